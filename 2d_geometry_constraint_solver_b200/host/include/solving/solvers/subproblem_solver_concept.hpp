// The static interface every sub-problem solver exposes (reference:
// src/solving/solvers/subproblem_solver_concept.hpp:31-36).
#pragma once

#include <concepts>

#include <gcs/model/gcs_data_structures.hpp>
#include <gcs/model/solve_result.hpp>

namespace Gcs::Solvers {

template <typename T>
concept SubproblemSolver = requires(const ConstraintGraph& constComponent, ConstraintGraph& component) {
    { T::matches(constComponent) } -> std::same_as<bool>;
    { T::solve(component) } -> std::same_as<SolveResult>;
};

}  // namespace Gcs::Solvers
