// classifyAndSolve: first-match dispatch over the eight sub-problem solvers (reference:
// src/solving/component_solver.hpp:31-66, same order, same Unsupported message).
#pragma once

#include <gcs/b200/leaf_batch.hpp>
#include <gcs/model/gcs_data_structures.hpp>
#include <gcs/model/solve_result.hpp>
#include "solving/solvers/line_angle_solvers.hpp"
#include "solving/solvers/point_line_solvers.hpp"
#include "solving/solvers/point_point_solvers.hpp"

namespace Gcs {

inline SolveResult classifyAndSolve(ConstraintGraph& component)
{
    const B200::SolverId id = B200::classify(component);
    if (id == B200::SolverId::None) return SolveResult::unsupported("No solver matches this component configuration");
    return B200::solveSingle(id, component);
}

}  // namespace Gcs
