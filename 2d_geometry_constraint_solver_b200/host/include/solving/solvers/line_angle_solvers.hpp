// Sub-problem solvers with the reference's names and static interface; matches() evaluates the
// reference's predicate, solve() packs the leaf, runs it through the CUDA path (a batch of one;
// many leaves at once go through Gcs::B200::solveLeaves) and writes the result back.
#pragma once

#include <gcs/b200/leaf_batch.hpp>
#include <gcs/export.hpp>
#include <gcs/model/gcs_data_structures.hpp>
#include <gcs/model/solve_result.hpp>
#include "solving/solvers/subproblem_solver_concept.hpp"

namespace Gcs::Solvers {

// line_angle_solvers.cpp:169-371: two unsolved lines with an angle, one point at distances
struct GCS_API ZeroFixedLLPAngleTriangleSolver {
    static bool matches(const ConstraintGraph& component) { return B200::matches(B200::SolverId::ZeroFixedLLPAngleTriangle, component); }
    static SolveResult solve(ConstraintGraph& component) { return B200::solveSingle(B200::SolverId::ZeroFixedLLPAngleTriangle, component); }
};
static_assert(SubproblemSolver<ZeroFixedLLPAngleTriangleSolver>);

// line_angle_solvers.cpp:377-567: solved line + solved point, free line at an angle and a distance
struct GCS_API FixedLineAndPointFreeLineSolver {
    static bool matches(const ConstraintGraph& component) { return B200::matches(B200::SolverId::FixedLineAndPointFreeLine, component); }
    static SolveResult solve(ConstraintGraph& component) { return B200::solveSingle(B200::SolverId::FixedLineAndPointFreeLine, component); }
};
static_assert(SubproblemSolver<FixedLineAndPointFreeLineSolver>);

}  // namespace Gcs::Solvers
