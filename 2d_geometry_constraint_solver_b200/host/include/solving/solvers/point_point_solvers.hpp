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

// point_point_solvers.cpp:14-85: three unsolved points, three distances; anchors P1 at the origin and P2 on the x axis
struct GCS_API ZeroFixedPointsTriangleSolver {
    static bool matches(const ConstraintGraph& component) { return B200::matches(B200::SolverId::ZeroFixedPointsTriangle, component); }
    static SolveResult solve(ConstraintGraph& component) { return B200::solveSingle(B200::SolverId::ZeroFixedPointsTriangle, component); }
};
static_assert(SubproblemSolver<ZeroFixedPointsTriangleSolver>);

// point_point_solvers.cpp:87-164: two solved points, one free point at given distances
struct GCS_API TwoFixedPointsDistanceSolver {
    static bool matches(const ConstraintGraph& component) { return B200::matches(B200::SolverId::TwoFixedPointsDistance, component); }
    static SolveResult solve(ConstraintGraph& component) { return B200::solveSingle(B200::SolverId::TwoFixedPointsDistance, component); }
};
static_assert(SubproblemSolver<TwoFixedPointsDistanceSolver>);

}  // namespace Gcs::Solvers
