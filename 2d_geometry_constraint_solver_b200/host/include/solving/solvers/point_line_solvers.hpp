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

// point_line_solvers.cpp:114-255: two unsolved points and an unsolved line, three distances
struct GCS_API ZeroFixedPPLTriangleSolver {
    static bool matches(const ConstraintGraph& component) { return B200::matches(B200::SolverId::ZeroFixedPPLTriangle, component); }
    static SolveResult solve(ConstraintGraph& component) { return B200::solveSingle(B200::SolverId::ZeroFixedPPLTriangle, component); }
};
static_assert(SubproblemSolver<ZeroFixedPPLTriangleSolver>);

// point_line_solvers.cpp:261-399: two solved points, free line at given distances
struct GCS_API TwoFixedPointsLineSolver {
    static bool matches(const ConstraintGraph& component) { return B200::matches(B200::SolverId::TwoFixedPointsLine, component); }
    static SolveResult solve(ConstraintGraph& component) { return B200::solveSingle(B200::SolverId::TwoFixedPointsLine, component); }
};
static_assert(SubproblemSolver<TwoFixedPointsLineSolver>);

// point_line_solvers.cpp:405-541: solved point + solved line, free point
struct GCS_API FixedPointAndLineFreePointSolver {
    static bool matches(const ConstraintGraph& component) { return B200::matches(B200::SolverId::FixedPointAndLineFreePoint, component); }
    static SolveResult solve(ConstraintGraph& component) { return B200::solveSingle(B200::SolverId::FixedPointAndLineFreePoint, component); }
};
static_assert(SubproblemSolver<FixedPointAndLineFreePointSolver>);

// point_line_solvers.cpp:547-695: two solved lines, free point
struct GCS_API TwoFixedLinesFreePointSolver {
    static bool matches(const ConstraintGraph& component) { return B200::matches(B200::SolverId::TwoFixedLinesFreePoint, component); }
    static SolveResult solve(ConstraintGraph& component) { return B200::solveSingle(B200::SolverId::TwoFixedLinesFreePoint, component); }
};
static_assert(SubproblemSolver<TwoFixedLinesFreePointSolver>);

}  // namespace Gcs::Solvers
