// The Merge3 step of the bottom-up plan solver as a consumer of the batched kernels: the three
// enumeration loops with a line in them, the rigid fallback, and the case order of a merge node
// (reference: src/constraint_solver/src/solving/bottom_up/merge3_pll_solver.cpp:15-189,
// merge3_lpp_solver.cpp:15-208, merge3_llp_solver.cpp:15-190, merge3_fallback_solver.cpp:13-78,
// and the dispatch in src/solving/bottom_up_plan_solver.cpp:393-431; the PPP loop is
// merge3_ppp_batched.hpp).
//
// Every loop has the reference's shape: for each child taken as the reference cluster, (a fixed
// element it shares with moving cluster A) x (one it shares with B) x (a free element A and B share
// outside the reference); per candidate ONE solve2D through a numeric helper -
//     PLL  two fixed points, free line      solveFreeLineFromFixedPoints          (K2, M3C:480-531)
//     LPP  fixed point + fixed line, point  solveFreePointFromFixedPointAndLine   (K3, M3C:533-562)
//     LLP  two fixed lines, free point      solveFreePointFromFixedLines          (K4, M3C:564-608)
// - then both moving clusters are placed by a two-anchor rigid fit, merged into the reference and
// scored; the best score wins, the first one on ties.  Here each loop runs twice over the same
// candidate order: pass 1 packs every candidate's equation pair into a Gcs::B200::Merge3Batch (one
// kernel launch per kind for the whole merge instead of one Newton solve per candidate), pass 2
// places / merges / scores with the solved elements.  Everything but the Newton solves is the host
// arithmetic of merge3_solver_common.hpp; results are those of the reference loops bit for bit
// (tests/test_merge3.py runs them against the reference's own solver classes).
//
// One difference, inherited from the numeric helpers: a candidate whose FIXED line is shorter than
// EPSILON is skipped (the reference solves a rank-deficient system there and scores whatever comes out).
#pragma once

#include <array>
#include <cstddef>
#include <optional>
#include <span>
#include <vector>

#include <gcs/export.hpp>
#include <gcs/model/gcs_data_structures.hpp>

#include "solving/bottom_up/merge3_ppp_batched.hpp"
#include "solving/bottom_up/merge3_solver_common.hpp"

namespace Gcs::B200 {

using Merge3Children = std::array<const Solvers::BottomUp::ClusterPose*, 3>;

struct Merge3Report {
    std::size_t candidates = 0;  // candidates packed for the Newton solve
    std::size_t scored = 0;      // candidates placed, merged and scored
    std::size_t launches = 0;    // kernel launches (one per equation-pair kind with candidates)
    double bestScore = 0.0;
};

GCS_API std::optional<Solvers::BottomUp::ClusterPose> solveMerge3Pll(
    const ConstraintGraph& sourceGraph, const Merge3Children& children, int device = 0, Merge3Report* report = nullptr);
GCS_API std::optional<Solvers::BottomUp::ClusterPose> solveMerge3Lpp(
    const ConstraintGraph& sourceGraph, const Merge3Children& children, int device = 0, Merge3Report* report = nullptr);
GCS_API std::optional<Solvers::BottomUp::ClusterPose> solveMerge3Llp(
    const ConstraintGraph& sourceGraph, const Merge3Children& children, int device = 0, Merge3Report* report = nullptr);

// merge3_fallback_solver.cpp:13-59: two fixed lines in the reference and a free LINE shared by the
// moving clusters - three lines fix no rigid placement
GCS_API bool detectUnsolvableMerge3Lll(const ConstraintGraph& sourceGraph, const Merge3Children& children);
// merge3_fallback_solver.cpp:61-78: child 1, then child 2, fitted onto child 0 over what they share
GCS_API std::optional<Solvers::BottomUp::ClusterPose> solveMerge3Fallback(const Merge3Children& children);

// What a merge node does (bottom_up_plan_solver.cpp:393-431): the first case, in this order, that
// produces a pose.  PPP first with its own launch (the common case); when it has no candidate the
// three line cases are enumerated together into ONE batch (at most one launch per kind K2 / K3 / K4)
// and read back in the reference's order, so the outcome is the sequential dispatch's.
enum class Merge3Case { Ppp = 0, Pll = 1, Lpp = 2, Llp = 3, Fallback = 4, Unsolvable = 5 };

struct Merge3NodeReport {
    Merge3Case solvedBy = Merge3Case::Unsolvable;
    std::size_t candidates = 0, scored = 0, launches = 0;
    double bestScore = 0.0;
};

GCS_API std::optional<Solvers::BottomUp::ClusterPose> solveMerge3Node(
    const ConstraintGraph& sourceGraph, const Merge3Children& children, int device = 0, Merge3NodeReport* report = nullptr);

// Every merge node of one level of a plan tree (nodes of a level do not depend on each other) through
// ONE Merge3Batch: pass 1 of every node - PPP, and the three line cases where PPP has no candidate -
// then at most one launch per equation-pair kind for the whole level, then pass 2 node by node in
// the reference's case order.  Results are solveMerge3Node's; what changes is the number of launches:
// 4 per level at most instead of up to 4 per node.  (A node whose PPP candidates all fail to place -
// possible, never seen - falls back to its own batch for the line cases.)
struct Merge3NodeInput {
    const ConstraintGraph* sourceGraph = nullptr;
    Merge3Children children {};
};

struct Merge3LevelReport {
    std::size_t nodes = 0, candidates = 0, launches = 0;
};

GCS_API std::vector<std::optional<Solvers::BottomUp::ClusterPose>> solveMerge3Level(std::span<const Merge3NodeInput> nodes, int device = 0,
    std::vector<Merge3NodeReport>* reports = nullptr, Merge3LevelReport* level = nullptr);

}  // namespace Gcs::B200
