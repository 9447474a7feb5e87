// The two passes of the PPP enumeration (merge3_ppp_batched.cpp) as separate steps over a caller-owned
// Merge3Batch, so that several merge nodes can share one batch (solveMerge3Level, merge3_batched.cpp).
#pragma once

#include <array>
#include <cstddef>
#include <optional>
#include <vector>

#include "solving/bottom_up/merge3_solver_common.hpp"

namespace Gcs::B200::detail {

// One candidate of the enumeration (merge3_ppp_solver.cpp:77-97): everything the second pass needs.
struct PppCandidate {
    std::size_t reference, movingA, movingB;
    ConstraintGraph::NodeIdType fixedA, fixedB, free;
    Eigen::Vector2d fixedAInGlobal, fixedBInGlobal;
    Merge3Batch::Handle handle;
};

void collectPpp(const ConstraintGraph& sourceGraph, const std::array<const Solvers::BottomUp::ClusterPose*, 3>& children, Merge3Batch& batch,
    std::vector<PppCandidate>& candidates);
std::optional<Solvers::BottomUp::ClusterPose> finishPpp(const ConstraintGraph& sourceGraph,
    const std::array<const Solvers::BottomUp::ClusterPose*, 3>& children, const Merge3Batch& batch, const std::vector<PppCandidate>& candidates,
    std::size_t& scored, double& bestScore);

}  // namespace Gcs::B200::detail
