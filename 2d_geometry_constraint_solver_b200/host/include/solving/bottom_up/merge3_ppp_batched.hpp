// The pairwise-anchor PPP merge of the bottom-up plan solver as a consumer of the batched kernels
// (reference: src/constraint_solver/src/solving/bottom_up/merge3_ppp_solver.cpp:18-214,
// Merge3PppSolver::solve).
//
// The reference enumerates, for each of the three children taken as the reference cluster, every
// (fixed point shared with moving cluster A) x (fixed point shared with moving cluster B) x (free
// point shared by A and B outside the reference), and for EACH candidate calls solve2D on two
// point-to-point distances (:135-143) and pickByTriangleOrientation (:145-150), then places both
// moving clusters by a two-point anchor fit, merges and scores (:152-192); the best score wins,
// the first one on ties (:188).  Here the enumeration runs twice over the same candidate order:
// once to collect every candidate's equation pair in a Gcs::B200::Merge3Batch (one kernel launch
// for the whole merge instead of one Newton solve per candidate), once to place / merge / score
// with the solved points.  Everything but the Newton solves is the host arithmetic of
// merge3_solver_common.hpp; results are those of the reference loop, bit for bit.
//
// The reference passes the three children through its plan-tree context (Merge3Context: plan node,
// child ids, pose map); the Merge3 solver classes and the plan builder are outside the accelerated
// path, so this entry point takes the three cluster poses directly.
#pragma once

#include <array>
#include <cstddef>
#include <optional>

#include <gcs/export.hpp>
#include <gcs/model/gcs_data_structures.hpp>

#include "solving/bottom_up/merge3_solver_common.hpp"

namespace Gcs::B200 {

struct Merge3PppReport {
    std::size_t candidates = 0;  // candidates that reached the Newton solve (the reference's attemptedCandidates counts those that also placed)
    std::size_t scored = 0;      // candidates placed, merged and scored
    std::size_t launches = 0;    // kernel launches (1 when there is any candidate)
    double bestScore = 0.0;
};

GCS_API std::optional<Solvers::BottomUp::ClusterPose> solveMerge3Ppp(const ConstraintGraph& sourceGraph,
    const std::array<const Solvers::BottomUp::ClusterPose*, 3>& children, int device = 0, Merge3PppReport* report = nullptr);

}  // namespace Gcs::B200
