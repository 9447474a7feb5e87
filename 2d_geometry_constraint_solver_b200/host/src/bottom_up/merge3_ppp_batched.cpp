// merge3_ppp_batched.cpp — see solving/bottom_up/merge3_ppp_batched.hpp.
#include <limits>
#include <unordered_set>
#include <vector>

#include "solving/bottom_up/merge3_ppp_batched.hpp"

#include "merge3_cases.hpp"
#include "merge3_pass2.hpp"

namespace Gcs::B200 {

namespace Bu = Solvers::BottomUp;
using Eigen::Vector2d;
using NodeId = ConstraintGraph::NodeIdType;


namespace detail {

// ---- pass 1: the reference's enumeration, collecting the equation pair of every candidate ----
void collectPpp(const ConstraintGraph& sourceGraph, const std::array<const Bu::ClusterPose*, 3>& children, Merge3Batch& batch,
    std::vector<PppCandidate>& candidates)
{
    for (std::size_t referenceIndex = 0; referenceIndex < 3; ++referenceIndex) {
        std::array<std::size_t, 2> moving {};
        std::size_t at = 0;
        for (std::size_t index = 0; index < 3; ++index)
            if (index != referenceIndex) moving[at++] = index;  // merge3_ppp_solver.cpp:34-42
        const Bu::ClusterPose& referenceCluster = *children[referenceIndex];
        const Bu::ClusterPose& movingClusterA = *children[moving[0]];
        const Bu::ClusterPose& movingClusterB = *children[moving[1]];

        std::unordered_set<NodeId> referenceElements;
        referenceElements.reserve(referenceCluster.size());
        for (const auto& entry : referenceCluster) referenceElements.insert(entry.first);

        const auto sharedRefA = Bu::clusterIntersectionByType(sourceGraph, referenceCluster, movingClusterA, true);
        const auto sharedRefB = Bu::clusterIntersectionByType(sourceGraph, referenceCluster, movingClusterB, true);
        const auto sharedAB = Bu::clusterIntersectionByType(sourceGraph, movingClusterA, movingClusterB, true);
        std::vector<NodeId> freeCandidates;
        for (const auto& id : sharedAB)
            if (!referenceElements.contains(id)) freeCandidates.push_back(id);  // :64-69

        for (const auto& fixedPointA : sharedRefA) {
            for (const auto& fixedPointB : sharedRefB) {
                if (fixedPointA == fixedPointB) continue;
                const auto fixedAInGlobal = Bu::getPointPosition(referenceCluster, fixedPointA);
                const auto fixedBInGlobal = Bu::getPointPosition(referenceCluster, fixedPointB);
                if (!fixedAInGlobal || !fixedBInGlobal) continue;
                for (const auto& freePointId : freeCandidates) {
                    if (freePointId == fixedPointA || freePointId == fixedPointB) continue;
                    const auto fixedAInMovingA = Bu::getPointPosition(movingClusterA, fixedPointA);
                    const auto freeInMovingA = Bu::getPointPosition(movingClusterA, freePointId);
                    const auto fixedBInMovingB = Bu::getPointPosition(movingClusterB, fixedPointB);
                    const auto freeInMovingB = Bu::getPointPosition(movingClusterB, freePointId);
                    if (!fixedAInMovingA || !freeInMovingA || !fixedBInMovingB || !freeInMovingB) continue;
                    const double distanceAFree = (*fixedAInMovingA - *freeInMovingA).norm();  // :112-117
                    const double distanceBFree = (*fixedBInMovingB - *freeInMovingB).norm();
                    if (distanceAFree < Bu::EPSILON || distanceBFree < Bu::EPSILON) continue;
                    const auto fixedACanvas = Bu::getPointCanvasPosition(sourceGraph, fixedPointA);
                    const auto fixedBCanvas = Bu::getPointCanvasPosition(sourceGraph, fixedPointB);
                    const auto freeCanvas = Bu::getPointCanvasPosition(sourceGraph, freePointId);
                    if (!fixedACanvas || !fixedBCanvas || !freeCanvas) continue;
                    // :135-150: two pointToPointDistance equations, solve2D from the default guesses,
                    // pickByTriangleOrientation - one row of the K1 batch
                    const auto h = batch.addFreePointFromFixedPoints(
                        *fixedAInGlobal, *fixedBInGlobal, distanceAFree, distanceBFree, *fixedACanvas, *fixedBCanvas, *freeCanvas);
                    candidates.push_back({ referenceIndex, moving[0], moving[1], fixedPointA, fixedPointB, freePointId, *fixedAInGlobal,
                        *fixedBInGlobal, h });
                }
            }
        }
    }

}

// ---- pass 2: place, merge, score - every candidate on its own (merge3_pass2.hpp), the first best score wins ----
std::optional<Bu::ClusterPose> finishPpp(const ConstraintGraph& sourceGraph, const std::array<const Bu::ClusterPose*, 3>& children,
    const Merge3Batch& batch, const std::vector<PppCandidate>& candidates, std::size_t& scored, double& bestScore)
{
    const auto build = [&](std::size_t i) -> std::optional<Bu::ClusterPose> {
        const PppCandidate& c = candidates[i];
        const Vector2d selectedFreePoint = batch.point(c.handle).value();
        const Bu::ClusterPose& referenceCluster = *children[c.reference];
        const auto transformedA = Bu::transformClusterByTwoPointAnchors(*children[c.movingA], c.fixedA, c.free, c.fixedAInGlobal, selectedFreePoint);
        const auto transformedB = Bu::transformClusterByTwoPointAnchors(*children[c.movingB], c.fixedB, c.free, c.fixedBInGlobal, selectedFreePoint);
        if (!transformedA || !transformedB) return std::nullopt;
        Bu::ClusterPose merged = referenceCluster;  // :160-175
        merged[c.free] = Bu::PointPose { selectedFreePoint };
        for (const auto& [elementId, pose] : *transformedA)
            if (!merged.contains(elementId)) merged.emplace(elementId, pose);
        for (const auto& [elementId, pose] : *transformedB)
            if (!merged.contains(elementId)) merged.emplace(elementId, pose);
        return merged;
    };
    return pickBestMergedPose(sourceGraph, candidates.size(), build, scored, bestScore);
}

}  // namespace detail

std::optional<Bu::ClusterPose> solveMerge3Ppp(const ConstraintGraph& sourceGraph, const std::array<const Bu::ClusterPose*, 3>& children,
    int device, Merge3PppReport* report)
{
    Merge3Batch batch;
    std::vector<detail::PppCandidate> candidates;
    detail::collectPpp(sourceGraph, children, batch, candidates);
    if (!candidates.empty()) batch.solve(device);  // every candidate's Newton solve + root selection: one launch
    std::size_t scored = 0;
    double bestScore = std::numeric_limits<double>::infinity();
    std::optional<Bu::ClusterPose> bestMergedPose = detail::finishPpp(sourceGraph, children, batch, candidates, scored, bestScore);
    if (report) {
        report->candidates = candidates.size();
        report->scored = scored;
        report->launches = batch.launches();
        report->bestScore = bestScore;
    }
    return bestMergedPose;
}

}  // namespace Gcs::B200
