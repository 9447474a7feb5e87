// merge3_batched.cpp — see solving/bottom_up/merge3_batched.hpp.
#include <limits>
#include <stdexcept>
#include <unordered_set>
#include <utility>
#include <vector>

#include "solving/bottom_up/merge3_batched.hpp"

#include "merge3_cases.hpp"
#include "merge3_pass2.hpp"

namespace Gcs::B200 {

namespace Bu = Solvers::BottomUp;
using Eigen::Vector2d;
using NodeId = ConstraintGraph::NodeIdType;

namespace {

// One candidate of an enumeration: what pass 2 needs to place both moving clusters.  Each moving
// cluster is fitted onto two anchors - the fixed element it shares with the reference, at its pose
// in the reference, and the free element, at its solved pose - in that order, as the reference
// builds its anchor arrays (PLL :109-139, LPP :120-149, LLP :106-135).
struct Candidate {
    std::size_t reference, movingA, movingB;
    NodeId fixedA, fixedB, free;
    Bu::ElementPose fixedAPose, fixedBPose;
    bool freeIsLine;
    Merge3Batch::Handle handle;
};

struct Enumeration {
    std::vector<Candidate> candidates;
    std::size_t skippedDegenerate = 0;
};

std::array<std::size_t, 2> movingOf(std::size_t referenceIndex)
{
    std::array<std::size_t, 2> moving {};
    std::size_t at = 0;
    for (std::size_t index = 0; index < 3; ++index)
        if (index != referenceIndex) moving[at++] = index;
    return moving;
}

std::unordered_set<NodeId> idsOf(const Bu::ClusterPose& cluster)
{
    std::unordered_set<NodeId> ids;
    ids.reserve(cluster.size());
    for (const auto& entry : cluster) ids.insert(entry.first);
    return ids;
}

// ---- pass 1 of each case: the reference's loops, every candidate's equation pair into `batch` ----

// merge3_pll_solver.cpp:25-107
void collectPll(const ConstraintGraph& g, const Merge3Children& children, Merge3Batch& batch, Enumeration& out)
{
    for (std::size_t referenceIndex = 0; referenceIndex < 3; ++referenceIndex) {
        const auto moving = movingOf(referenceIndex);
        const Bu::ClusterPose& referenceCluster = *children[referenceIndex];
        const Bu::ClusterPose& movingClusterA = *children[moving[0]];
        const Bu::ClusterPose& movingClusterB = *children[moving[1]];
        const auto referenceElements = idsOf(referenceCluster);
        const auto sharedRefAPoints = Bu::clusterIntersectionByType(g, referenceCluster, movingClusterA, true);
        const auto sharedRefBPoints = Bu::clusterIntersectionByType(g, referenceCluster, movingClusterB, true);
        const auto sharedABLines = Bu::clusterIntersectionByType(g, movingClusterA, movingClusterB, false);
        std::vector<NodeId> freeLineCandidates;
        for (const auto& lineId : sharedABLines)
            if (!referenceElements.contains(lineId)) freeLineCandidates.push_back(lineId);

        for (const auto& fixedPointA : sharedRefAPoints) {
            for (const auto& fixedPointB : sharedRefBPoints) {
                if (fixedPointA == fixedPointB) continue;
                const auto fixedAInGlobal = Bu::getPointPosition(referenceCluster, fixedPointA);
                const auto fixedBInGlobal = Bu::getPointPosition(referenceCluster, fixedPointB);
                const auto fixedACanvas = Bu::getPointCanvasPosition(g, fixedPointA);
                const auto fixedBCanvas = Bu::getPointCanvasPosition(g, fixedPointB);
                if (!fixedAInGlobal || !fixedBInGlobal || !fixedACanvas || !fixedBCanvas) continue;
                for (const auto& freeLineId : freeLineCandidates) {
                    const auto freeLineCanvas = Bu::getLineCanvasPose(g, freeLineId);
                    const auto freeLineInMovingA = Bu::getLinePosition(movingClusterA, freeLineId);
                    const auto freeLineInMovingB = Bu::getLinePosition(movingClusterB, freeLineId);
                    const auto fixedAInMovingA = Bu::getPointPosition(movingClusterA, fixedPointA);
                    const auto fixedBInMovingB = Bu::getPointPosition(movingClusterB, fixedPointB);
                    if (!freeLineCanvas || !freeLineInMovingA || !freeLineInMovingB || !fixedAInMovingA || !fixedBInMovingB) continue;
                    const double distanceA = Bu::pointToLineDistanceAbs(*fixedAInMovingA, *freeLineInMovingA);
                    const double distanceB = Bu::pointToLineDistanceAbs(*fixedBInMovingB, *freeLineInMovingB);
                    const auto h = batch.addFreeLineFromFixedPoints(
                        *fixedAInGlobal, *fixedBInGlobal, distanceA, distanceB, *fixedACanvas, *fixedBCanvas, *freeLineCanvas);
                    out.candidates.push_back({ referenceIndex, moving[0], moving[1], fixedPointA, fixedPointB, freeLineId,
                        Bu::PointPose { *fixedAInGlobal }, Bu::PointPose { *fixedBInGlobal }, true, h });
                }
            }
        }
    }
}

// merge3_lpp_solver.cpp:25-118: moving cluster A is the one sharing a POINT with the reference, B the
// one sharing a LINE; both assignments of the two moving children are tried
void collectLpp(const ConstraintGraph& g, const Merge3Children& children, Merge3Batch& batch, Enumeration& out)
{
    for (std::size_t referenceIndex = 0; referenceIndex < 3; ++referenceIndex) {
        const auto base = movingOf(referenceIndex);
        for (const auto& moving : std::array<std::array<std::size_t, 2>, 2> { base, std::array<std::size_t, 2> { base[1], base[0] } }) {
            const Bu::ClusterPose& referenceCluster = *children[referenceIndex];
            const Bu::ClusterPose& pointCluster = *children[moving[0]];
            const Bu::ClusterPose& lineCluster = *children[moving[1]];
            const auto referenceElements = idsOf(referenceCluster);
            const auto sharedRefPoints = Bu::clusterIntersectionByType(g, referenceCluster, pointCluster, true);
            const auto sharedRefLines = Bu::clusterIntersectionByType(g, referenceCluster, lineCluster, false);
            const auto sharedFreePoints = Bu::clusterIntersectionByType(g, pointCluster, lineCluster, true);

            for (const auto& fixedPointId : sharedRefPoints) {
                for (const auto& fixedLineId : sharedRefLines) {
                    const auto fixedPointGlobal = Bu::getPointPosition(referenceCluster, fixedPointId);
                    const auto fixedLineGlobal = Bu::getLinePosition(referenceCluster, fixedLineId);
                    const auto fixedPointCanvas = Bu::getPointCanvasPosition(g, fixedPointId);
                    const auto fixedLineCanvas = Bu::getLineCanvasPose(g, fixedLineId);
                    if (!fixedPointGlobal || !fixedLineGlobal || !fixedPointCanvas || !fixedLineCanvas) continue;
                    for (const auto& freePointId : sharedFreePoints) {
                        if (referenceElements.contains(freePointId)) continue;
                        const auto freePointInPointCluster = Bu::getPointPosition(pointCluster, freePointId);
                        const auto fixedPointInPointCluster = Bu::getPointPosition(pointCluster, fixedPointId);
                        const auto freePointInLineCluster = Bu::getPointPosition(lineCluster, freePointId);
                        const auto fixedLineInLineCluster = Bu::getLinePosition(lineCluster, fixedLineId);
                        const auto freePointCanvas = Bu::getPointCanvasPosition(g, freePointId);
                        if (!freePointInPointCluster || !fixedPointInPointCluster || !freePointInLineCluster || !fixedLineInLineCluster
                            || !freePointCanvas)
                            continue;
                        const double distanceToPoint = (*freePointInPointCluster - *fixedPointInPointCluster).norm();
                        const double distanceToLine = Bu::pointToLineDistanceAbs(*freePointInLineCluster, *fixedLineInLineCluster);
                        Merge3Batch::Handle h;
                        try {
                            h = batch.addFreePointFromFixedPointAndLine(*fixedPointGlobal, *fixedLineGlobal, distanceToPoint, distanceToLine,
                                *fixedPointCanvas, *fixedLineCanvas, *freePointCanvas);
                        } catch (const std::domain_error&) {  // degenerate fixed line: see the header
                            ++out.skippedDegenerate;
                            continue;
                        }
                        out.candidates.push_back({ referenceIndex, moving[0], moving[1], fixedPointId, fixedLineId, freePointId,
                            Bu::PointPose { *fixedPointGlobal }, *fixedLineGlobal, false, h });
                    }
                }
            }
        }
    }
}

// merge3_llp_solver.cpp:25-104
void collectLlp(const ConstraintGraph& g, const Merge3Children& children, Merge3Batch& batch, Enumeration& out)
{
    for (std::size_t referenceIndex = 0; referenceIndex < 3; ++referenceIndex) {
        const auto moving = movingOf(referenceIndex);
        const Bu::ClusterPose& referenceCluster = *children[referenceIndex];
        const Bu::ClusterPose& movingClusterA = *children[moving[0]];
        const Bu::ClusterPose& movingClusterB = *children[moving[1]];
        const auto referenceElements = idsOf(referenceCluster);
        const auto sharedRefALines = Bu::clusterIntersectionByType(g, referenceCluster, movingClusterA, false);
        const auto sharedRefBLines = Bu::clusterIntersectionByType(g, referenceCluster, movingClusterB, false);
        const auto sharedABPoints = Bu::clusterIntersectionByType(g, movingClusterA, movingClusterB, true);

        for (const auto& fixedLineAId : sharedRefALines) {
            for (const auto& fixedLineBId : sharedRefBLines) {
                if (fixedLineAId == fixedLineBId) continue;
                const auto fixedLineAGlobal = Bu::getLinePosition(referenceCluster, fixedLineAId);
                const auto fixedLineBGlobal = Bu::getLinePosition(referenceCluster, fixedLineBId);
                const auto fixedLineACanvas = Bu::getLineCanvasPose(g, fixedLineAId);
                const auto fixedLineBCanvas = Bu::getLineCanvasPose(g, fixedLineBId);
                if (!fixedLineAGlobal || !fixedLineBGlobal || !fixedLineACanvas || !fixedLineBCanvas) continue;
                for (const auto& freePointId : sharedABPoints) {
                    if (referenceElements.contains(freePointId)) continue;
                    const auto freePointInA = Bu::getPointPosition(movingClusterA, freePointId);
                    const auto freePointInB = Bu::getPointPosition(movingClusterB, freePointId);
                    const auto fixedLineAInA = Bu::getLinePosition(movingClusterA, fixedLineAId);
                    const auto fixedLineBInB = Bu::getLinePosition(movingClusterB, fixedLineBId);
                    const auto freePointCanvas = Bu::getPointCanvasPosition(g, freePointId);
                    if (!freePointInA || !freePointInB || !fixedLineAInA || !fixedLineBInB || !freePointCanvas) continue;
                    const double distanceToA = Bu::pointToLineDistanceAbs(*freePointInA, *fixedLineAInA);
                    const double distanceToB = Bu::pointToLineDistanceAbs(*freePointInB, *fixedLineBInB);
                    Merge3Batch::Handle h;
                    try {
                        h = batch.addFreePointFromFixedLines(*fixedLineAGlobal, *fixedLineBGlobal, distanceToA, distanceToB, *fixedLineACanvas,
                            *fixedLineBCanvas, *freePointCanvas);
                    } catch (const std::domain_error&) {
                        ++out.skippedDegenerate;
                        continue;
                    }
                    out.candidates.push_back({ referenceIndex, moving[0], moving[1], fixedLineAId, fixedLineBId, freePointId, *fixedLineAGlobal,
                        *fixedLineBGlobal, false, h });
                }
            }
        }
    }
}

// ---- pass 2, the same for the three cases: place, merge, score - in the enumeration's order ----
std::optional<Bu::ClusterPose> finish(
    const ConstraintGraph& g, const Merge3Children& children, const Merge3Batch& batch, const Enumeration& e, Merge3Report* report)
{
    const auto build = [&](std::size_t i) -> std::optional<Bu::ClusterPose> {
        const Candidate& c = e.candidates[i];
        Bu::ElementPose solved;
        if (c.freeIsLine) {
            const auto line = batch.line(c.handle);
            if (!line) return std::nullopt;  // the helper's std::nullopt
            solved = *line;
        } else {
            const auto point = batch.point(c.handle);
            if (!point) return std::nullopt;
            solved = Bu::PointPose { *point };
        }
        const std::array<std::pair<NodeId, Bu::ElementPose>, 2> anchorsA { std::pair { c.fixedA, c.fixedAPose }, std::pair { c.free, solved } };
        const std::array<std::pair<NodeId, Bu::ElementPose>, 2> anchorsB { std::pair { c.fixedB, c.fixedBPose }, std::pair { c.free, solved } };
        const auto transformedA = Bu::transformClusterByAnchors(*children[c.movingA], anchorsA);
        const auto transformedB = Bu::transformClusterByAnchors(*children[c.movingB], anchorsB);
        if (!transformedA || !transformedB) return std::nullopt;
        Bu::ClusterPose merged = *children[c.reference];
        merged[c.free] = solved;
        for (const auto& [elementId, pose] : *transformedA)
            if (!merged.contains(elementId)) merged.emplace(elementId, pose);
        for (const auto& [elementId, pose] : *transformedB)
            if (!merged.contains(elementId)) merged.emplace(elementId, pose);
        return merged;
    };
    std::size_t scored = 0;
    double bestScore = std::numeric_limits<double>::infinity();
    auto bestMergedPose = detail::pickBestMergedPose(g, e.candidates.size(), build, scored, bestScore);  // on every host thread
    if (report) {
        report->candidates = e.candidates.size();
        report->scored = scored;
        report->bestScore = bestScore;
    }
    return bestMergedPose;
}

using Collect = void (*)(const ConstraintGraph&, const Merge3Children&, Merge3Batch&, Enumeration&);

std::optional<Bu::ClusterPose> solveOne(
    Collect collect, const ConstraintGraph& g, const Merge3Children& children, int device, Merge3Report* report)
{
    Merge3Batch batch;
    Enumeration e;
    collect(g, children, batch, e);
    if (!e.candidates.empty()) batch.solve(device);
    auto merged = finish(g, children, batch, e, report);
    if (report) report->launches = batch.launches();
    return merged;
}

}  // namespace

std::optional<Bu::ClusterPose> solveMerge3Pll(const ConstraintGraph& g, const Merge3Children& children, int device, Merge3Report* report)
{
    return solveOne(collectPll, g, children, device, report);
}

std::optional<Bu::ClusterPose> solveMerge3Lpp(const ConstraintGraph& g, const Merge3Children& children, int device, Merge3Report* report)
{
    return solveOne(collectLpp, g, children, device, report);
}

std::optional<Bu::ClusterPose> solveMerge3Llp(const ConstraintGraph& g, const Merge3Children& children, int device, Merge3Report* report)
{
    return solveOne(collectLlp, g, children, device, report);
}

bool detectUnsolvableMerge3Lll(const ConstraintGraph& g, const Merge3Children& children)
{
    for (std::size_t referenceIndex = 0; referenceIndex < 3; ++referenceIndex) {
        const auto moving = movingOf(referenceIndex);
        const Bu::ClusterPose& referenceCluster = *children[referenceIndex];
        const Bu::ClusterPose& movingClusterA = *children[moving[0]];
        const Bu::ClusterPose& movingClusterB = *children[moving[1]];
        const auto referenceElements = idsOf(referenceCluster);
        const auto sharedRefALines = Bu::clusterIntersectionByType(g, referenceCluster, movingClusterA, false);
        const auto sharedRefBLines = Bu::clusterIntersectionByType(g, referenceCluster, movingClusterB, false);
        const auto sharedABLines = Bu::clusterIntersectionByType(g, movingClusterA, movingClusterB, false);
        if (sharedRefALines.empty() || sharedRefBLines.empty()) continue;
        for (const auto& freeLineId : sharedABLines)
            if (!referenceElements.contains(freeLineId)) return true;
    }
    return false;
}

std::optional<Bu::ClusterPose> solveMerge3Fallback(const Merge3Children& children)
{
    const auto firstMerge = Bu::mergeChildClusterIntoReference(*children[0], *children[1]);
    if (!firstMerge) return std::nullopt;
    return Bu::mergeChildClusterIntoReference(*firstMerge, *children[2]);
}

std::optional<Bu::ClusterPose> solveMerge3Node(const ConstraintGraph& g, const Merge3Children& children, int device, Merge3NodeReport* report)
{
    Merge3NodeReport local;
    Merge3NodeReport& r = report ? *report : local;
    r = Merge3NodeReport {};

    Merge3PppReport ppp;
    if (auto merged = solveMerge3Ppp(g, children, device, &ppp)) {
        r = { Merge3Case::Ppp, ppp.candidates, ppp.scored, ppp.launches, ppp.bestScore };
        return merged;
    }
    r.candidates = ppp.candidates, r.launches = ppp.launches;

    // the three cases with a line in them: one batch, read back in the reference's order
    Merge3Batch batch;
    std::array<Enumeration, 3> e;
    collectPll(g, children, batch, e[0]);
    collectLpp(g, children, batch, e[1]);
    collectLlp(g, children, batch, e[2]);
    if (batch.size() != 0) batch.solve(device);
    r.launches += batch.launches();
    for (std::size_t k = 0; k < 3; ++k) {
        Merge3Report one;
        auto merged = finish(g, children, batch, e[k], &one);
        r.candidates += one.candidates;
        if (merged) {
            r.solvedBy = static_cast<Merge3Case>(static_cast<int>(Merge3Case::Pll) + static_cast<int>(k));
            r.scored = one.scored;
            r.bestScore = one.bestScore;
            return merged;
        }
    }
    if (detectUnsolvableMerge3Lll(g, children)) {
        r.solvedBy = Merge3Case::Unsolvable;
        return std::nullopt;
    }
    auto fallback = solveMerge3Fallback(children);
    r.solvedBy = fallback ? Merge3Case::Fallback : Merge3Case::Unsolvable;
    return fallback;
}

std::vector<std::optional<Bu::ClusterPose>> solveMerge3Level(
    std::span<const Merge3NodeInput> nodes, int device, std::vector<Merge3NodeReport>* reports, Merge3LevelReport* level)
{
    struct NodeWork {
        std::vector<detail::PppCandidate> ppp;
        std::array<Enumeration, 3> line;  // PLL, LPP, LLP: collected only where PPP has no candidate
        bool lineCollected = false;
    };
    Merge3Batch batch;
    std::vector<NodeWork> work(nodes.size());
    for (std::size_t k = 0; k < nodes.size(); ++k) {
        const ConstraintGraph& g = *nodes[k].sourceGraph;
        detail::collectPpp(g, nodes[k].children, batch, work[k].ppp);
        if (!work[k].ppp.empty()) continue;
        collectPll(g, nodes[k].children, batch, work[k].line[0]);
        collectLpp(g, nodes[k].children, batch, work[k].line[1]);
        collectLlp(g, nodes[k].children, batch, work[k].line[2]);
        work[k].lineCollected = true;
    }
    if (batch.size() != 0) batch.solve(device);
    std::size_t launches = batch.launches(), candidates = 0;

    std::vector<std::optional<Bu::ClusterPose>> out(nodes.size());
    std::vector<Merge3NodeReport> local(nodes.size());
    for (std::size_t k = 0; k < nodes.size(); ++k) {
        const ConstraintGraph& g = *nodes[k].sourceGraph;
        const Merge3Children& children = nodes[k].children;
        Merge3NodeReport& r = local[k];
        r.candidates = work[k].ppp.size();
        if (!work[k].ppp.empty()) {
            std::size_t scored = 0;
            double best = 0.0;
            if ((out[k] = detail::finishPpp(g, children, batch, work[k].ppp, scored, best))) {
                r.solvedBy = Merge3Case::Ppp, r.scored = scored, r.bestScore = best;
                candidates += r.candidates;
                continue;
            }
        }
        const Merge3Batch* lineBatch = &batch;
        Merge3Batch own;
        if (!work[k].lineCollected) {  // PPP had candidates and none placed: this node's line cases on their own
            collectPll(g, children, own, work[k].line[0]);
            collectLpp(g, children, own, work[k].line[1]);
            collectLlp(g, children, own, work[k].line[2]);
            if (own.size() != 0) own.solve(device);
            launches += own.launches();
            lineBatch = &own;
        }
        bool done = false;
        for (std::size_t c = 0; c < 3 && !done; ++c) {
            Merge3Report one;
            out[k] = finish(g, children, *lineBatch, work[k].line[c], &one);
            r.candidates += one.candidates;
            if (out[k]) {
                r.solvedBy = static_cast<Merge3Case>(static_cast<int>(Merge3Case::Pll) + static_cast<int>(c));
                r.scored = one.scored, r.bestScore = one.bestScore;
                done = true;
            }
        }
        candidates += r.candidates;
        if (done) continue;
        if (detectUnsolvableMerge3Lll(g, children)) {
            r.solvedBy = Merge3Case::Unsolvable;
            continue;
        }
        out[k] = solveMerge3Fallback(children);
        r.solvedBy = out[k] ? Merge3Case::Fallback : Merge3Case::Unsolvable;
    }
    if (reports) *reports = std::move(local);
    if (level) *level = { nodes.size(), candidates, launches };
    return out;
}

}  // namespace Gcs::B200
