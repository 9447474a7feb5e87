// Numeric helpers of the bottom-up Merge3 solvers, with the reference's names and signatures
// (reference: src/constraint_solver/src/solving/bottom_up/merge3_solver_common.hpp:17-93,
// merge3_solver_common.cpp).  SURVEY.md section 8f rank 3: the second consumer of the batched
// Newton kernels.
//
// The three solveFree* helpers (merge3_solver_common.cpp:480-608) and the point-from-two-points
// step of Merge3PppSolver (merge3_ppp_solver.cpp:135-153) are the reference's K2 / K3 / K4 / K1
// equation pairs with value-level arguments.  The Merge3 solvers call them inside candidate
// enumeration loops (reference cluster x shared element pairs x free candidates), one solve2D per
// candidate.  Here a loop collects its candidates in a Gcs::B200::Merge3Batch and gets them back
// from one kernel launch per kind; the single-call functions below keep the reference's signatures
// and are batches of one.  There is no host implementation of the Newton iteration.
//
// Everything else in the reference file (pose accessors, Procrustes fit, pose scoring) is plain
// host arithmetic and is provided as such.
#pragma once

#include <array>
#include <cstddef>
#include <optional>
#include <span>
#include <utility>
#include <vector>

#include <gcs/b200/leaf_batch.hpp>
#include <gcs/export.hpp>
#include <gcs/math/matrix2d.hpp>
#include <gcs/model/elements.hpp>

#include "solving/bottom_up/plan_pose_types.hpp"

namespace Gcs::Solvers::BottomUp {

struct RigidTransform {
    Eigen::Matrix2d rotation;
    Eigen::Vector2d translation;
};

constexpr double MIN_LINE_LENGTH = 50.0;  // merge3_solver_common.hpp:24
constexpr double EPSILON = 1e-9;          // merge3_solver_common.hpp:25

GCS_API std::optional<LinePose> poseAsLine(const ElementPose& pose);
GCS_API std::optional<PointPose> poseAsPoint(const ElementPose& pose);
GCS_API Eigen::Vector2d lineMidpoint(const LinePose& line);
GCS_API std::optional<Eigen::Vector2d> lineUnitDirection(const LinePose& line);

// Least-squares rotation + translation source -> target (2-D Procrustes through the SVD of the
// 2x2 covariance, reflection removed): merge3_solver_common.cpp:96-160.
GCS_API std::optional<RigidTransform> estimateRigidTransform(
    const std::vector<Eigen::Vector2d>& sourcePoints, const std::vector<Eigen::Vector2d>& targetPoints);
GCS_API ElementPose applyRigidTransform(const ElementPose& pose, const RigidTransform& transform);

GCS_API std::optional<ClusterPose> mergeChildClusterIntoReference(ClusterPose referenceCluster, const ClusterPose& movingCluster);

GCS_API std::optional<Eigen::Vector2d> getPointPosition(const ClusterPose& cluster, ConstraintGraph::NodeIdType elementId);
GCS_API std::optional<Eigen::Vector2d> getPointCanvasPosition(const ConstraintGraph& graph, ConstraintGraph::NodeIdType elementId);
GCS_API std::optional<LinePose> getLinePosition(const ClusterPose& cluster, ConstraintGraph::NodeIdType elementId);
GCS_API std::optional<LinePose> getLineCanvasPose(const ConstraintGraph& graph, ConstraintGraph::NodeIdType elementId);

GCS_API bool isPointElement(const ConstraintGraph& graph, ConstraintGraph::NodeIdType elementId);
GCS_API bool isLineElement(const ConstraintGraph& graph, ConstraintGraph::NodeIdType elementId);

GCS_API std::vector<ConstraintGraph::NodeIdType> clusterIntersectionByType(
    const ConstraintGraph& graph, const ClusterPose& first, const ClusterPose& second, bool selectPoints);

GCS_API std::optional<ClusterPose> transformClusterByTwoPointAnchors(const ClusterPose& movingCluster,
    ConstraintGraph::NodeIdType fixedPoint, ConstraintGraph::NodeIdType freePoint, const Eigen::Vector2d& fixedPointGlobal,
    const Eigen::Vector2d& freePointGlobal);

GCS_API std::optional<ClusterPose> transformClusterByAnchors(
    const ClusterPose& movingCluster, std::span<const std::pair<ConstraintGraph::NodeIdType, ElementPose>> anchors);

GCS_API double scoreMergedPose(const ConstraintGraph& sourceGraph, const ClusterPose& mergedPose);

GCS_API double safeCanvasLineLength(const Line& line);
GCS_API double lineLength(const LinePose& line);
GCS_API double pointToLineDistanceAbs(const Eigen::Vector2d& point, const LinePose& line);

// ---- the numeric helpers: one solve2D-equivalent each, on the device (batch of one) ----
// Deliberate difference: a FIXED line shorter than EPSILON makes the reference substitute
// MIN_LINE_LENGTH for the length inside the point-to-line residual while the direction stays
// (near) zero - a rank-deficient solve whose result is arbitrary; these throw std::domain_error
// there instead (the kernel takes the length from the endpoints).
GCS_API std::optional<LinePose> solveFreeLineFromFixedPoints(const Eigen::Vector2d& fixedPointA,
    const Eigen::Vector2d& fixedPointB, double distanceA, double distanceB, const Eigen::Vector2d& canvasPointA,
    const Eigen::Vector2d& canvasPointB, const LinePose& canvasFreeLine);

GCS_API std::optional<Eigen::Vector2d> solveFreePointFromFixedPointAndLine(const Eigen::Vector2d& fixedPoint,
    const LinePose& fixedLine, double distanceToPoint, double distanceToLine, const Eigen::Vector2d& canvasFixedPoint,
    const LinePose& canvasFixedLine, const Eigen::Vector2d& canvasFreePoint);

GCS_API std::optional<Eigen::Vector2d> solveFreePointFromFixedLines(const LinePose& fixedLineA, const LinePose& fixedLineB,
    double distanceToLineA, double distanceToLineB, const LinePose& canvasLineA, const LinePose& canvasLineB,
    const Eigen::Vector2d& canvasFreePoint);

// The numeric step Merge3PppSolver::solve performs per candidate (merge3_ppp_solver.cpp:135-153):
// two point-to-point distances from the default guesses + pickByTriangleOrientation.
GCS_API Eigen::Vector2d solveFreePointFromFixedPoints(const Eigen::Vector2d& fixedPointA, const Eigen::Vector2d& fixedPointB,
    double distanceA, double distanceB, const Eigen::Vector2d& canvasPointA, const Eigen::Vector2d& canvasPointB,
    const Eigen::Vector2d& canvasFreePoint);

}  // namespace Gcs::Solvers::BottomUp

namespace Gcs::B200 {

// Collects the candidate sub-problems of a Merge3 enumeration and solves them with one kernel
// launch per equation-pair kind.  add*() packs one candidate (canvas-side signs and flags are host
// work, as for the top-down leaves) and returns its handle; cases the reference answers without
// any numerics (std::nullopt) are recorded as such and never reach the device.
class GCS_API Merge3Batch {
public:
    using Handle = std::size_t;

    Merge3Batch();

    Handle addFreePointFromFixedPoints(const Eigen::Vector2d& fixedPointA, const Eigen::Vector2d& fixedPointB, double distanceA,
        double distanceB, const Eigen::Vector2d& canvasPointA, const Eigen::Vector2d& canvasPointB,
        const Eigen::Vector2d& canvasFreePoint);
    Handle addFreeLineFromFixedPoints(const Eigen::Vector2d& fixedPointA, const Eigen::Vector2d& fixedPointB, double distanceA,
        double distanceB, const Eigen::Vector2d& canvasPointA, const Eigen::Vector2d& canvasPointB,
        const Solvers::BottomUp::LinePose& canvasFreeLine);
    Handle addFreePointFromFixedPointAndLine(const Eigen::Vector2d& fixedPoint, const Solvers::BottomUp::LinePose& fixedLine,
        double distanceToPoint, double distanceToLine, const Eigen::Vector2d& canvasFixedPoint,
        const Solvers::BottomUp::LinePose& canvasFixedLine, const Eigen::Vector2d& canvasFreePoint);
    Handle addFreePointFromFixedLines(const Solvers::BottomUp::LinePose& fixedLineA, const Solvers::BottomUp::LinePose& fixedLineB,
        double distanceToLineA, double distanceToLineB, const Solvers::BottomUp::LinePose& canvasLineA,
        const Solvers::BottomUp::LinePose& canvasLineB, const Eigen::Vector2d& canvasFreePoint);

    std::size_t size() const { return m_entries.size(); }
    std::size_t launches() const { return m_launches; }
    // rows of one kind as the kernel receives them (inspection / tests)
    const KindBatch& rows(int kind) const { return m_rows[static_cast<std::size_t>(kind)]; }
    int kindOf(Handle h) const { return m_entries.at(h).kind; }        // 0: answered without numerics
    std::size_t rowOf(Handle h) const { return m_entries.at(h).row; }

    // One launch per kind that has rows.  Throws std::runtime_error when the CUDA library cannot
    // run: there is no CPU fallback.
    void solve(int device = 0);

    // Results (after solve()); std::nullopt where the reference returns std::nullopt.
    std::optional<Eigen::Vector2d> point(Handle h) const;
    std::optional<Solvers::BottomUp::LinePose> line(Handle h) const;

private:
    struct Entry {
        int kind = 0;
        std::size_t row = 0;
    };
    Handle push(int kind, const PackedLeaf& row);
    std::vector<Entry> m_entries;
    mutable std::array<KindBatch, GCS_KIND_COUNT + 1> m_rows;  // descriptor() / out() are non-const views
    std::size_t m_launches = 0;
    bool m_solved = false;
};

}  // namespace Gcs::B200
