// Bottom-up Merge3 numeric helpers over the batched CUDA path (reference:
// src/constraint_solver/src/solving/bottom_up/merge3_solver_common.cpp; the point-from-two-points
// step: merge3_ppp_solver.cpp:135-153).  See the header for the split: Newton solves go to the
// device through Gcs::B200::Merge3Batch, the rest is host arithmetic in the reference's
// evaluation order (the tests compare with the reference build bit for bit).
#include "solving/bottom_up/merge3_solver_common.hpp"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <limits>
#include <stdexcept>
#include <string>

#include "gcs_b200.h"
#include <gcs/math/svd2x2.hpp>
#include "solving/solvers/heuristics.hpp"

using Eigen::Matrix2d;
using Eigen::Vector2d;

namespace Gcs::Solvers::BottomUp {

namespace {

// A line contributes its midpoint and midpoint + unit direction as a pair of anchor points
bool pushLineAnchors(const LinePose& source, const LinePose& target, std::vector<Vector2d>& src, std::vector<Vector2d>& dst)
{
    const auto sd = lineUnitDirection(source), td = lineUnitDirection(target);
    if (!sd || !td) return false;
    const Vector2d sc = lineMidpoint(source), tc = lineMidpoint(target);
    src.push_back(sc), dst.push_back(tc);
    src.push_back(sc + *sd), dst.push_back(tc + *td);
    return true;
}

ClusterPose transformed(const ClusterPose& cluster, const RigidTransform& t)
{
    ClusterPose out;
    out.reserve(cluster.size());
    for (const auto& [id, pose] : cluster) out.emplace(id, applyRigidTransform(pose, t));
    return out;
}

}  // namespace

std::optional<LinePose> poseAsLine(const ElementPose& pose)
{
    if (const auto* l = std::get_if<LinePose>(&pose)) return *l;
    return std::nullopt;
}

std::optional<PointPose> poseAsPoint(const ElementPose& pose)
{
    if (const auto* p = std::get_if<PointPose>(&pose)) return *p;
    return std::nullopt;
}

Vector2d lineMidpoint(const LinePose& line) { return (line.p1 + line.p2) / 2.0; }

std::optional<Vector2d> lineUnitDirection(const LinePose& line)
{
    const Vector2d d = line.p2 - line.p1;
    const double len = d.norm();
    if (len < EPSILON) return std::nullopt;
    return d / len;
}

std::optional<RigidTransform> estimateRigidTransform(const std::vector<Vector2d>& sourcePoints, const std::vector<Vector2d>& targetPoints)
{
    const std::size_t n = sourcePoints.size();
    if (n != targetPoints.size() || n == 0) return std::nullopt;
    if (n == 1) return RigidTransform { Matrix2d::Identity(), targetPoints[0] - sourcePoints[0] };

    Vector2d sc = Vector2d::Zero(), tc = Vector2d::Zero();
    for (std::size_t i = 0; i < n; ++i) sc += sourcePoints[i], tc += targetPoints[i];
    const double count = static_cast<double>(n);
    sc = sc / count, tc = tc / count;

    Matrix2d cov = Matrix2d::Zero();
    for (std::size_t i = 0; i < n; ++i) {
        const Vector2d s = sourcePoints[i] - sc, t = targetPoints[i] - tc;
        cov(0, 0) += s.x() * t.x(), cov(0, 1) += s.x() * t.y();
        cov(1, 0) += s.y() * t.x(), cov(1, 1) += s.y() * t.y();
    }
    Matrix2d u, v;
    Gcs::Math::jacobiSvd2x2(cov, u, v);
    Matrix2d rot = v * u.transpose();
    if (rot.determinant() < 0.0) {  // a reflection: flip the second right singular vector
        v(0, 1) *= -1.0, v(1, 1) *= -1.0;
        rot = v * u.transpose();
    }
    return RigidTransform { rot, tc - rot * sc };
}

ElementPose applyRigidTransform(const ElementPose& pose, const RigidTransform& t)
{
    if (const auto* p = std::get_if<PointPose>(&pose)) return PointPose { t.rotation * p->position + t.translation };
    const auto& l = std::get<LinePose>(pose);
    return LinePose { t.rotation * l.p1 + t.translation, t.rotation * l.p2 + t.translation };
}

std::optional<ClusterPose> mergeChildClusterIntoReference(ClusterPose referenceCluster, const ClusterPose& movingCluster)
{
    std::vector<Vector2d> src, dst;
    for (const auto& [id, moving] : movingCluster) {
        const auto ref = referenceCluster.find(id);
        if (ref == referenceCluster.end()) continue;
        if (const auto* mp = std::get_if<PointPose>(&moving)) {
            const auto* rp = std::get_if<PointPose>(&ref->second);
            if (!rp) return std::nullopt;
            src.push_back(mp->position), dst.push_back(rp->position);
            continue;
        }
        const auto* rl = std::get_if<LinePose>(&ref->second);
        if (!rl || !pushLineAnchors(std::get<LinePose>(moving), *rl, src, dst)) return std::nullopt;
    }
    const auto t = estimateRigidTransform(src, dst);
    if (!t) return std::nullopt;
    for (const auto& [id, moving] : movingCluster)
        if (!referenceCluster.contains(id)) referenceCluster.emplace(id, applyRigidTransform(moving, *t));
    return referenceCluster;
}

std::optional<Vector2d> getPointPosition(const ClusterPose& cluster, ConstraintGraph::NodeIdType id)
{
    const auto it = cluster.find(id);
    if (it == cluster.end()) return std::nullopt;
    if (const auto* p = std::get_if<PointPose>(&it->second)) return p->position;
    return std::nullopt;
}

std::optional<Vector2d> getPointCanvasPosition(const ConstraintGraph& graph, ConstraintGraph::NodeIdType id)
{
    const auto e = graph.getElement(id);
    if (!e || !e->isElementType<Point>()) return std::nullopt;
    return e->getElement<Point>().canvasPosition;
}

std::optional<LinePose> getLinePosition(const ClusterPose& cluster, ConstraintGraph::NodeIdType id)
{
    const auto it = cluster.find(id);
    if (it == cluster.end()) return std::nullopt;
    return poseAsLine(it->second);
}

std::optional<LinePose> getLineCanvasPose(const ConstraintGraph& graph, ConstraintGraph::NodeIdType id)
{
    const auto e = graph.getElement(id);
    if (!e || !e->isElementType<Line>()) return std::nullopt;
    const auto& l = e->getElement<Line>();
    return LinePose { l.canvasP1, l.canvasP2 };
}

bool isPointElement(const ConstraintGraph& graph, ConstraintGraph::NodeIdType id)
{
    const auto e = graph.getElement(id);
    return e && e->isElementType<Point>();
}

bool isLineElement(const ConstraintGraph& graph, ConstraintGraph::NodeIdType id)
{
    const auto e = graph.getElement(id);
    return e && e->isElementType<Line>();
}

std::vector<ConstraintGraph::NodeIdType> clusterIntersectionByType(
    const ConstraintGraph& graph, const ClusterPose& first, const ClusterPose& second, bool selectPoints)
{
    std::vector<ConstraintGraph::NodeIdType> shared;
    for (const auto& entry : first) {
        const auto id = entry.first;
        if (!second.contains(id)) continue;
        if (selectPoints ? isPointElement(graph, id) : isLineElement(graph, id)) shared.push_back(id);
    }
    std::sort(shared.begin(), shared.end());
    shared.erase(std::unique(shared.begin(), shared.end()), shared.end());
    return shared;
}

std::optional<ClusterPose> transformClusterByTwoPointAnchors(const ClusterPose& movingCluster, ConstraintGraph::NodeIdType fixedPoint,
    ConstraintGraph::NodeIdType freePoint, const Vector2d& fixedPointGlobal, const Vector2d& freePointGlobal)
{
    const auto fixedLocal = getPointPosition(movingCluster, fixedPoint);
    const auto freeLocal = getPointPosition(movingCluster, freePoint);
    if (!fixedLocal || !freeLocal) return std::nullopt;
    const auto t = estimateRigidTransform({ *fixedLocal, *freeLocal }, { fixedPointGlobal, freePointGlobal });
    if (!t) return std::nullopt;
    return transformed(movingCluster, *t);
}

std::optional<ClusterPose> transformClusterByAnchors(
    const ClusterPose& movingCluster, std::span<const std::pair<ConstraintGraph::NodeIdType, ElementPose>> anchors)
{
    std::vector<Vector2d> src, dst;
    for (const auto& [id, target] : anchors) {
        const auto it = movingCluster.find(id);
        if (it == movingCluster.end()) return std::nullopt;
        if (const auto* tp = std::get_if<PointPose>(&target)) {
            const auto* sp = std::get_if<PointPose>(&it->second);
            if (!sp) return std::nullopt;
            src.push_back(sp->position), dst.push_back(tp->position);
            continue;
        }
        const auto* sl = std::get_if<LinePose>(&it->second);
        if (!sl || !pushLineAnchors(*sl, std::get<LinePose>(target), src, dst)) return std::nullopt;
    }
    const auto t = estimateRigidTransform(src, dst);
    if (!t) return std::nullopt;
    return transformed(movingCluster, *t);
}

double scoreMergedPose(const ConstraintGraph& sourceGraph, const ClusterPose& mergedPose)
{
    // squared canvas distance of every point / line midpoint, + 100 (1 - |cos|) of the angle
    // between solved and canvas direction of every line: merge3_solver_common.cpp:411-456
    double score = 0.0;
    std::size_t terms = 0;
    for (const auto& [id, pose] : mergedPose) {
        if (const auto* p = std::get_if<PointPose>(&pose)) {
            const auto canvas = getPointCanvasPosition(sourceGraph, id);
            if (!canvas) continue;
            score += (p->position - *canvas).squaredNorm();
            ++terms;
            continue;
        }
        const auto& line = std::get<LinePose>(pose);
        const auto canvas = getLineCanvasPose(sourceGraph, id);
        if (!canvas) continue;
        score += (lineMidpoint(line) - lineMidpoint(*canvas)).squaredNorm();
        const auto sd = lineUnitDirection(line), cd = lineUnitDirection(*canvas);
        if (sd && cd) score += (1.0 - std::abs(sd->dot(*cd))) * 100.0;
        ++terms;
    }
    return terms == 0 ? std::numeric_limits<double>::infinity() : score;
}

double safeCanvasLineLength(const Line& line)
{
    const double len = (line.canvasP2 - line.canvasP1).norm();
    return len < EPSILON ? MIN_LINE_LENGTH : len;
}

double lineLength(const LinePose& line)
{
    const double len = (line.p2 - line.p1).norm();
    return len < EPSILON ? MIN_LINE_LENGTH : len;
}

double pointToLineDistanceAbs(const Vector2d& point, const LinePose& line)
{
    return std::abs(Solvers::signedDistanceToLine(point, line.p1, line.p2));
}

// ---- single-call forms of the numeric helpers: batches of one ----
std::optional<LinePose> solveFreeLineFromFixedPoints(const Vector2d& fixedPointA, const Vector2d& fixedPointB, double distanceA,
    double distanceB, const Vector2d& canvasPointA, const Vector2d& canvasPointB, const LinePose& canvasFreeLine)
{
    B200::Merge3Batch batch;
    const auto h = batch.addFreeLineFromFixedPoints(fixedPointA, fixedPointB, distanceA, distanceB, canvasPointA, canvasPointB, canvasFreeLine);
    batch.solve();
    return batch.line(h);
}

std::optional<Vector2d> solveFreePointFromFixedPointAndLine(const Vector2d& fixedPoint, const LinePose& fixedLine, double distanceToPoint,
    double distanceToLine, const Vector2d& canvasFixedPoint, const LinePose& canvasFixedLine, const Vector2d& canvasFreePoint)
{
    B200::Merge3Batch batch;
    const auto h = batch.addFreePointFromFixedPointAndLine(
        fixedPoint, fixedLine, distanceToPoint, distanceToLine, canvasFixedPoint, canvasFixedLine, canvasFreePoint);
    batch.solve();
    return batch.point(h);
}

std::optional<Vector2d> solveFreePointFromFixedLines(const LinePose& fixedLineA, const LinePose& fixedLineB, double distanceToLineA,
    double distanceToLineB, const LinePose& canvasLineA, const LinePose& canvasLineB, const Vector2d& canvasFreePoint)
{
    B200::Merge3Batch batch;
    const auto h = batch.addFreePointFromFixedLines(
        fixedLineA, fixedLineB, distanceToLineA, distanceToLineB, canvasLineA, canvasLineB, canvasFreePoint);
    batch.solve();
    return batch.point(h);
}

Vector2d solveFreePointFromFixedPoints(const Vector2d& fixedPointA, const Vector2d& fixedPointB, double distanceA, double distanceB,
    const Vector2d& canvasPointA, const Vector2d& canvasPointB, const Vector2d& canvasFreePoint)
{
    B200::Merge3Batch batch;
    const auto h = batch.addFreePointFromFixedPoints(fixedPointA, fixedPointB, distanceA, distanceB, canvasPointA, canvasPointB, canvasFreePoint);
    batch.solve();
    return batch.point(h).value();
}

}  // namespace Gcs::Solvers::BottomUp

namespace Gcs::B200 {

namespace Bu = Solvers::BottomUp;

namespace {

int sign3(double x) { return (x > 0) - (x < 0); }  // heuristics.hpp:54
double signOf(double x) { return (x > 0.0) ? 1.0 : -1.0; }  // signForDistance, merge3_solver_common.cpp:21-24 (zero -> -1)

void requireFixedLine(const Bu::LinePose& l, const char* who)
{
    if ((l.p2 - l.p1).norm() < Bu::EPSILON)
        throw std::domain_error(std::string(who)
            + ": fixed line shorter than 1e-9 (the reference substitutes MIN_LINE_LENGTH inside the residual and "
              "solves a rank-deficient system; not reproduced)");
}

}  // namespace

Merge3Batch::Merge3Batch()
{
    for (int k = 1; k <= GCS_KIND_COUNT; ++k) m_rows[static_cast<std::size_t>(k)] = KindBatch(k);
}

Merge3Batch::Handle Merge3Batch::push(int kind, const PackedLeaf& row)
{
    m_solved = false;
    Entry e;
    e.kind = kind;
    if (kind != 0) {
        auto& b = m_rows[static_cast<std::size_t>(kind)];
        e.row = b.size();
        b.push(row);
    }
    m_entries.push_back(e);
    return m_entries.size() - 1;
}

// merge3_ppp_solver.cpp:135-153
Merge3Batch::Handle Merge3Batch::addFreePointFromFixedPoints(const Vector2d& a, const Vector2d& b, double distanceA, double distanceB,
    const Vector2d& canvasA, const Vector2d& canvasB, const Vector2d& canvasFree)
{
    PackedLeaf r;
    r.kind = GCS_KIND_PP;
    r.in[0] = a.x(), r.in[1] = a.y(), r.in[2] = distanceA;
    r.in[3] = b.x(), r.in[4] = b.y(), r.in[5] = distanceB;
    r.code = GCS_MAKE_CODE(sign3(Solvers::triangleOrientation(canvasA, canvasB, canvasFree)), 0, 0);
    return push(GCS_KIND_PP, r);
}

// merge3_solver_common.cpp:480-531
Merge3Batch::Handle Merge3Batch::addFreeLineFromFixedPoints(const Vector2d& a, const Vector2d& b, double distanceA, double distanceB,
    const Vector2d& canvasA, const Vector2d& canvasB, const Bu::LinePose& canvasFree)
{
    const double sa = Solvers::signedDistanceToLine(canvasA, canvasFree.p1, canvasFree.p2);
    const double sb = Solvers::signedDistanceToLine(canvasB, canvasFree.p1, canvasFree.p2);
    Vector2d dir = canvasFree.p2 - canvasFree.p1;
    if (dir.norm() < Bu::EPSILON) dir = Vector2d { 1.0, 0.0 };
    PackedLeaf r;
    r.kind = GCS_KIND_SDD;
    r.in[0] = a.x(), r.in[1] = a.y(), r.in[2] = b.x(), r.in[3] = b.y();
    r.in[4] = signOf(sa) * distanceA;
    r.in[5] = signOf(sb) * distanceB;
    r.in[6] = -dir.y() / dir.norm(), r.in[7] = dir.x() / dir.norm();  // canvas unit normal: guess 0, guess 1 = its negation
    r.in[8] = Bu::lineLength(canvasFree);
    r.code = GCS_MAKE_CODE(sign3(sa), sign3(sb), 0);
    return push(GCS_KIND_SDD, r);
}

// merge3_solver_common.cpp:533-562
Merge3Batch::Handle Merge3Batch::addFreePointFromFixedPointAndLine(const Vector2d& fixedPoint, const Bu::LinePose& fixedLine,
    double distanceToPoint, double distanceToLine, const Vector2d& canvasFixedPoint, const Bu::LinePose& canvasFixedLine,
    const Vector2d& canvasFree)
{
    requireFixedLine(fixedLine, "solveFreePointFromFixedPointAndLine");
    const double cs = Solvers::signedDistanceToLine(canvasFree, canvasFixedLine.p1, canvasFixedLine.p2);
    PackedLeaf r;
    r.kind = GCS_KIND_PPL;
    r.in[0] = fixedPoint.x(), r.in[1] = fixedPoint.y(), r.in[2] = distanceToPoint;
    r.in[3] = fixedLine.p1.x(), r.in[4] = fixedLine.p1.y(), r.in[5] = fixedLine.p2.x(), r.in[6] = fixedLine.p2.y();
    r.in[7] = signOf(cs) * distanceToLine;
    r.in[8] = canvasFree.x(), r.in[9] = canvasFree.y();
    const Vector2d canvasFoot = Solvers::perpendicularFoot(canvasFixedPoint, canvasFixedLine.p1, canvasFixedLine.p2);
    const double ori = Solvers::triangleOrientation(canvasFixedPoint, canvasFoot, canvasFree);
    r.code = GCS_MAKE_CODE(sign3(ori), 0, (std::abs(ori) < GCS_COLLINEAR_EPSILON) ? GCS_CODE_COLLINEAR : 0);
    return push(GCS_KIND_PPL, r);
}

// merge3_solver_common.cpp:564-608
Merge3Batch::Handle Merge3Batch::addFreePointFromFixedLines(const Bu::LinePose& lineA, const Bu::LinePose& lineB, double distanceToLineA,
    double distanceToLineB, const Bu::LinePose& canvasA, const Bu::LinePose& canvasB, const Vector2d& canvasFree)
{
    const auto solverX = Solvers::lineLineIntersection(lineA.p1, lineA.p2, lineB.p1, lineB.p2);
    const auto canvasX = Solvers::lineLineIntersection(canvasA.p1, canvasA.p2, canvasB.p1, canvasB.p2);
    const bool framed = solverX.has_value() && canvasX.has_value();
    // with both intersections the reference needs the unit directions of line A in both spaces and
    // answers nullopt without looking at the candidates when either is degenerate (:588-592)
    if (framed && (!Bu::lineUnitDirection(lineA) || !Bu::lineUnitDirection(canvasA))) return push(0, PackedLeaf {});
    requireFixedLine(lineA, "solveFreePointFromFixedLines");
    requireFixedLine(lineB, "solveFreePointFromFixedLines");
    const double ca = Solvers::signedDistanceToLine(canvasFree, canvasA.p1, canvasA.p2);
    const double cb = Solvers::signedDistanceToLine(canvasFree, canvasB.p1, canvasB.p2);
    PackedLeaf r;
    r.kind = GCS_KIND_PLL;
    r.in[0] = lineA.p1.x(), r.in[1] = lineA.p1.y(), r.in[2] = lineA.p2.x(), r.in[3] = lineA.p2.y(), r.in[4] = signOf(ca) * distanceToLineA;
    r.in[5] = lineB.p1.x(), r.in[6] = lineB.p1.y(), r.in[7] = lineB.p2.x(), r.in[8] = lineB.p2.y(), r.in[9] = signOf(cb) * distanceToLineB;
    r.in[10] = canvasFree.x(), r.in[11] = canvasFree.y();
    if (!canvasX) {
        r.code = GCS_MAKE_CODE(0, 0, GCS_CODE_CANVAS_PARALLEL);  // nearest to the canvas point (:604-607)
    } else {
        const Vector2d canvasRef = *canvasX + Bu::lineUnitDirection(canvasA).value_or(Vector2d { 0.0, 0.0 });
        const double ori = Solvers::triangleOrientation(*canvasX, canvasRef, canvasFree);
        r.code = GCS_MAKE_CODE(sign3(ori), 0, (std::abs(ori) < GCS_COLLINEAR_EPSILON) ? GCS_CODE_COLLINEAR : 0);
    }
    return push(GCS_KIND_PLL, r);
}

void Merge3Batch::solve(int device)
{
    if (m_solved) return;
    for (int k = 1; k <= GCS_KIND_COUNT; ++k) {
        auto& b = m_rows[static_cast<std::size_t>(k)];
        if (b.size() == 0) continue;
        gcs_b200_batch d = b.descriptor();
        const int rc = gcs_b200_solve_host(&d, device);
        if (rc != GCS_OK)
            throw std::runtime_error(std::string("Merge3Batch: gcs_b200_solve_host failed (") + std::to_string(rc) + "): "
                + gcs_b200_last_error() + " - the Merge3 numeric helpers run on the CUDA path only");
        ++m_launches;
    }
    m_solved = true;
}

std::optional<Vector2d> Merge3Batch::point(Handle h) const
{
    const Entry& e = m_entries.at(h);
    if (e.kind == 0) return std::nullopt;
    if (!m_solved) throw std::logic_error("Merge3Batch::point before solve()");
    auto& b = m_rows[static_cast<std::size_t>(e.kind)];
    if (gcs_b200_kind_out_cols(e.kind) != 2) throw std::logic_error("Merge3Batch::point on a line-valued entry");
    return Vector2d { b.out(0)[e.row], b.out(1)[e.row] };
}

std::optional<Bu::LinePose> Merge3Batch::line(Handle h) const
{
    const Entry& e = m_entries.at(h);
    if (e.kind == 0) return std::nullopt;
    if (!m_solved) throw std::logic_error("Merge3Batch::line before solve()");
    auto& b = m_rows[static_cast<std::size_t>(e.kind)];
    if (gcs_b200_kind_out_cols(e.kind) != 4) throw std::logic_error("Merge3Batch::line on a point-valued entry");
    return Bu::LinePose { Vector2d { b.out(0)[e.row], b.out(1)[e.row] }, Vector2d { b.out(2)[e.row], b.out(3)[e.row] } };
}

}  // namespace Gcs::B200
