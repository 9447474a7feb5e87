// ref_merge3_driver.cpp — C entry points over the REFERENCE'S OWN bottom-up Merge3 numeric helpers
// (src/constraint_solver/src/solving/bottom_up/merge3_solver_common.cpp, compiled where it lies
// under /root/reference against the stand-in headers of oracle/ref_shim).  TEST INFRASTRUCTURE ONLY.
//
// Reference code behind these entry points, unmodified:
//   solveFreeLineFromFixedPoints          merge3_solver_common.cpp:480-531
//   solveFreePointFromFixedPointAndLine   merge3_solver_common.cpp:533-562
//   solveFreePointFromFixedLines          merge3_solver_common.cpp:564-608
//   estimateRigidTransform / applyRigidTransform   :96-160, :162-178
//   scoreMergedPose                       :411-456
//   Merge3PppSolver::solve (the whole enumeration loop)   merge3_ppp_solver.cpp:18-214
//   Merge3PllSolver / Merge3LppSolver / Merge3LlpSolver::solve   merge3_pll_solver.cpp:15-189, merge3_lpp_solver.cpp:15-208,
//                                                                  merge3_llp_solver.cpp:15-190
//   detectUnsolvableMerge3Lll, Merge3FallbackSolver::solve        merge3_fallback_solver.cpp:13-78
// The point-from-two-points step is inlined in Merge3PppSolver::solve (merge3_ppp_solver.cpp:135-153);
// gcs_ref_m3_point_pp repeats those three calls on the reference's own functions.
// Third-party arithmetic (Eigen's JacobiSVD, ColPivHouseholderQR, autodiff) is the stand-ins'.
#include <cmath>
#include <cstdint>
#include <limits>
#include <memory>
#include <unordered_map>
#include <vector>

#include "solving/bottom_up/merge3_fallback_solver.hpp"
#include "solving/bottom_up/merge3_llp_solver.hpp"
#include "solving/bottom_up/merge3_lpp_solver.hpp"
#include "solving/bottom_up/merge3_pll_solver.hpp"
#include "solving/bottom_up/merge3_ppp_solver.hpp"
#include "solving/bottom_up/merge3_solver_common.hpp"
#include "solving/equations/equation_primitives.hpp"
#include "solving/equations/newton_raphson.hpp"
#include "solving/solvers/heuristics.hpp"
#include <gcs/model/elements.hpp>
#include <gcs/model/gcs_data_structures.hpp>

using Eigen::Vector2d;
namespace Bu = Gcs::Solvers::BottomUp;

#define REF_API extern "C" __attribute__((visibility("default")))

namespace {
Vector2d v2(const double* p) { return Vector2d(p[0], p[1]); }
Bu::LinePose ln(const double* p) { return Bu::LinePose { .p1 = v2(p), .p2 = v2(p + 2) }; }
const double kNaN = std::numeric_limits<double>::quiet_NaN();
}  // namespace

// rows of 12: fixedA(2) fixedB(2) distA distB canvasA(2) canvasB(2) canvasFree(2) -> point (2)
REF_API void gcs_ref_m3_point_pp(int64_t n, const double* rows, double* out)
{
    for (int64_t i = 0; i < n; ++i) {
        const double* r = rows + 12 * i;
        auto f = Gcs::Equations::pointToPointDistance(r[0], r[1], r[4]);
        auto g = Gcs::Equations::pointToPointDistance(r[2], r[3], r[5]);
        const auto cand = Gcs::Equations::solve2D(f, g);
        const Vector2d pick = Gcs::Solvers::pickByTriangleOrientation(v2(r + 6), v2(r + 8), v2(r + 10), v2(r), v2(r + 2), cand[0], cand[1]);
        out[2 * i] = pick.x(), out[2 * i + 1] = pick.y();
    }
}

// rows of 14: fixedA(2) fixedB(2) distA distB canvasA(2) canvasB(2) canvasFreeLine(4) -> line (4)
REF_API void gcs_ref_m3_free_line(int64_t n, const double* rows, double* out, uint8_t* ok)
{
    for (int64_t i = 0; i < n; ++i) {
        const double* r = rows + 14 * i;
        const auto res = Bu::solveFreeLineFromFixedPoints(v2(r), v2(r + 2), r[4], r[5], v2(r + 6), v2(r + 8), ln(r + 10));
        ok[i] = res.has_value();
        for (int c = 0; c < 4; ++c) out[4 * i + c] = kNaN;
        if (res) out[4 * i] = res->p1.x(), out[4 * i + 1] = res->p1.y(), out[4 * i + 2] = res->p2.x(), out[4 * i + 3] = res->p2.y();
    }
}

// rows of 16: fixedPoint(2) fixedLine(4) distPoint distLine canvasFixedPoint(2) canvasFixedLine(4) canvasFree(2)
REF_API void gcs_ref_m3_point_pl(int64_t n, const double* rows, double* out, uint8_t* ok)
{
    for (int64_t i = 0; i < n; ++i) {
        const double* r = rows + 16 * i;
        const auto res = Bu::solveFreePointFromFixedPointAndLine(v2(r), ln(r + 2), r[6], r[7], v2(r + 8), ln(r + 10), v2(r + 14));
        ok[i] = res.has_value();
        out[2 * i] = res ? res->x() : kNaN, out[2 * i + 1] = res ? res->y() : kNaN;
    }
}

// rows of 20: fixedLineA(4) fixedLineB(4) distA distB canvasLineA(4) canvasLineB(4) canvasFree(2)
REF_API void gcs_ref_m3_point_ll(int64_t n, const double* rows, double* out, uint8_t* ok)
{
    for (int64_t i = 0; i < n; ++i) {
        const double* r = rows + 20 * i;
        const auto res = Bu::solveFreePointFromFixedLines(ln(r), ln(r + 4), r[8], r[9], ln(r + 10), ln(r + 14), v2(r + 18));
        ok[i] = res.has_value();
        out[2 * i] = res ? res->x() : kNaN, out[2 * i + 1] = res ? res->y() : kNaN;
    }
}

// estimateRigidTransform on npts point pairs (src, dst: x0 y0 x1 y1 ...); out6 = R00 R01 R10 R11 tx ty
REF_API int gcs_ref_m3_rigid_transform(int npts, const double* src, const double* dst, double* out6)
{
    std::vector<Vector2d> s, t;
    for (int i = 0; i < npts; ++i) s.push_back(v2(src + 2 * i)), t.push_back(v2(dst + 2 * i));
    const auto tr = Bu::estimateRigidTransform(s, t);
    if (!tr) return 0;
    out6[0] = tr->rotation(0, 0), out6[1] = tr->rotation(0, 1), out6[2] = tr->rotation(1, 0), out6[3] = tr->rotation(1, 1);
    out6[4] = tr->translation.x(), out6[5] = tr->translation.y();
    return 1;
}

// scoreMergedPose: elements i = 0..n_el-1 of a graph (type 0 point: canvas/pose x,y; type 1 line:
// x1,y1,x2,y2); in_pose[i] != 0 puts element i into the merged pose, inserted in index order.
REF_API double gcs_ref_m3_score(int n_el, const int32_t* type, const double* canvas4, const double* pose4, const uint8_t* in_pose)
{
    Gcs::ConstraintGraph g;
    Bu::ClusterPose merged;
    for (int i = 0; i < n_el; ++i) {
        const double* c = canvas4 + 4 * i;
        const double* p = pose4 + 4 * i;
        const auto node = g.getGraph().addNode();
        if (type[i] == 0)
            g.addElement(node, std::make_shared<Gcs::Element>(Gcs::Point(v2(c))));
        else
            g.addElement(node, std::make_shared<Gcs::Element>(Gcs::Line(v2(c), v2(c + 2))));
        if (!in_pose[i]) continue;
        if (type[i] == 0)
            merged.emplace(node, Bu::PointPose { .position = v2(p) });
        else
            merged.emplace(node, ln(p));
    }
    return Bu::scoreMergedPose(g, merged);
}

// The reference's own Merge3 solver classes on three child clusters of one sketch.
// Elements i = 0..n_el-1 (type 0 point: canvas x,y; 1 line: x1,y1,x2,y2).  Cluster c = 0..2 holds
// counts[c] elements: ids / pose4 concatenated in the order they are inserted into the cluster's
// pose map.  which = 0 Merge3PppSolver, 1 Merge3PllSolver, 2 Merge3LppSolver, 3 Merge3LlpSolver,
// 4 Merge3FallbackSolver, 5 the case order of a merge node (bottom_up_plan_solver.cpp:393-431: the
// first of PPP, PLL, LPP, LLP that returns a pose; else nothing when detectUnsolvableMerge3Lll;
// else the fallback) - *solved_by gets the case that produced the pose (5: none).
// Returns the size of the merged pose (0: none), out_ids ascending, out_pose4 per id.
REF_API int gcs_ref_m3_merge(int which, int n_el, const int32_t* type, const double* canvas4, const int32_t* counts, const int32_t* ids,
    const double* pose4, int32_t* out_ids, double* out_pose4, int32_t* solved_by)
{
    Gcs::ConstraintGraph g;
    std::vector<Gcs::ConstraintGraph::NodeIdType> nodes;
    for (int i = 0; i < n_el; ++i) {
        const double* c = canvas4 + 4 * i;
        nodes.push_back(g.getGraph().addNode());
        if (type[i] == 0)
            g.addElement(nodes.back(), std::make_shared<Gcs::Element>(Gcs::Point(v2(c))));
        else
            g.addElement(nodes.back(), std::make_shared<Gcs::Element>(Gcs::Line(v2(c), v2(c + 2))));
    }
    std::unordered_map<MathUtils::GeneralTreeNodeId, Bu::ClusterPose> poses;
    std::vector<MathUtils::GeneralTreeNodeId> children;
    int at = 0;
    for (int c = 0; c < 3; ++c) {
        Bu::ClusterPose pose;
        for (int k = 0; k < counts[c]; ++k, ++at) {
            const double* p = pose4 + 4 * at;
            if (type[ids[at]] == 0)
                pose.emplace(nodes[ids[at]], Bu::PointPose { .position = v2(p) });
            else
                pose.emplace(nodes[ids[at]], ln(p));
        }
        const MathUtils::GeneralTreeNodeId child { c + 1 };
        children.push_back(child);
        poses.emplace(child, std::move(pose));
    }
    Gcs::PlanNode node { .kind = Gcs::PlanNodeKind::Merge3,
        .info = Gcs::Merge3Info { .output = Gcs::ClusterId { 9 }, .inputs = { Gcs::ClusterId { 1 }, Gcs::ClusterId { 2 }, Gcs::ClusterId { 3 } }, .outputElements = {} } };
    const Bu::Merge3Context context { .sourceGraph = g, .node = node, .children = children, .solvedNodePose = poses };
    std::optional<Bu::ClusterPose> merged;
    int by = which;
    switch (which) {
    case 0: merged = Bu::Merge3PppSolver::solve(context); break;
    case 1: merged = Bu::Merge3PllSolver::solve(context); break;
    case 2: merged = Bu::Merge3LppSolver::solve(context); break;
    case 3: merged = Bu::Merge3LlpSolver::solve(context); break;
    case 4: merged = Bu::Merge3FallbackSolver::solve(context); break;
    default:
        by = 0;
        if ((merged = Bu::Merge3PppSolver::solve(context))) break;
        by = 1;
        if ((merged = Bu::Merge3PllSolver::solve(context))) break;
        by = 2;
        if ((merged = Bu::Merge3LppSolver::solve(context))) break;
        by = 3;
        if ((merged = Bu::Merge3LlpSolver::solve(context))) break;
        by = 5;
        if (Bu::detectUnsolvableMerge3Lll(context)) break;
        merged = Bu::Merge3FallbackSolver::solve(context);
        by = merged ? 4 : 5;
    }
    if (solved_by) *solved_by = by;
    if (!merged) return 0;
    int n = 0;
    for (int i = 0; i < n_el; ++i) {
        const auto it = merged->find(nodes[i]);
        if (it == merged->end()) continue;
        out_ids[n] = i;
        double* o = out_pose4 + 4 * n;
        o[0] = o[1] = o[2] = o[3] = 0.0;
        if (const auto* pp = std::get_if<Bu::PointPose>(&it->second))
            o[0] = pp->position.x(), o[1] = pp->position.y();
        else {
            const auto& l = std::get<Bu::LinePose>(it->second);
            o[0] = l.p1.x(), o[1] = l.p1.y(), o[2] = l.p2.x(), o[3] = l.p2.y();
        }
        ++n;
    }
    return n;
}

REF_API int gcs_ref_m3_ppp_merge(int n_el, const int32_t* type, const double* canvas4, const int32_t* counts, const int32_t* ids,
    const double* pose4, int32_t* out_ids, double* out_pose4)
{
    return gcs_ref_m3_merge(0, n_el, type, canvas4, counts, ids, pose4, out_ids, out_pose4, nullptr);
}
