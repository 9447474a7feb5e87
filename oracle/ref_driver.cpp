// ref_driver.cpp — C entry points over the REFERENCE'S OWN sources, compiled where they lie
// under /root/reference against the stand-in headers in oracle/ref_shim (Eigen, autodiff,
// <flat_map>, spdlog, OGDF bridge are absent in this image).  TEST INFRASTRUCTURE ONLY.
//
// Reference code that runs unmodified behind these entry points:
//   solving/equations/newton_raphson.hpp      (Equations::solve2D, constants, default guesses)
//   solving/equations/equation_primitives.hpp (the five primitives in use)
//   solving/solvers/heuristics.hpp            (every pick* / geometric helper)
//   solving/solvers/{point_point,point_line,line_angle}_solvers.cpp (the 8 matches()/solve())
//   solving/component_solver.hpp              (classifyAndSolve dispatch order)
//   model/{elements,constraints}.cpp, gcs/model/*.hpp, structures/{simple_graph,property_map}.hpp
// What is NOT reference code: the third-party arithmetic inside the stand-ins, and the nine
// trivial ConstraintGraph members in ref_graph_members.cpp (their home TU,
// src/model/gcs_data_structures.cpp, needs GCC >= 14 deducing-this and OGDF).
#include <array>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <memory>
#include <vector>

#include "../include/gcs_b200.h"

#include "solving/component_solver.hpp"
#include "solving/equations/equation_primitives.hpp"
#include "solving/equations/newton_raphson.hpp"
#include "solving/solvers/heuristics.hpp"
#include <gcs/model/constraints.hpp>
#include <gcs/model/elements.hpp>
#include <gcs/model/gcs_data_structures.hpp>
#include <gcs/model/solve_result.hpp>

#ifdef _OPENMP
#include <omp.h>
#endif

using Eigen::Vector2d;
using autodiff::dual;
namespace Eq = Gcs::Equations;
namespace Hs = Gcs::Solvers;

namespace {

// Wraps an equation so that the number of evaluations can be read back: solve2D evaluates each
// equation three times per loop iteration (two derivatives + one value, newton_raphson.hpp:66-77),
// so a run that broke at iteration i made 3*(i+1) calls; 3000 calls = i >= 999.
template <typename F>
struct Counted {
    F f;
    long* calls;
    dual operator()(dual x, dual y) const
    {
        ++*calls;
        return f(x, y);
    }
};

template <typename F, typename G>
void run_solve2d(const F& f, const G& g, const std::array<Vector2d, 2>* guesses, bool count,
    std::array<Vector2d, 2>& cand, int iters[2], int conv[2])
{
    iters[0] = iters[1] = -1;
    conv[0] = conv[1] = -1;
    if (!count) {
        cand = guesses ? Eq::solve2D(f, g, *guesses) : Eq::solve2D(f, g);
        return;
    }
    const std::array<Vector2d, 2> gs = guesses ? *guesses : Eq::DEFAULT_SPATIAL_GUESSES;
    for (int s = 0; s < 2; ++s) {
        long calls = 0;
        Counted<F> cf { f, &calls };
        std::array<Vector2d, 2> twice { gs[s], gs[s] };
        auto r = Eq::solve2D(cf, g, twice);
        cand[s] = r[0];
        const long per_run = calls / 2;  // both runs are identical
        const int evals = (int)(per_run / 3);
        iters[s] = evals - 1;
        conv[s] = 1;
        if (evals >= Eq::MAXIMUM_ITERATIONS) {
            // broke at i = 999 or ran out: decide with the reference's own test on one more step
            iters[s] = Eq::MAXIMUM_ITERATIONS;  // reported as ">= 999"; callers treat 999/1000 alike
            conv[s] = 0;
        }
    }
}

// index of the candidate the reference's heuristic returned; 2 = unobservable (both candidates
// carry the same bits, e.g. both seeds reached the same root)
int which(const Vector2d& chosen, const std::array<Vector2d, 2>& cand)
{
    auto same = [](double a, double b) { return std::memcmp(&a, &b, 8) == 0 || (a != a && b != b); };
    const bool is0 = same(chosen.x(), cand[0].x()) && same(chosen.y(), cand[0].y());
    const bool is1 = same(chosen.x(), cand[1].x()) && same(chosen.y(), cand[1].y());
    if (is0 && is1) return 2;
    return is0 ? 0 : 1;
}

void solve_one(const gcs_b200_batch* b, int64_t i, bool count)
{
    const int64_t n = b->n;
    double in[GCS_MAX_IN_COLS];
    const int nin = gcs_b200_kind_in_cols(b->kind);
    for (int c = 0; c < nin; ++c) in[c] = b->in[c][i];
    const uint8_t code = b->code[i];
    const int sign0 = GCS_CODE_SIGN0(code), sign1 = GCS_CODE_SIGN1(code);
    std::array<Vector2d, 2> cand;
    int iters[2], conv[2], root = 0;
    double out[4] = { NAN, NAN, NAN, NAN };
    std::array<Vector2d, 2> gs;
    const std::array<Vector2d, 2>* gp = nullptr;
    if (b->guesses) {
        for (int s = 0; s < 2; ++s) gs[s] = Vector2d(b->guesses[(s * 2 + 0) * n + i], b->guesses[(s * 2 + 1) * n + i]);
        gp = &gs;
    }
    // a canvas triangle whose orientation has the wanted three-valued sign
    const Vector2d cA(0.0, 0.0), cB(1.0, 0.0), cF(0.0, (double)sign0);

    switch (b->kind) {
    case GCS_KIND_PP: {
        auto f = Eq::pointToPointDistance(in[0], in[1], in[2]);
        auto g = Eq::pointToPointDistance(in[3], in[4], in[5]);
        run_solve2d(f, g, gp, count, cand, iters, conv);
        Vector2d pick = Hs::pickByTriangleOrientation(cA, cB, cF, Vector2d(in[0], in[1]), Vector2d(in[3], in[4]), cand[0], cand[1]);
        root = which(pick, cand);
        out[0] = pick.x(), out[1] = pick.y();
        break;
    }
    case GCS_KIND_SDD: {
        const Vector2d p1(in[0], in[1]), p2(in[2], in[3]);
        const Vector2d delta = p2 - p1;
        auto f = Eq::lineNormalSignedDistanceDiff(delta.x(), delta.y(), in[4], in[5]);
        auto g = Eq::unitNormalConstraint();
        const Vector2d cn(in[6], in[7]);
        if (!gp) {
            gs = { cn, -cn };
            gp = &gs;
        }
        run_solve2d(f, g, gp, count, cand, iters, conv);
        const double off0 = cand[0].dot(p1) - in[4];
        const double off1 = cand[1].dot(p1) - in[4];
        auto [nx, ny, off] = Hs::pickLineBySignedDistances((double)sign0, (double)sign1, cand[0], cand[1], p1, p2, off0, off1);
        root = which(Vector2d(nx, ny), cand);
        out[0] = nx, out[1] = ny, out[2] = off;  // (normal, offset): endpoints need the solver TU
        break;
    }
    case GCS_KIND_PPL: {
        Gcs::Line ln;
        ln.updateElementPosition(Vector2d(in[3], in[4]), Vector2d(in[5], in[6]));
        auto f = Eq::pointToPointDistance(in[0], in[1], in[2]);
        auto g = Eq::pointToLineDistance(ln.p1.x(), ln.p1.y(), ln.p2.x(), ln.p2.y(), in[7], ln.length());
        run_solve2d(f, g, gp, count, cand, iters, conv);
        const Vector2d fixedPt(in[0], in[1]);
        const Vector2d foot = Hs::perpendicularFoot(fixedPt, ln.p1, ln.p2);
        Vector2d a = cA, bb = cB, fr = cF;
        if (code & GCS_CODE_COLLINEAR) {
            fr = Vector2d(in[8], in[9]);
            a = Vector2d(in[8] - 1.0, in[9]);
            bb = Vector2d(in[8] + 1.0, in[9]);
        }
        Vector2d pick = Hs::pickByTriangleOrientationWithFallback(a, bb, fr, fixedPt, foot, cand[0], cand[1]);
        root = which(pick, cand);
        out[0] = pick.x(), out[1] = pick.y();
        break;
    }
    case GCS_KIND_PLL: {
        Gcs::Line l1, l2;
        l1.updateElementPosition(Vector2d(in[0], in[1]), Vector2d(in[2], in[3]));
        l2.updateElementPosition(Vector2d(in[5], in[6]), Vector2d(in[7], in[8]));
        auto f = Eq::pointToLineDistance(l1.p1.x(), l1.p1.y(), l1.p2.x(), l1.p2.y(), in[4], l1.length());
        auto g = Eq::pointToLineDistance(l2.p1.x(), l2.p1.y(), l2.p2.x(), l2.p2.y(), in[9], l2.length());
        run_solve2d(f, g, gp, count, cand, iters, conv);
        auto si = Hs::lineLineIntersection(l1.p1, l1.p2, l2.p1, l2.p2);
        const Vector2d freeCanvas(in[10], in[11]);
        Vector2d pick;
        if (si && !(code & GCS_CODE_CANVAS_PARALLEL)) {
            const Vector2d sref = *si + l1.unitDirection();
            Vector2d a = cA, bb = cB, fr = cF;
            if (code & GCS_CODE_COLLINEAR) {
                fr = freeCanvas;
                a = Vector2d(in[10] - 1.0, in[11]);
                bb = Vector2d(in[10] + 1.0, in[11]);
            }
            pick = Hs::pickByTriangleOrientationWithFallback(a, bb, fr, *si, sref, cand[0], cand[1]);
        } else {
            const double d0 = (cand[0] - freeCanvas).squaredNorm();
            const double d1 = (cand[1] - freeCanvas).squaredNorm();
            pick = (d0 <= d1) ? cand[0] : cand[1];  // point_line_solvers.cpp:676-681
        }
        root = which(pick, cand);
        out[0] = pick.x(), out[1] = pick.y();
        break;
    }
    case GCS_KIND_ANG: {
        const Vector2d fd(in[0], in[1]);
        auto f = Eq::lineNormalAngleConstraint(fd.x(), fd.y(), fd.norm(), in[2]);
        auto g = Eq::unitNormalConstraint();
        const Vector2d cn(in[3], in[4]);
        if (!gp) {
            gs = { cn, -cn };
            gp = &gs;
        }
        run_solve2d(f, g, gp, count, cand, iters, conv);
        const Vector2d cfd(in[5], in[6]);
        const Vector2d cfree((double)sign0 * -cfd.y(), (double)sign0 * cfd.x());
        Vector2d pick = Hs::pickLineNormalByAngleOrientation(cfd, cfree, cand[0], cand[1]);
        root = which(pick, cand);
        const double off = pick.dot(Vector2d(in[7], in[8])) - in[9];
        out[0] = pick.x(), out[1] = pick.y(), out[2] = off;
        break;
    }
    }
    const int nout = gcs_b200_kind_out_cols(b->kind);
    for (int c = 0; c < nout; ++c)
        if (b->out[c]) b->out[c][i] = out[c];
    for (int s = 0; s < 2; ++s) {
        if (b->cand) {
            b->cand[(s * 2 + 0) * n + i] = cand[s].x();
            b->cand[(s * 2 + 1) * n + i] = cand[s].y();
        }
        if (b->iters) b->iters[s * n + i] = (int16_t)iters[s];
        if (b->converged) b->converged[s * n + i] = (uint8_t)conv[s];
    }
    if (b->root_index) b->root_index[i] = (uint8_t)root;
}

}  // namespace

extern "C" {

// Kind tables duplicated here so that this library does not depend on the product's .so.
int gcs_b200_kind_in_cols(int kind)
{
    static const int t[GCS_KIND_COUNT + 1] = { 0, 6, 9, 10, 12, 13 };
    return (kind >= 1 && kind <= GCS_KIND_COUNT) ? t[kind] : 0;
}
int gcs_b200_kind_out_cols(int kind)
{
    static const int t[GCS_KIND_COUNT + 1] = { 0, 2, 4, 2, 2, 4 };
    return (kind >= 1 && kind <= GCS_KIND_COUNT) ? t[kind] : 0;
}

// The reference's solve2D + heuristics on a host batch (2 seeds).  For the line kinds out[] is
// (nx, ny, offset, NaN): the endpoints come out of the solver TUs, see gcs_ref_component_solve.
// count_iters != 0 runs every seed twice through a counting wrapper to recover the number of
// loop iterations the reference executed.
__attribute__((visibility("default"))) int gcs_ref_solve_batch(const gcs_b200_batch* b, int count_iters, int threads)
{
    if (!b || b->kind < 1 || b->kind > GCS_KIND_COUNT || b->n_seeds != 2 || b->mem != GCS_MEM_HOST) return GCS_E_INVALID;
    (void)threads;
#ifdef _OPENMP
    if (threads < 1) threads = omp_get_max_threads();
#pragma omp parallel for schedule(static) num_threads(threads)
#endif
    for (int64_t i = 0; i < b->n; ++i) solve_one(b, i, count_iters != 0);
    return GCS_OK;
}

__attribute__((visibility("default"))) int gcs_ref_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// ---- component level: a 3-element leaf through the reference's classifyAndSolve -------------
typedef struct gcs_ref_element {
    int32_t type;      // 0 = Point, 1 = Line
    int32_t is_set;    // in: already solved (position valid); out: m_isSet after the solve
    double canvas[4];  // point: x,y ; line: x1,y1,x2,y2
    double pos[4];     // solver-space position, same layout
} gcs_ref_element;

typedef struct gcs_ref_edge {
    int32_t a, b;   // element indices
    int32_t type;   // 0 = Distance, 1 = Angle, 2 = virtual edge (no constraint)
    int32_t flip;   // AngleConstraint::flipOrientation
    double value;   // distance, or angle in radians
} gcs_ref_edge;

// returns SolveStatus (0 Success, 1 Unsupported, 2 Failed) or -1 on an exception
__attribute__((visibility("default"))) int gcs_ref_component_solve(
    int n_el, gcs_ref_element* el, int n_edges, const gcs_ref_edge* edges)
{
    try {
        Gcs::ConstraintGraph g;
        std::vector<Gcs::ConstraintGraph::NodeIdType> nodes;
        std::vector<std::shared_ptr<Gcs::Element>> elems;
        for (int i = 0; i < n_el; ++i) {
            auto node = g.getGraph().addNode();
            std::shared_ptr<Gcs::Element> e;
            if (el[i].type == 0) {
                e = std::make_shared<Gcs::Element>(Gcs::Point(Vector2d(el[i].canvas[0], el[i].canvas[1])));
                if (el[i].is_set) e->updateElementPosition(Vector2d(el[i].pos[0], el[i].pos[1]));
            } else {
                e = std::make_shared<Gcs::Element>(Gcs::Line(Vector2d(el[i].canvas[0], el[i].canvas[1]), Vector2d(el[i].canvas[2], el[i].canvas[3])));
                if (el[i].is_set) e->updateElementPosition(Vector2d(el[i].pos[0], el[i].pos[1]), Vector2d(el[i].pos[2], el[i].pos[3]));
            }
            g.addElement(node, e);
            nodes.push_back(node);
            elems.push_back(e);
        }
        for (int k = 0; k < n_edges; ++k) {
            const auto& ed = edges[k];
            if (ed.type == 2) {
                g.addVirtualEdge(nodes[ed.a], nodes[ed.b]);
                continue;
            }
            auto eid = g.getGraph().addEdge(nodes[ed.a], nodes[ed.b]).value();
            std::shared_ptr<Gcs::Constraint> c;
            if (ed.type == 0)
                c = std::make_shared<Gcs::Constraint>(Gcs::DistanceConstraint(ed.value));
            else
                c = std::make_shared<Gcs::Constraint>(Gcs::AngleConstraint(ed.value, ed.flip != 0));
            g.addConstraint(eid, c);
        }
        const Gcs::SolveResult r = Gcs::classifyAndSolve(g);
        for (int i = 0; i < n_el; ++i) {
            el[i].is_set = elems[i]->isElementSet() ? 1 : 0;
            if (el[i].type == 0) {
                const auto& p = elems[i]->getElement<Gcs::Point>();
                el[i].pos[0] = p.position.x(), el[i].pos[1] = p.position.y();
            } else {
                const auto& l = elems[i]->getElement<Gcs::Line>();
                el[i].pos[0] = l.p1.x(), el[i].pos[1] = l.p1.y(), el[i].pos[2] = l.p2.x(), el[i].pos[3] = l.p2.y();
            }
        }
        return (int)r.status;
    } catch (const std::exception& ex) {
        if (std::getenv("GCS_REF_DEBUG")) std::fprintf(stderr, "[gcs_ref] exception: %s\n", ex.what());
        return -1;
    } catch (...) {
        return -1;
    }
}

// ---- leaf list level: the reference's sequential loop over leaves that share elements -------
// (DeficitStreeBasedTopDownStrategy::solveGcs = for_each(leaves, classifyAndSolve),
// stree_top_down_strategy.cpp:41-45).  leaf_elems: 3 element indices per leaf (node order);
// edge_offsets: n_leaves + 1 offsets into edges (a, b = element indices).  status[l] = SolveStatus
// of leaf l, or -1 from the leaf that threw on (the loop stops there, as an exception would).
__attribute__((visibility("default"))) int gcs_ref_leaves_solve(int n_el, gcs_ref_element* el, int n_leaves,
    const int32_t* leaf_elems, const int32_t* edge_offsets, const gcs_ref_edge* edges, int32_t* status)
{
    std::vector<std::shared_ptr<Gcs::Element>> elems;
    for (int i = 0; i < n_el; ++i) {
        std::shared_ptr<Gcs::Element> e;
        if (el[i].type == 0) {
            e = std::make_shared<Gcs::Element>(Gcs::Point(Vector2d(el[i].canvas[0], el[i].canvas[1])));
            if (el[i].is_set) e->updateElementPosition(Vector2d(el[i].pos[0], el[i].pos[1]));
        } else {
            e = std::make_shared<Gcs::Element>(Gcs::Line(Vector2d(el[i].canvas[0], el[i].canvas[1]), Vector2d(el[i].canvas[2], el[i].canvas[3])));
            if (el[i].is_set) e->updateElementPosition(Vector2d(el[i].pos[0], el[i].pos[1]), Vector2d(el[i].pos[2], el[i].pos[3]));
        }
        elems.push_back(e);
    }
    int rc = 0;
    for (int l = 0; l < n_leaves; ++l) status[l] = -2;  // not reached
    for (int l = 0; l < n_leaves && rc == 0; ++l) {
        try {
            Gcs::ConstraintGraph g;
            Gcs::ConstraintGraph::NodeIdType nodes[3];
            for (int i = 0; i < 3; ++i) {
                nodes[i] = g.getGraph().addNode();
                g.addElement(nodes[i], elems[leaf_elems[3 * l + i]]);
            }
            auto nodeOf = [&](int global) {
                for (int i = 0; i < 3; ++i)
                    if (leaf_elems[3 * l + i] == global) return nodes[i];
                throw std::runtime_error("edge endpoint outside the leaf");
            };
            for (int k = edge_offsets[l]; k < edge_offsets[l + 1]; ++k) {
                const auto& ed = edges[k];
                if (ed.type == 2) {
                    g.addVirtualEdge(nodeOf(ed.a), nodeOf(ed.b));
                    continue;
                }
                auto eid = g.getGraph().addEdge(nodeOf(ed.a), nodeOf(ed.b)).value();
                if (ed.type == 0)
                    g.addConstraint(eid, std::make_shared<Gcs::Constraint>(Gcs::DistanceConstraint(ed.value)));
                else
                    g.addConstraint(eid, std::make_shared<Gcs::Constraint>(Gcs::AngleConstraint(ed.value, ed.flip != 0)));
            }
            status[l] = (int32_t)Gcs::classifyAndSolve(g).status;
        } catch (...) {
            status[l] = -1;
            rc = -1;
        }
    }
    for (int i = 0; i < n_el; ++i) {
        el[i].is_set = elems[i]->isElementSet() ? 1 : 0;
        if (el[i].type == 0) {
            const auto& p = elems[i]->getElement<Gcs::Point>();
            el[i].pos[0] = p.position.x(), el[i].pos[1] = p.position.y();
        } else {
            const auto& ln = elems[i]->getElement<Gcs::Line>();
            el[i].pos[0] = ln.p1.x(), el[i].pos[1] = ln.p1.y(), el[i].pos[2] = ln.p2.x(), el[i].pos[3] = ln.p2.y();
        }
    }
    return rc;
}

}  // extern "C"
