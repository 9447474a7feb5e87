// ref_model_driver.cpp — C entry point over the REFERENCE'S OWN gui/src/constraint_model.cpp
// (compiled where it lies under /root/reference; it has no GTK dependency), for the step that
// follows the solve: ConstraintModel::applySolverToCanvasTransform (constraint_model.cpp:394-501),
// reached through the public solveConstraintSystem() (:362-382).  TEST INFRASTRUCTURE ONLY.
//
// solveConstraintSystem() builds a strategy, runs GeometricConstraintSystem and then applies the
// solver->canvas transform.  The decomposition / orchestration translation units need OGDF and
// GCC >= 14, so here GeometricConstraintSystem::solveGeometricConstraintSystem is a hook that
// installs the solver-space positions the caller supplies (the "solve" is an input of this test),
// and the two strategy classes get inert virtuals so their vtables exist.  Everything after the
// hook - pairing, centroids, covariance, SVD, reflection fix, application to points and lines - is
// the reference's code; the SVD arithmetic is the Eigen stand-in's (oracle/ref_shim).
#include <cstdint>
#include <functional>
#include <stdexcept>
#include <vector>

#include "constraint_model.hpp"
#include <gcs/decomposition/bottom_up/bottom_up_strategy.hpp>
#include <gcs/decomposition/top_down/stree_top_down_strategy.hpp>
#include <gcs/orchestration/geometric_constraint_system.hpp>

namespace {
thread_local std::function<void(Gcs::ConstraintGraph&)> g_solve_hook;
[[noreturn]] void notBuilt() { throw std::logic_error("decomposition is not part of the reference test build"); }
}  // namespace

namespace Gcs {
void GeometricConstraintSystem::solveGeometricConstraintSystem(ConstraintGraph& g)
{
    if (g_solve_hook) g_solve_hook(g);
}
Constrainedness DeficitStreeBasedTopDownStrategy::checkConstraintGraphConstrainedness(const ConstraintGraph&) { notBuilt(); }
bool DeficitStreeBasedTopDownStrategy::resolve(ConstraintGraph&) { notBuilt(); }
std::vector<ConstraintGraph> DeficitStreeBasedTopDownStrategy::decomposeConstraintGraph(ConstraintGraph&) { notBuilt(); }
void DeficitStreeBasedTopDownStrategy::solveGcs(std::vector<ConstraintGraph>&) { notBuilt(); }
Constrainedness BottomUpDrPlanStrategy::checkConstraintGraphConstrainedness(const ConstraintGraph&) { notBuilt(); }
bool BottomUpDrPlanStrategy::resolve(ConstraintGraph&) { notBuilt(); }
std::vector<ConstraintGraph> BottomUpDrPlanStrategy::decomposeConstraintGraph(ConstraintGraph&) { notBuilt(); }
void BottomUpDrPlanStrategy::solveGcs(std::vector<ConstraintGraph>&) { notBuilt(); }
// referenced by ConstraintModel::removeElement / removeConstraint, never called here
ConstraintGraphError ConstraintGraph::removeElement(NodeIdType) { notBuilt(); }
ConstraintGraphError ConstraintGraph::removeConstraintEdge(EdgeIdType) { notBuilt(); }
}  // namespace Gcs

// Elements i = 0..n_el-1 in insertion order (type 0 point: x,y; 1 line: x1,y1,x2,y2).
// constraints: (a, b, type 0 distance / 1 angle, flip) + value (distance, or angle in DEGREES as
// the GUI and the JSON file store it).  solved[i] != 0: element i gets solver position pos4[i].
// Outputs: canvas4 after solveConstraintSystem(); angle_rad[k] = what the graph stores for
// constraint k (NaN for a rejected constraint).  Returns the number of accepted constraints.
extern "C" __attribute__((visibility("default"))) int gcs_ref_model_solve_transform(int n_el, const int32_t* type,
    const double* canvas4, const double* pos4, const uint8_t* solved, int n_con, const int32_t* con4, const double* value,
    double* canvas4_out, double* stored_value)
{
    Gui::ConstraintModel model;
    std::vector<Gui::ElementId> ids;
    for (int i = 0; i < n_el; ++i) {
        const double* c = canvas4 + 4 * i;
        ids.push_back(type[i] == 0 ? model.addPoint(c[0], c[1]) : model.addLine(c[0], c[1], c[2], c[3]));
    }
    int accepted = 0;
    for (int k = 0; k < n_con; ++k) {
        const int32_t* c = con4 + 4 * k;
        const auto id = (c[2] == 0) ? model.addDistanceConstraint(ids[c[0]], ids[c[1]], value[k])
                                    : model.addAngleConstraint(ids[c[0]], ids[c[1]], value[k], c[3] != 0);
        stored_value[k] = id ? model.getConstraintValue(*id).value_or(__builtin_nan("")) : __builtin_nan("");
        accepted += id.has_value();
    }
    g_solve_hook = [&](Gcs::ConstraintGraph& g) {
        int i = 0;
        for (const auto& [node, element] : g.getElementMap()) {  // ascending node id = insertion order
            const double* p = pos4 + 4 * i;
            if (solved[i]) {
                if (type[i] == 0)
                    element->updateElementPosition(Eigen::Vector2d(p[0], p[1]));
                else
                    element->updateElementPosition(Eigen::Vector2d(p[0], p[1]), Eigen::Vector2d(p[2], p[3]));
            }
            ++i;
        }
    };
    model.solveConstraintSystem();
    g_solve_hook = nullptr;
    for (int i = 0; i < n_el; ++i) {
        double* o = canvas4_out + 4 * i;
        o[0] = o[1] = o[2] = o[3] = 0.0;
        if (type[i] == 0) {
            const auto p = model.getPointCanvasPosition(ids[i]);
            if (p) o[0] = p->first, o[1] = p->second;
        } else {
            const auto l = model.getLineCanvasEndpoints(ids[i]);
            if (l) o[0] = l->first.first, o[1] = l->first.second, o[2] = l->second.first, o[3] = l->second.second;
        }
    }
    return accepted;
}
