// Client of the reference's GUI model (gui/src/constraint_model.cpp), which is compiled UNMODIFIED
// from /root/reference against THIS repo's host headers and linked with libgcs_host.so: the
// drop-in check of SURVEY.md section 8b at the level a maintainer would do it (swap the include
// path and the library, rebuild the client).  Builds the 3-4-5 triangle of BASELINE configs[0]
// plus a point-point-line cluster; `solve` runs ConstraintModel::solveConstraintSystem() (decompose
// -> batched device solve -> solver->canvas transform) and prints the canvas coordinates.
#include <cmath>
#include <cstdio>
#include <cstring>

#include "constraint_model.hpp"

int main(int argc, char** argv)
{
    const bool solve = argc > 1 && std::strcmp(argv[1], "solve") == 0;
    Gui::ConstraintModel model;
    const auto a = model.addPoint(100.0, 100.0);
    const auto b = model.addPoint(200.0, 100.0);
    const auto c = model.addPoint(150.0, 200.0);
    const auto l = model.addLine(90.0, 300.0, 260.0, 310.0);
    bool ok = model.addDistanceConstraint(a, b, 3.0).has_value();
    ok = model.addDistanceConstraint(b, c, 4.0).has_value() && ok;
    ok = model.addDistanceConstraint(a, c, 5.0).has_value() && ok;
    ok = model.addDistanceConstraint(a, l, 2.0).has_value() && ok;
    ok = model.addDistanceConstraint(c, l, 1.0).has_value() && ok;
    ok = !model.addAngleConstraint(a, l, 30.0).has_value() && ok;  // a point is not a line: rejected
    if (!ok) {
        std::puts("constraint acceptance differs");
        return 2;
    }
    std::printf("status %s\n", model.getStatusText().c_str());
    if (!solve) return 0;
    const std::string err = model.solveConstraintSystem();
    if (!err.empty()) {
        std::printf("solve failed: %s\n", err.c_str());
        return 3;
    }
    for (auto id : { a, b, c }) {
        const auto p = model.getPointCanvasPosition(id);
        if (!p || !model.isElementSolved(id)) return 4;
        std::printf("point %.17g %.17g\n", p->first, p->second);
    }
    const auto e = model.getLineCanvasEndpoints(l);
    if (!e || !model.isElementSolved(l)) return 5;
    std::printf("line %.17g %.17g %.17g %.17g\n", e->first.first, e->first.second, e->second.first, e->second.second);
    return 0;
}
