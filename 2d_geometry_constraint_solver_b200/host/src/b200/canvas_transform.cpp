// Solver -> canvas rigid motion (reference: gui/src/constraint_model.cpp:394-501).
#include <gcs/b200/canvas_transform.hpp>

#include <vector>

#include <gcs/math/matrix2d.hpp>
#include <gcs/math/svd2x2.hpp>

using Eigen::Matrix2d;
using Eigen::Vector2d;

namespace Gcs::B200 {

namespace {

template <typename Map>
void moveSolved(const Map& elements, const Matrix2d* rot, const Vector2d& shift)
{
    auto place = [&](const Vector2d& p) { return rot ? (*rot) * p + shift : p + shift; };
    for (const auto& [node, element] : elements) {
        if (!element || !element->isElementSet()) continue;
        if (element->template isElementType<Point>()) {
            auto& pt = element->template getElement<Point>();
            pt.canvasPosition = place(pt.position);
        } else if (element->template isElementType<Line>()) {
            auto& ln = element->template getElement<Line>();
            ln.canvasP1 = place(ln.p1);
            ln.canvasP2 = place(ln.p2);
        }
    }
}

}  // namespace

void applySolverToCanvasTransform(ConstraintGraph& graph)
{
    const auto& elements = graph.getElementMap();
    std::vector<Vector2d> solver, canvas;  // solved points only, ascending node id
    for (const auto& [node, element] : elements) {
        if (!element || !element->isElementSet() || !element->isElementType<Point>()) continue;
        const auto& pt = element->getElement<Point>();
        solver.push_back(pt.position), canvas.push_back(pt.canvasPosition);
    }
    const std::size_t n = solver.size();
    if (n == 0) return;
    if (n == 1) {  // rotation undetermined: translate the single point back (:421-440)
        moveSolved(elements, nullptr, canvas[0] - solver[0]);
        return;
    }
    Vector2d cs = Vector2d::Zero(), cc = Vector2d::Zero();
    for (std::size_t i = 0; i < n; ++i) cs += solver[i], cc += canvas[i];
    const double count = static_cast<double>(n);
    cs = cs / count, cc = cc / count;
    Matrix2d h = Matrix2d::Zero();
    for (std::size_t i = 0; i < n; ++i) {
        const Vector2d a = solver[i] - cs, b = canvas[i] - cc;
        h(0, 0) += a.x() * b.x(), h(0, 1) += a.x() * b.y();
        h(1, 0) += a.y() * b.x(), h(1, 1) += a.y() * b.y();
    }
    Matrix2d u, v;
    Math::jacobiSvd2x2(h, u, v);
    Matrix2d fix = Matrix2d::Identity();
    fix(1, 1) = (v * u.transpose()).determinant();  // -1 for a reflection: keeps R a proper rotation
    const Matrix2d rot = (v * fix) * u.transpose();
    moveSolved(elements, &rot, cc - rot * cs);
}

}  // namespace Gcs::B200
