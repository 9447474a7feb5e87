// Members of the element value types (reference: src/constraint_solver/src/model/elements.cpp).
// The packer reads Line::length() / unitDirection() / midpoint() (point_line_solvers.cpp:506,
// :636-673; line_angle_solvers.cpp:551), so those keep Eigen's operation order: norm =
// sqrt(x*x + y*y), normalized = v / norm when the squared norm is positive, midpoint =
// (p1 + p2) / 2.0.  The text forms are the reference's, character for character.
#include <atomic>
#include <format>

#include <gcs/model/elements.hpp>

namespace Gcs {

namespace {

using Vec = Eigen::Vector2d;

Vec origin() { return Vec::Zero(); }

// "x,y" as std::format prints two doubles
std::string xy(const Vec& v) { return std::format("{},{}", v.x(), v.y()); }

}  // namespace

// ---- Point: a canvas position and, once solved, a solver-space position ----
Point::Point() : Point(origin()) {}

Point::Point(const Vec& canvasPos) : canvasPosition(canvasPos), position(origin()) {}

void Point::updateElementPosition(const Vec& newPosition) { position = newPosition; }

std::string Point::getTypeName() const { return "Point"; }

std::string Point::toString() const { return "CanvasCoords(" + xy(canvasPosition) + "), Calculated(" + xy(position) + ")"; }

// ---- FixedRadiusCircle: kept for the variant's shape; no solver handles it ----
FixedRadiusCircle::FixedRadiusCircle() : FixedRadiusCircle(origin(), 0.0) {}

FixedRadiusCircle::FixedRadiusCircle(const Vec& centerPos, double r) : position(centerPos), fixedRadius(r) {}

void FixedRadiusCircle::updateElementPosition(const Vec& newPosition) { position = newPosition; }

std::string FixedRadiusCircle::getTypeName() const { return "FixedRadiusCircle"; }

std::string FixedRadiusCircle::toString() const
{
    return std::format("FixedRadiusCircle(x: {}, y: {}, radius: {})", position.x(), position.y(), fixedRadius);
}

// ---- Line: two canvas endpoints and, once solved, two solver-space endpoints ----
Line::Line() : Line(origin(), origin()) {}

Line::Line(const Vec& canvasEndpoint1, const Vec& canvasEndpoint2)
    : canvasP1(canvasEndpoint1), canvasP2(canvasEndpoint2), p1(origin()), p2(origin())
{
}

void Line::updateElementPosition(const Vec& newP1, const Vec& newP2) { p1 = newP1, p2 = newP2; }

std::string Line::getTypeName() const { return "Line"; }

std::string Line::toString() const
{
    return "Line(CanvasP1:(" + xy(canvasP1) + "), CanvasP2:(" + xy(canvasP2) + "), CalcP1:(" + xy(p1) + "), CalcP2:(" + xy(p2) + "))";
}

Vec Line::direction() const { return p2 - p1; }

double Line::length() const { return direction().norm(); }

Vec Line::unitDirection() const { return direction().normalized(); }

Vec Line::normal() const
{
    const Vec d = direction();
    return { -d.y(), d.x() };
}

Vec Line::midpoint() const { return (p1 + p2) / 2.0; }

// ---- Element: forwards to the active alternative ----
namespace {
std::atomic<std::uint64_t> g_serial { 1 };
}

std::uint64_t Element::nextSerial() { return g_serial.fetch_add(1, std::memory_order_relaxed); }

void Element::skipSerials(std::uint64_t count) { g_serial.fetch_add(count, std::memory_order_relaxed); }

std::string Element::getElementName() const
{
    return std::visit([](const auto& shape) { return shape.getTypeName(); }, m_element);
}

std::string Element::toString() const
{
    return std::visit([](const auto& shape) { return shape.toString(); }, m_element);
}

}  // namespace Gcs
