// Constraint value types at the boundary (reference: includes/gcs/model/constraints.hpp:39-124).
// A constraint is a tagged value on an edge of the constraint graph; the solver path reads
// distances (between two points or a point and a line) and angles (between two lines, RADIANS in
// the model - degrees exist only in the GUI and its JSON files).  The alternatives without a value
// answer getConstraintValue() with ConstraintError::NoValue, which is what makes
// ConstraintGraph::getConstraintBetweenNodes(...)->getConstraintValue().value() throw on a virtual
// edge, a quirk the packer reproduces.
#pragma once

#include <expected>
#include <string>
#include <variant>

#include <gcs/export.hpp>

namespace Gcs {

enum class ConstraintError { NoValue };

namespace detail {
using ConstraintValue = std::expected<double, ConstraintError>;
inline ConstraintValue noValue() { return std::unexpected(ConstraintError::NoValue); }
}  // namespace detail

// |P - Q| = distance, or the unsigned point-to-line distance
struct GCS_API DistanceConstraint {
    explicit DistanceConstraint(double d) : distance(d) {}
    double distance;

    detail::ConstraintValue getConstraintValue() const { return distance; }
    std::string getTypeName() const { return "Distance"; }
};

// angle between two lines in radians; flipOrientation selects the supplementary side
// (line_angle_solvers.cpp:322-326 negates the canvas direction of the free line when it is set)
struct GCS_API AngleConstraint {
    explicit AngleConstraint(double a, bool flip = false) : angle(a), flipOrientation(flip) {}
    double angle;
    bool flipOrientation = false;

    detail::ConstraintValue getConstraintValue() const { return angle; }
    std::string getTypeName() const { return "Angle"; }
};

// declared by the reference, used by no solver
struct GCS_API TangencyConstraint {
    explicit TangencyConstraint(double d) : angle(d) {}
    double angle;

    detail::ConstraintValue getConstraintValue() const { return angle; }
    std::string getTypeName() const { return "Tangency"; }
};

// the two value-less alternatives
struct GCS_API PointOnLineConstraint {
    explicit PointOnLineConstraint() {}
    detail::ConstraintValue getConstraintValue() const { return detail::noValue(); }
    std::string getTypeName() const { return "PointOnLine"; }
};

struct GCS_API VirtualConstraint {
    explicit VirtualConstraint() {}
    detail::ConstraintValue getConstraintValue() const { return detail::noValue(); }
    std::string getTypeName() const { return "Virtual"; }
};

// alternative order = the reference's (index() is observable)
using ConstraintVariant
    = std::variant<DistanceConstraint, TangencyConstraint, AngleConstraint, PointOnLineConstraint, VirtualConstraint>;

class GCS_API Constraint final {
public:
    template <typename Kind>
    explicit Constraint(const Kind& value) : m_constraint(value) {}

    template <typename Kind>
    bool isConstraintType() const { return std::holds_alternative<Kind>(m_constraint); }
    template <typename Kind>
    Kind* getConstraintAs() { return std::get_if<Kind>(&m_constraint); }
    template <typename Kind>
    const Kind* getConstraintAs() const { return std::get_if<Kind>(&m_constraint); }

    detail::ConstraintValue getConstraintValue() const
    {
        return std::visit([](const auto& held) { return held.getConstraintValue(); }, m_constraint);
    }
    std::string getConstraintName() const
    {
        return std::visit([](const auto& held) { return held.getTypeName(); }, m_constraint);
    }

private:
    ConstraintVariant m_constraint;
};

}  // namespace Gcs
