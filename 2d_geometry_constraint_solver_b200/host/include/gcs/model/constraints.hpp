// Constraint value types (reference: includes/gcs/model/constraints.hpp:39-124).  Angles are
// radians in the model (degrees only in the GUI / JSON).
#pragma once

#include <expected>
#include <string>
#include <variant>

#include <gcs/export.hpp>

namespace Gcs {

enum class ConstraintError { NoValue };

struct GCS_API DistanceConstraint {
    double distance;
    explicit DistanceConstraint(double d) : distance { d } {}
    std::string getTypeName() const { return "Distance"; }
    std::expected<double, ConstraintError> getConstraintValue() const { return distance; }
};

struct GCS_API TangencyConstraint {
    double angle;
    explicit TangencyConstraint(double d) : angle { d } {}
    std::string getTypeName() const { return "Tangency"; }
    std::expected<double, ConstraintError> getConstraintValue() const { return angle; }
};

struct GCS_API AngleConstraint {
    double angle;
    bool flipOrientation = false;
    explicit AngleConstraint(double a, bool flip = false) : angle { a }, flipOrientation { flip } {}
    std::string getTypeName() const { return "Angle"; }
    std::expected<double, ConstraintError> getConstraintValue() const { return angle; }
};

struct GCS_API PointOnLineConstraint {
    explicit PointOnLineConstraint() {}
    std::string getTypeName() const { return "PointOnLine"; }
    std::expected<double, ConstraintError> getConstraintValue() const { return std::unexpected(ConstraintError::NoValue); }
};

struct GCS_API VirtualConstraint {
    explicit VirtualConstraint() {}
    std::string getTypeName() const { return "Virtual"; }
    std::expected<double, ConstraintError> getConstraintValue() const { return std::unexpected(ConstraintError::NoValue); }
};

using ConstraintVariant
    = std::variant<DistanceConstraint, TangencyConstraint, AngleConstraint, PointOnLineConstraint, VirtualConstraint>;

class GCS_API Constraint final {
public:
    template <typename T>
    explicit Constraint(const T& c) : m_constraint { c } {}

    template <typename T>
    bool isConstraintType() const { return std::holds_alternative<T>(m_constraint); }
    template <typename T>
    const T* getConstraintAs() const { return std::get_if<T>(&m_constraint); }
    template <typename T>
    T* getConstraintAs() { return std::get_if<T>(&m_constraint); }

    std::string getConstraintName() const
    {
        return std::visit([](const auto& c) { return c.getTypeName(); }, m_constraint);
    }
    std::expected<double, ConstraintError> getConstraintValue() const
    {
        return std::visit([](const auto& c) { return c.getConstraintValue(); }, m_constraint);
    }

private:
    ConstraintVariant m_constraint;
};

}  // namespace Gcs
