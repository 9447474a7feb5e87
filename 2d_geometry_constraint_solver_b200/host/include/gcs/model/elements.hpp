// Geometry element value types kept at the boundary (reference:
// includes/gcs/model/elements.hpp:24-158, src/model/elements.cpp).  Same names, members and
// semantics: `canvas*` is the user's sketch, `position` / `p1,p2` the solver-space result;
// Element::updateElementPosition stores the result and marks the element solved (m_isSet).
#pragma once

#include <string>
#include <utility>
#include <variant>

#include <gcs/export.hpp>
#include <gcs/math/vector2d.hpp>

namespace Gcs {

struct GCS_API Point {
    Eigen::Vector2d canvasPosition;
    Eigen::Vector2d position;

    Point();
    explicit Point(const Eigen::Vector2d& canvasPos);
    std::string getTypeName() const;
    std::string toString() const;
    void updateElementPosition(const Eigen::Vector2d& newPosition);
};

struct GCS_API FixedRadiusCircle {
    Eigen::Vector2d position;
    double fixedRadius;

    FixedRadiusCircle();
    explicit FixedRadiusCircle(const Eigen::Vector2d& centerPos, double r);
    std::string getTypeName() const;
    std::string toString() const;
    void updateElementPosition(const Eigen::Vector2d& newPosition);
};

struct GCS_API Line {
    Eigen::Vector2d canvasP1;
    Eigen::Vector2d canvasP2;
    Eigen::Vector2d p1;
    Eigen::Vector2d p2;

    Line();
    explicit Line(const Eigen::Vector2d& canvasEndpoint1, const Eigen::Vector2d& canvasEndpoint2);
    std::string getTypeName() const;
    std::string toString() const;
    void updateElementPosition(const Eigen::Vector2d& newP1, const Eigen::Vector2d& newP2);

    Eigen::Vector2d direction() const;      // p2 - p1
    Eigen::Vector2d unitDirection() const;  // (p2 - p1).normalized()
    Eigen::Vector2d normal() const;         // (-dir.y, dir.x)
    double length() const;                  // (p2 - p1).norm()
    Eigen::Vector2d midpoint() const;       // (p1 + p2) / 2.0
};

using ElementVariant = std::variant<Point, FixedRadiusCircle, Line>;

class GCS_API Element final {
public:
    template <typename T>
    explicit Element(const T& e) : m_element { e } {}

    template <typename T>
    bool isElementType() const { return std::holds_alternative<T>(m_element); }
    template <typename T>
    T& getElement() { return std::get<T>(m_element); }
    template <typename T>
    const T& getElement() const { return std::get<T>(m_element); }

    std::string getElementName() const;
    std::string toString() const;
    bool isElementSet() const { return m_isSet; }

    // Point / circle: one vector; Line: two vectors.  A call whose arguments do not fit the
    // active alternative is ignored (the reference asserts), the element stays unset.
    template <typename... Parameters>
    void updateElementPosition(Parameters&&... params)
    {
        std::visit(
            [&](auto& elem) {
                if constexpr (requires { elem.updateElementPosition(std::forward<Parameters>(params)...); }) {
                    elem.updateElementPosition(std::forward<Parameters>(params)...);
                    m_isSet = true;
                }
            },
            m_element);
    }

private:
    ElementVariant m_element;
    bool m_isSet = false;
};

}  // namespace Gcs
