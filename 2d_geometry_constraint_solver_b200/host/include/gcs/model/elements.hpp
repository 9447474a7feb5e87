// Geometry element value types kept at the boundary (reference:
// includes/gcs/model/elements.hpp:24-158, src/model/elements.cpp).  Same type names, member names
// and semantics, because host code written against the reference must compile unchanged:
//   canvas*            where the user drew the element (never touched by the solver path; the
//                      solver->canvas transform of gcs/b200/canvas_transform.hpp rewrites it)
//   position / p1,p2   the solver-space result
//   Element            a variant over the shapes + the "already solved" flag (m_isSet), which
//                      Element::updateElementPosition sets and the eight matches() predicates read.
#pragma once

#include <cstdint>

#include <string>
#include <utility>
#include <variant>

#include <gcs/export.hpp>
#include <gcs/math/vector2d.hpp>

namespace Gcs {

struct GCS_API Point {
    Point();
    explicit Point(const Eigen::Vector2d& canvasPos);

    Eigen::Vector2d canvasPosition;
    Eigen::Vector2d position;

    void updateElementPosition(const Eigen::Vector2d& newPosition);
    std::string getTypeName() const;  // "Point"
    std::string toString() const;
};

// Part of the variant in the reference; no sub-problem solver reads or writes it.
struct GCS_API FixedRadiusCircle {
    FixedRadiusCircle();
    explicit FixedRadiusCircle(const Eigen::Vector2d& centerPos, double r);

    Eigen::Vector2d position;
    double fixedRadius;

    void updateElementPosition(const Eigen::Vector2d& newPosition);
    std::string getTypeName() const;  // "FixedRadiusCircle"
    std::string toString() const;
};

struct GCS_API Line {
    Line();
    explicit Line(const Eigen::Vector2d& canvasEndpoint1, const Eigen::Vector2d& canvasEndpoint2);

    Eigen::Vector2d canvasP1;
    Eigen::Vector2d canvasP2;
    Eigen::Vector2d p1;
    Eigen::Vector2d p2;

    void updateElementPosition(const Eigen::Vector2d& newP1, const Eigen::Vector2d& newP2);
    std::string getTypeName() const;  // "Line"
    std::string toString() const;

    // solver-space helpers (the packer reads length() and midpoint(); operation order in elements.cpp)
    Eigen::Vector2d direction() const;      // p2 - p1
    double length() const;                  // |p2 - p1|
    Eigen::Vector2d unitDirection() const;  // direction / length, or direction itself when it is zero
    Eigen::Vector2d normal() const;         // (-direction.y, direction.x), not normalised
    Eigen::Vector2d midpoint() const;       // (p1 + p2) / 2
};

using ElementVariant = std::variant<Point, FixedRadiusCircle, Line>;

class GCS_API Element final {
public:
    template <typename Shape>
    explicit Element(const Shape& shape) : m_element(shape) {}

    bool isElementSet() const { return m_isSet; }

    template <typename Shape>
    bool isElementType() const { return std::holds_alternative<Shape>(m_element); }
    template <typename Shape>
    const Shape& getElement() const { return std::get<Shape>(m_element); }
    template <typename Shape>
    Shape& getElement() { return std::get<Shape>(m_element); }

    // One vector for a point / circle, two for a line; marks the element solved.  Arguments that
    // do not fit the active alternative leave the element untouched and unset (the reference
    // asserts there).
    template <typename... Args>
    void updateElementPosition(Args&&... args)
    {
        std::visit(
            [&](auto& shape) {
                if constexpr (requires { shape.updateElementPosition(std::forward<Args>(args)...); }) {
                    shape.updateElementPosition(std::forward<Args>(args)...);
                    m_isSet = true;
                }
            },
            m_element);
    }

    std::string getElementName() const;
    std::string toString() const;

    // Serial number of this Element OBJECT: unique in the process, handed out in construction order,
    // so the elements of one sketch sit in a narrow range of it.  The batched leaf scheduler
    // (gcs/b200/leaf_batch.hpp) indexes its per-element tables by it instead of hashing pointers.
    // Not part of the element's value: a copy is a new object with a new number, assignment keeps
    // the number of the object assigned to.
    std::uint64_t serial() const { return m_serial; }
    // leaves `count` numbers unused (tests: elements whose numbers lie far apart)
    static void skipSerials(std::uint64_t count);

    Element(const Element& other) : m_element(other.m_element), m_isSet(other.m_isSet) {}
    Element(Element&& other) noexcept : m_element(std::move(other.m_element)), m_isSet(other.m_isSet) {}
    Element& operator=(const Element& other)
    {
        m_element = other.m_element, m_isSet = other.m_isSet;
        return *this;
    }
    Element& operator=(Element&& other) noexcept
    {
        m_element = std::move(other.m_element), m_isSet = other.m_isSet;
        return *this;
    }

private:
    ElementVariant m_element;
    bool m_isSet = false;
    const std::uint64_t m_serial = nextSerial();
    static std::uint64_t nextSerial();
};

}  // namespace Gcs
