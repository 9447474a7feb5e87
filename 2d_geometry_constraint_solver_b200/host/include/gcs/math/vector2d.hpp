// Eigen::Vector2d as the boundary types use it (reference: model/elements.hpp:24-94).  Real Eigen
// is used when the build finds it; this image has none, so a minimal value type with the members
// the solver path touches stands in.  Every operation is one IEEE rounding per arithmetic step, in
// the order Eigen evaluates it for a fixed 2-vector (dot = x*x' + y*y'; norm = sqrt(squaredNorm);
// normalized: z = squaredNorm, z > 0 ? v / sqrt(z) : v; v / s is a true division).
#pragma once

#if __has_include(<Eigen/src/Core/Matrix.h>) && !defined(GCS_B200_NO_EIGEN)
#include <Eigen/Core>
#else
#include <cmath>

namespace Eigen {

class Vector2d {
public:
    Vector2d() : m_x(0.0), m_y(0.0) {}
    Vector2d(double x, double y) : m_x(x), m_y(y) {}
    static Vector2d Zero() { return Vector2d(0.0, 0.0); }

    double& x() { return m_x; }
    double& y() { return m_y; }
    double x() const { return m_x; }
    double y() const { return m_y; }
    double& operator()(int i) { return i == 0 ? m_x : m_y; }
    double operator()(int i) const { return i == 0 ? m_x : m_y; }
    double& operator[](int i) { return i == 0 ? m_x : m_y; }
    double operator[](int i) const { return i == 0 ? m_x : m_y; }

    double dot(const Vector2d& o) const { return m_x * o.m_x + m_y * o.m_y; }
    double squaredNorm() const { return m_x * m_x + m_y * m_y; }
    double norm() const { return std::sqrt(squaredNorm()); }
    Vector2d normalized() const
    {
        const double z = squaredNorm();
        if (z > 0.0) {
            const double n = std::sqrt(z);
            return Vector2d(m_x / n, m_y / n);
        }
        return *this;
    }
    Vector2d operator-() const { return Vector2d(-m_x, -m_y); }
    Vector2d& operator+=(const Vector2d& o) { m_x += o.m_x; m_y += o.m_y; return *this; }
    Vector2d& operator-=(const Vector2d& o) { m_x -= o.m_x; m_y -= o.m_y; return *this; }
    Vector2d& operator/=(double s) { m_x /= s; m_y /= s; return *this; }
    Vector2d& operator*=(double s) { m_x *= s; m_y *= s; return *this; }
    // v.transpose(): a row view, only ever the right operand of an outer product u * v.transpose()
    struct Row {
        double a, b;
    };
    Row transpose() const { return { m_x, m_y }; }
    friend Vector2d operator+(const Vector2d& a, const Vector2d& b) { return Vector2d(a.m_x + b.m_x, a.m_y + b.m_y); }
    friend Vector2d operator-(const Vector2d& a, const Vector2d& b) { return Vector2d(a.m_x - b.m_x, a.m_y - b.m_y); }
    friend Vector2d operator*(double s, const Vector2d& a) { return Vector2d(s * a.m_x, s * a.m_y); }
    friend Vector2d operator*(const Vector2d& a, double s) { return Vector2d(a.m_x * s, a.m_y * s); }
    friend Vector2d operator/(const Vector2d& a, double s) { return Vector2d(a.m_x / s, a.m_y / s); }
    friend bool operator==(const Vector2d& a, const Vector2d& b) { return a.m_x == b.m_x && a.m_y == b.m_y; }

private:
    double m_x, m_y;
};

}  // namespace Eigen
#endif
