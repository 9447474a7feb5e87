// Eigen::Matrix2d as RigidTransform uses it (reference:
// src/solving/bottom_up/merge3_solver_common.hpp:19-22).  Real Eigen is used when the build finds
// it; this image has none, so a minimal 2x2 value type stands in: coefficient access, identity,
// transpose, determinant, the coefficient-wise 2x2 product and matrix * vector, each arithmetic
// step one IEEE rounding in the order Eigen's lazy 2x2 product evaluates it.
#pragma once

#include <gcs/math/vector2d.hpp>

#if !(__has_include(<Eigen/src/Core/Matrix.h>) && !defined(GCS_B200_NO_EIGEN))
namespace Eigen {

class Matrix2d {
public:
    Matrix2d() : m_d { { 0.0, 0.0 }, { 0.0, 0.0 } } {}
    static Matrix2d Zero() { return Matrix2d(); }
    static Matrix2d Identity()
    {
        Matrix2d m;
        m.m_d[0][0] = 1.0, m.m_d[1][1] = 1.0;
        return m;
    }
    double& operator()(int i, int j) { return m_d[i][j]; }
    double operator()(int i, int j) const { return m_d[i][j]; }
    Matrix2d transpose() const
    {
        Matrix2d t;
        t.m_d[0][0] = m_d[0][0], t.m_d[0][1] = m_d[1][0], t.m_d[1][0] = m_d[0][1], t.m_d[1][1] = m_d[1][1];
        return t;
    }
    double determinant() const { return m_d[0][0] * m_d[1][1] - m_d[1][0] * m_d[0][1]; }
    friend Matrix2d operator*(const Matrix2d& a, const Matrix2d& b)
    {
        Matrix2d r;
        for (int i = 0; i < 2; ++i)
            for (int j = 0; j < 2; ++j) r.m_d[i][j] = a.m_d[i][0] * b.m_d[0][j] + a.m_d[i][1] * b.m_d[1][j];
        return r;
    }
    friend Vector2d operator*(const Matrix2d& a, const Vector2d& v)
    {
        return Vector2d(a.m_d[0][0] * v.x() + a.m_d[0][1] * v.y(), a.m_d[1][0] * v.x() + a.m_d[1][1] * v.y());
    }

private:
    double m_d[2][2];
};

}  // namespace Eigen
#endif
