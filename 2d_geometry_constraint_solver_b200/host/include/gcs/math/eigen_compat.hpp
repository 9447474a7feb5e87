// The slice of Eigen's 2-D fixed-size interface that code written against the reference uses
// besides Vector2d (gui/src/constraint_model.cpp:394-501, merge3_solver_common.cpp:96-160):
// compound assignment on vectors, the outer product u * v.transpose(), Matrix2d accumulation and
// JacobiSVD<Matrix2d>(m, ComputeFullU | ComputeFullV).  Only for builds without real Eigen; with
// Eigen on the include path none of this is seen.
#pragma once

#include <gcs/math/matrix2d.hpp>
#include <gcs/math/svd2x2.hpp>

#if !(__has_include(<Eigen/src/Core/Matrix.h>) && !defined(GCS_B200_NO_EIGEN)) || defined(GCS_B200_EIGEN_COMPAT)
namespace Eigen {

enum : unsigned { ComputeFullU = 0x04, ComputeFullV = 0x10 };

inline Matrix2d operator*(const Vector2d& u, const Vector2d::Row& vt)
{
    Matrix2d m;
    m(0, 0) = u.x() * vt.a, m(0, 1) = u.x() * vt.b;
    m(1, 0) = u.y() * vt.a, m(1, 1) = u.y() * vt.b;
    return m;
}

inline Matrix2d& operator+=(Matrix2d& m, const Matrix2d& o)
{
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j) m(i, j) += o(i, j);
    return m;
}

template <typename M>
class JacobiSVD;

template <>
class JacobiSVD<Matrix2d> {
public:
    JacobiSVD(const Matrix2d& m, unsigned /*ComputeFullU | ComputeFullV*/) { Gcs::Math::jacobiSvd2x2(m, m_u, m_v); }
    const Matrix2d& matrixU() const { return m_u; }
    const Matrix2d& matrixV() const { return m_v; }

private:
    Matrix2d m_u, m_v;
};

}  // namespace Eigen
#endif
