// Stand-in for the slice of Eigen (conanfile.py:11 `eigen/[~5]`, not vendored) that the reference's
// Newton-Raphson path uses: Vector2d / Matrix2d / Matrix<dual,2,1> value types with the handful of
// members the solvers call, and Matrix2d::colPivHouseholderQr().solve().
//
// TEST INFRASTRUCTURE.  Restated from knowledge of Eigen's published algorithms
// (ColPivHouseholderQR.h, Householder.h, TriangularSolverVector.h, Dot.h), not copied:
//   - a.dot(b) = a0*b0 + a1*b1; squaredNorm = x*x + y*y; norm = sqrt(squaredNorm);
//     normalized(): z = squaredNorm; z > 0 ? v / sqrt(z) : v; v / s is a true division.
//   - colPivHouseholderQr: see `ColPivHouseholderQR2::compute` below.
#pragma once

#include <cfloat>
#include <algorithm>
#include <cmath>
#include <initializer_list>
#include <utility>

namespace Eigen {

template <typename T, int R, int C>
class Matrix;

template <typename T>
class Matrix<T, 2, 1> {
public:
    Matrix() : m_d {} {}
    Matrix(const T& a, const T& b) : m_d { a, b } {}
    static Matrix Zero() { return Matrix(T(0), T(0)); }
    T& x() { return m_d[0]; }
    T& y() { return m_d[1]; }
    const T& x() const { return m_d[0]; }
    const T& y() const { return m_d[1]; }
    T& operator()(int i) { return m_d[i]; }
    const T& operator()(int i) const { return m_d[i]; }
    T& operator[](int i) { return m_d[i]; }
    const T& operator[](int i) const { return m_d[i]; }

    T dot(const Matrix& o) const { return m_d[0] * o.m_d[0] + m_d[1] * o.m_d[1]; }
    T squaredNorm() const { return m_d[0] * m_d[0] + m_d[1] * m_d[1]; }
    T norm() const { using std::sqrt; return sqrt(squaredNorm()); }
    Matrix normalized() const
    {
        using std::sqrt;
        const T z = squaredNorm();
        if (z > T(0)) return *this / sqrt(z);
        return *this;
    }
    Matrix operator-() const { return Matrix(-m_d[0], -m_d[1]); }
    friend Matrix operator+(const Matrix& a, const Matrix& b) { return Matrix(a.m_d[0] + b.m_d[0], a.m_d[1] + b.m_d[1]); }
    friend Matrix operator-(const Matrix& a, const Matrix& b) { return Matrix(a.m_d[0] - b.m_d[0], a.m_d[1] - b.m_d[1]); }
    friend Matrix operator*(const T& s, const Matrix& a) { return Matrix(s * a.m_d[0], s * a.m_d[1]); }
    friend Matrix operator*(const Matrix& a, const T& s) { return Matrix(a.m_d[0] * s, a.m_d[1] * s); }
    friend Matrix operator/(const Matrix& a, const T& s) { return Matrix(a.m_d[0] / s, a.m_d[1] / s); }
    Matrix& operator+=(const Matrix& o) { m_d[0] += o.m_d[0]; m_d[1] += o.m_d[1]; return *this; }
    Matrix& operator-=(const Matrix& o) { m_d[0] -= o.m_d[0]; m_d[1] -= o.m_d[1]; return *this; }
    Matrix& operator/=(const T& s) { m_d[0] /= s; m_d[1] /= s; return *this; }
    Matrix& operator*=(const T& s) { m_d[0] *= s; m_d[1] *= s; return *this; }
    // v.transpose(): only ever used as the right operand of an outer product (merge3_solver_common.cpp:135)
    struct Transposed { T a, b; };
    Transposed transpose() const { return Transposed { m_d[0], m_d[1] }; }

private:
    T m_d[2];
};

using Vector2d = Matrix<double, 2, 1>;

// Eigen::ColPivHouseholderQR<Matrix2d>, restated.  Column-major storage: col(k) = (q[0][k], q[1][k]).
class ColPivHouseholderQR2 {
public:
    explicit ColPivHouseholderQR2(const double m[2][2]) { compute(m); }

    Vector2d solve(const Vector2d& rhs) const
    {
        // _solve_impl: nonzeroPivots(), Q^T applied as H0 then H1 (length = nonzero pivots),
        // column-major upper triangular solve with exact-zero skips, un-permutation.
        Vector2d dst(0.0, 0.0);
        if (m_nonzero_pivots == 0) return dst;
        double c[2] = { rhs(0), rhs(1) };
        for (int k = 0; k < m_nonzero_pivots; ++k) {
            if (k == 1) {
                c[1] *= (1.0 - m_hcoeffs[1]);  // rows()==1 branch of applyHouseholderOnTheLeft
            } else if (m_hcoeffs[0] != 0.0) {
                double tmp = m_qr[1][0] * c[1];  // essential^T * bottom
                tmp += c[0];
                c[0] -= m_hcoeffs[0] * tmp;
                c[1] -= tmp * (m_hcoeffs[0] * m_qr[1][0]);
            }
        }
        const int size = m_nonzero_pivots;
        for (int k = 0; k < size; ++k) {
            const int i = size - k - 1;
            if (c[i] != 0.0) {
                c[i] /= m_qr[i][i];
                const int r = size - k - 1;  // rows above i
                for (int s = 0; s < r; ++s) c[s] -= c[i] * m_qr[s][i];
            }
        }
        for (int i = 0; i < m_nonzero_pivots; ++i) dst(m_perm[i]) = c[i];
        for (int i = m_nonzero_pivots; i < 2; ++i) dst(m_perm[i]) = 0.0;
        return dst;
    }

private:
    void compute(const double m[2][2])
    {
        const int rows = 2, cols = 2, size = 2;
        for (int i = 0; i < 2; ++i)
            for (int j = 0; j < 2; ++j) m_qr[i][j] = m[i][j];
        double norms_updated[2], norms_direct[2];
        for (int k = 0; k < cols; ++k) {
            norms_direct[k] = std::sqrt(m_qr[0][k] * m_qr[0][k] + m_qr[1][k] * m_qr[1][k]);
            norms_updated[k] = norms_direct[k];
        }
        double maxn = norms_updated[0];
        if (norms_updated[1] > maxn) maxn = norms_updated[1];
        const double th = maxn * DBL_EPSILON;
        const double threshold_helper = (th * th) / double(rows);
        const double norm_downdate_threshold = std::sqrt(DBL_EPSILON);
        m_nonzero_pivots = size;
        int transpositions[2] = { 0, 1 };
        for (int k = 0; k < size; ++k) {
            int biggest = k;
            for (int j = k + 1; j < cols; ++j)
                if (norms_updated[j] > norms_updated[biggest]) biggest = j;
            const double biggest_sq = norms_updated[biggest] * norms_updated[biggest];
            if (m_nonzero_pivots == size && biggest_sq < threshold_helper * double(rows - k)) m_nonzero_pivots = k;
            transpositions[k] = biggest;
            if (k != biggest) {
                for (int i = 0; i < rows; ++i) std::swap(m_qr[i][k], m_qr[i][biggest]);
                std::swap(norms_updated[k], norms_updated[biggest]);
                std::swap(norms_direct[k], norms_direct[biggest]);
            }
            // makeHouseholderInPlace on col(k).tail(rows-k)
            double beta, tau;
            {
                double tail_sq = 0.0;
                for (int i = k + 1; i < rows; ++i) tail_sq += m_qr[i][k] * m_qr[i][k];
                const double c0 = m_qr[k][k];
                if (rows - k == 1 || tail_sq <= DBL_MIN) {
                    tau = 0.0;
                    beta = c0;
                    for (int i = k + 1; i < rows; ++i) m_qr[i][k] = 0.0;
                } else {
                    beta = std::sqrt(c0 * c0 + tail_sq);
                    if (c0 >= 0.0) beta = -beta;
                    for (int i = k + 1; i < rows; ++i) m_qr[i][k] = m_qr[i][k] / (c0 - beta);
                    tau = (beta - c0) / beta;
                }
            }
            m_hcoeffs[k] = tau;
            m_qr[k][k] = beta;
            // bottomRightCorner(rows-k, cols-k-1).applyHouseholderOnTheLeft(essential, tau, ws)
            if (rows - k > 1 && tau != 0.0) {
                for (int j = k + 1; j < cols; ++j) {
                    double tmp = m_qr[k + 1][k] * m_qr[k + 1][j];
                    tmp += m_qr[k][j];
                    m_qr[k][j] -= tau * tmp;
                    m_qr[k + 1][j] -= tmp * (tau * m_qr[k + 1][k]);
                }
            }
            // norm down-date
            for (int j = k + 1; j < cols; ++j) {
                if (norms_updated[j] != 0.0) {
                    double temp = std::abs(m_qr[k][j]) / norms_updated[j];
                    temp = (1.0 + temp) * (1.0 - temp);
                    temp = temp < 0.0 ? 0.0 : temp;
                    const double ratio = norms_updated[j] / norms_direct[j];
                    const double temp2 = temp * (ratio * ratio);
                    if (temp2 <= norm_downdate_threshold) {
                        double sq = 0.0;
                        for (int i = k + 1; i < rows; ++i) sq += m_qr[i][j] * m_qr[i][j];
                        norms_direct[j] = std::sqrt(sq);
                        norms_updated[j] = norms_direct[j];
                    } else {
                        norms_updated[j] *= std::sqrt(temp);
                    }
                }
            }
        }
        m_perm[0] = 0, m_perm[1] = 1;
        for (int k = 0; k < size; ++k) std::swap(m_perm[k], m_perm[transpositions[k]]);
    }

    double m_qr[2][2];
    double m_hcoeffs[2];
    int m_perm[2];
    int m_nonzero_pivots;
};

template <>
class Matrix<double, 2, 2> {
public:
    Matrix() : m_d {} {}
    double& operator()(int i, int j) { return m_d[i][j]; }
    const double& operator()(int i, int j) const { return m_d[i][j]; }
    ColPivHouseholderQR2 colPivHouseholderQr() const { return ColPivHouseholderQR2(m_d); }

    // ---- the members merge3_solver_common.cpp:96-160 (estimateRigidTransform) touches ----
    static Matrix Zero() { return Matrix(); }
    static Matrix Identity()
    {
        Matrix m;
        m.m_d[0][0] = 1.0, m.m_d[1][1] = 1.0;
        return m;
    }
    Matrix transpose() const
    {
        Matrix t;
        for (int i = 0; i < 2; ++i)
            for (int j = 0; j < 2; ++j) t.m_d[i][j] = m_d[j][i];
        return t;
    }
    double determinant() const { return m_d[0][0] * m_d[1][1] - m_d[1][0] * m_d[0][1]; }  // bruteforce_det2
    Matrix& operator+=(const Matrix& o)
    {
        for (int i = 0; i < 2; ++i)
            for (int j = 0; j < 2; ++j) m_d[i][j] += o.m_d[i][j];
        return *this;
    }
    // lazy coefficient product: (A B)(i,j) = A(i,0) B(0,j) + A(i,1) B(1,j)
    friend Matrix operator*(const Matrix& a, const Matrix& b)
    {
        Matrix r;
        for (int i = 0; i < 2; ++i)
            for (int j = 0; j < 2; ++j) r.m_d[i][j] = a.m_d[i][0] * b.m_d[0][j] + a.m_d[i][1] * b.m_d[1][j];
        return r;
    }
    friend Vector2d operator*(const Matrix& a, const Vector2d& v)
    {
        return Vector2d(a.m_d[0][0] * v(0) + a.m_d[0][1] * v(1), a.m_d[1][0] * v(0) + a.m_d[1][1] * v(1));
    }
    struct ColRef {
        Matrix* m;
        int j;
        ColRef& operator*=(double s)
        {
            m->m_d[0][j] *= s, m->m_d[1][j] *= s;
            return *this;
        }
    };
    ColRef col(int j) { return ColRef { this, j }; }

private:
    double m_d[2][2];
};

using Matrix2d = Matrix<double, 2, 2>;

// outer product u v^T (merge3_solver_common.cpp:135)
inline Matrix2d operator*(const Vector2d& u, const Vector2d::Transposed& vt)
{
    Matrix2d r;
    r(0, 0) = u(0) * vt.a, r(0, 1) = u(0) * vt.b;
    r(1, 0) = u(1) * vt.a, r(1, 1) = u(1) * vt.b;
    return r;
}

// Eigen::JacobiSVD<Matrix2d>(m, ComputeFullU | ComputeFullV), restated from the published
// algorithm (JacobiSVD.h: scaling by the largest |coefficient|, two-sided Jacobi sweeps on the
// (p,q) = (1,0) block with real_2x2_jacobi_svd + JacobiRotation::makeJacobi (Jacobi.h), sign fix of
// U, descending sort).  THIRD-PARTY ARITHMETIC, NOT PINNED: no real Eigen in this image.
enum { ComputeFullU = 0x04, ComputeFullV = 0x10 };

struct JacobiRotation2 {
    double c = 1.0, s = 0.0;
    JacobiRotation2 transpose() const { return JacobiRotation2 { c, -s }; }
    friend JacobiRotation2 operator*(const JacobiRotation2& a, const JacobiRotation2& b)
    {
        return JacobiRotation2 { a.c * b.c - a.s * b.s, a.c * b.s + a.s * b.c };
    }
    // makeJacobi(x, y, z) for the symmetric 2x2 [[x y],[y z]]
    bool makeJacobi(double x, double y, double z)
    {
        const double deno = 2.0 * std::abs(y);
        if (deno < DBL_MIN) {
            c = 1.0, s = 0.0;
            return false;
        }
        const double tau = (x - z) / deno;
        const double w = std::sqrt(tau * tau + 1.0);
        const double t = (tau > 0.0) ? 1.0 / (tau + w) : 1.0 / (tau - w);
        const double sign_t = t > 0.0 ? 1.0 : -1.0;
        const double n = 1.0 / std::sqrt(t * t + 1.0);
        s = -sign_t * (y / std::abs(y)) * std::abs(t) * n;
        c = n;
        return true;
    }
};

// apply_rotation_in_the_plane on two length-2 vectors given by accessors
inline void jacobi_apply(double& x0, double& y0, double& x1, double& y1, const JacobiRotation2& j)
{
    if (j.c == 1.0 && j.s == 0.0) return;
    const double a0 = x0, b0 = y0, a1 = x1, b1 = y1;
    x0 = j.c * a0 + j.s * b0, y0 = -j.s * a0 + j.c * b0;
    x1 = j.c * a1 + j.s * b1, y1 = -j.s * a1 + j.c * b1;
}

template <typename M>
class JacobiSVD;

template <>
class JacobiSVD<Matrix2d> {
public:
    JacobiSVD(const Matrix2d& m, unsigned /*options: full U and V*/) { compute(m); }
    const Matrix2d& matrixU() const { return m_u; }
    const Matrix2d& matrixV() const { return m_v; }
    double singularValue(int i) const { return m_sv[i]; }

private:
    // rows p,q of `a` <- J applied on the left
    static void onTheLeft(Matrix2d& a, int p, int q, const JacobiRotation2& j) { jacobi_apply(a(p, 0), a(q, 0), a(p, 1), a(q, 1), j); }
    // columns p,q of `a` <- J applied on the right (apply_rotation_in_the_plane with J^T)
    static void onTheRight(Matrix2d& a, int p, int q, const JacobiRotation2& j) { jacobi_apply(a(0, p), a(0, q), a(1, p), a(1, q), j.transpose()); }

    void compute(const Matrix2d& matrix)
    {
        using std::abs;
        const double considerAsZero = DBL_MIN, precision = 2.0 * DBL_EPSILON;
        double scale = 0.0;
        bool nan = false;
        for (int i = 0; i < 2; ++i)
            for (int j = 0; j < 2; ++j) {
                const double a = abs(matrix(i, j));
                if (a != a) nan = true;
                if (a > scale) scale = a;
            }
        m_u = Matrix2d::Identity(), m_v = Matrix2d::Identity();
        m_sv[0] = m_sv[1] = 0.0;
        if (nan || !std::isfinite(scale)) return;  // InvalidInput
        if (scale == 0.0) scale = 1.0;
        Matrix2d w;
        for (int i = 0; i < 2; ++i)
            for (int j = 0; j < 2; ++j) w(i, j) = matrix(i, j) / scale;
        double maxDiag = std::max(abs(w(0, 0)), abs(w(1, 1)));
        bool finished = false;
        while (!finished) {
            finished = true;
            const int p = 1, q = 0;
            const double threshold = std::max(considerAsZero, precision * maxDiag);
            if (abs(w(p, q)) > threshold || abs(w(q, p)) > threshold) {
                finished = false;
                // real_2x2_jacobi_svd(w, p, q, &j_left, &j_right)
                Matrix2d m;
                m(0, 0) = w(p, p), m(0, 1) = w(p, q), m(1, 0) = w(q, p), m(1, 1) = w(q, q);
                JacobiRotation2 rot1;
                const double t = m(0, 0) + m(1, 1);
                const double d = m(1, 0) - m(0, 1);
                if (abs(d) < DBL_MIN) {
                    rot1.s = 0.0, rot1.c = 1.0;
                } else {
                    const double u = t / d;
                    const double tmp = std::sqrt(1.0 + u * u);
                    rot1.s = 1.0 / tmp;
                    rot1.c = u / tmp;
                }
                onTheLeft(m, 0, 1, rot1);
                JacobiRotation2 j_right;
                j_right.makeJacobi(m(0, 0), m(0, 1), m(1, 1));
                const JacobiRotation2 j_left = rot1 * j_right.transpose();
                onTheLeft(w, p, q, j_left);
                onTheRight(m_u, p, q, j_left.transpose());
                onTheRight(w, p, q, j_right);
                onTheRight(m_v, p, q, j_right);
                maxDiag = std::max(maxDiag, std::max(abs(w(p, p)), abs(w(q, q))));
            }
        }
        for (int i = 0; i < 2; ++i) {
            const double a = w(i, i);
            m_sv[i] = abs(a);
            if (a < 0.0) m_u(0, i) = -m_u(0, i), m_u(1, i) = -m_u(1, i);
        }
        m_sv[0] *= scale, m_sv[1] *= scale;
        if (m_sv[1] > m_sv[0]) {  // descending order; maxCoeff keeps the first of equals
            std::swap(m_sv[0], m_sv[1]);
            std::swap(m_u(0, 0), m_u(0, 1)), std::swap(m_u(1, 0), m_u(1, 1));
            std::swap(m_v(0, 0), m_v(0, 1)), std::swap(m_v(1, 0), m_v(1, 1));
        }
    }

    Matrix2d m_u, m_v;
    double m_sv[2];
};

}  // namespace Eigen
