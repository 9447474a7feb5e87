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
#include <cmath>
#include <initializer_list>

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

private:
    double m_d[2][2];
};

using Matrix2d = Matrix<double, 2, 2>;

}  // namespace Eigen
