// Full SVD of a 2x2 matrix, shared by the Procrustes fits of the bottom-up Merge3 helpers
// (reference: merge3_solver_common.cpp:139-140) and of the solver->canvas transform
// (reference: gui/src/constraint_model.cpp:467-470), both of which construct
// Eigen::JacobiSVD<Matrix2d>(m, ComputeFullU | ComputeFullV).
#include <gcs/math/svd2x2.hpp>

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <utility>

using Eigen::Matrix2d;

namespace Gcs::Math {

namespace {

// Plane rotation (c, s) as Eigen's JacobiRotation stores it.
struct Rot {
    double c = 1.0, s = 0.0;
    Rot transposed() const { return { c, -s }; }
};
Rot compose(const Rot& a, const Rot& b) { return { a.c * b.c - a.s * b.s, a.c * b.s + a.s * b.c }; }

// (x, y) <- (c x + s y, -s x + c y) on two coefficient pairs; the identity rotation is skipped
void rotatePairs(double& x0, double& y0, double& x1, double& y1, const Rot& j)
{
    if (j.c == 1.0 && j.s == 0.0) return;
    const double a0 = x0, b0 = y0, a1 = x1, b1 = y1;
    x0 = j.c * a0 + j.s * b0, y0 = -j.s * a0 + j.c * b0;
    x1 = j.c * a1 + j.s * b1, y1 = -j.s * a1 + j.c * b1;
}
void rotateRows(Matrix2d& m, int p, int q, const Rot& j) { rotatePairs(m(p, 0), m(q, 0), m(p, 1), m(q, 1), j); }
void rotateCols(Matrix2d& m, int p, int q, const Rot& j) { rotatePairs(m(0, p), m(0, q), m(1, p), m(1, q), j.transposed()); }

// Jacobi rotation that diagonalises the symmetric [[x y],[y z]]
Rot symmetricJacobi(double x, double y, double z)
{
    const double deno = 2.0 * std::abs(y);
    if (deno < DBL_MIN) return {};
    const double tau = (x - z) / deno;
    const double w = std::sqrt(tau * tau + 1.0);
    const double t = (tau > 0.0) ? 1.0 / (tau + w) : 1.0 / (tau - w);
    const double n = 1.0 / std::sqrt(t * t + 1.0);
    const double signT = t > 0.0 ? 1.0 : -1.0;
    return { n, -signT * (y / std::abs(y)) * std::abs(t) * n };
}

// Full SVD of a 2x2 by two-sided Jacobi rotations, the way Eigen::JacobiSVD<Matrix2d> with
// ComputeFullU | ComputeFullV proceeds (merge3_solver_common.cpp:139-140 constructs exactly that):
// scale by the largest |coefficient|, sweep the (1,0) block until both off-diagonals are below
// 2 eps * max|diag|, flip U's column where the diagonal came out negative, sort descending.
// Third-party algorithm restated; the image has no Eigen to pin it against.
}  // namespace

void jacobiSvd2x2(const Matrix2d& a, Matrix2d& u, Matrix2d& v)
{
    u = Matrix2d::Identity(), v = Matrix2d::Identity();
    double scale = 0.0;
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j) {
            const double m = std::abs(a(i, j));
            if (m != m) return;
            scale = std::max(scale, m);
        }
    if (!std::isfinite(scale)) return;
    if (scale == 0.0) scale = 1.0;
    Matrix2d w;
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j) w(i, j) = a(i, j) / scale;
    double maxDiag = std::max(std::abs(w(0, 0)), std::abs(w(1, 1)));
    constexpr int p = 1, q = 0;
    for (;;) {
        const double threshold = std::max(DBL_MIN, (2.0 * DBL_EPSILON) * maxDiag);
        if (!(std::abs(w(p, q)) > threshold || std::abs(w(q, p)) > threshold)) break;
        Matrix2d m;
        m(0, 0) = w(p, p), m(0, 1) = w(p, q), m(1, 0) = w(q, p), m(1, 1) = w(q, q);
        Rot first;  // makes the block symmetric
        const double trace = m(0, 0) + m(1, 1), skew = m(1, 0) - m(0, 1);
        if (!(std::abs(skew) < DBL_MIN)) {
            const double r = trace / skew;
            const double h = std::sqrt(1.0 + r * r);
            first.s = 1.0 / h;
            first.c = r / h;
        }
        rotateRows(m, 0, 1, first);
        const Rot right = symmetricJacobi(m(0, 0), m(0, 1), m(1, 1));
        const Rot left = compose(first, right.transposed());
        rotateRows(w, p, q, left);
        rotateCols(u, p, q, left.transposed());
        rotateCols(w, p, q, right);
        rotateCols(v, p, q, right);
        maxDiag = std::max(maxDiag, std::max(std::abs(w(p, p)), std::abs(w(q, q))));
    }
    double sv[2];
    for (int i = 0; i < 2; ++i) {
        sv[i] = std::abs(w(i, i)) * scale;
        if (w(i, i) < 0.0) u(0, i) = -u(0, i), u(1, i) = -u(1, i);
    }
    if (sv[1] > sv[0]) {
        std::swap(u(0, 0), u(0, 1)), std::swap(u(1, 0), u(1, 1));
        std::swap(v(0, 0), v(0, 1)), std::swap(v(1, 0), v(1, 1));
    }
}


}  // namespace Gcs::Math
