/*
 * gcs_oracle.c — CPU restatement of the reference's Newton-Raphson sub-system path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (2d_geometry_constraint_solver_b200/,
 * include/) may include, link or call this file; only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py use it, and only as the checker / the
 * reported CPU baseline.
 *
 * PARITY STATUS: "parity unpinned" at the third-party boundary.  The reference
 * (SolyomBalint/2D_geometry_constraint_solver, /root/reference) has no tests, golden vectors or
 * fixtures for this path (doc/milestones.md:8), and its arithmetic lives in two libraries that
 * are not vendored: autodiff v1.1.2 (CMakeLists.txt:22-34) and Eigen (conanfile.py:11).  This
 * file restates (a) the reference's own code literally, file:line cited at each function, and
 * (b) the published algorithms of autodiff's forward-mode `dual` expression evaluation and of
 * Eigen's ColPivHouseholderQR for a 2x2 system.  It is pinned against (1) closed-form analytic
 * roots (tests/test_oracle.py) and (2) the reference's own solve2D / primitives / heuristics /
 * solver sources compiled from /root/reference against small stand-in headers for the two
 * libraries (oracle/ref_shim, built into oracle/_ref/; tests/golden/ holds the vectors).
 *
 * All arithmetic: IEEE-754 binary64, one rounding per operation, no contraction.  Build with
 * -O2 -ffp-contract=off (the reference is built for baseline x86-64 without -march, so it never
 * fuses; conan_profiles/Linux/LinuxGccStd20ReleaseProfile).
 */
#include "gcs_oracle.h"

#include <float.h>
#include <math.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define TOL GCS_CONVERGENCE_THRESHOLD
#define MAXIT GCS_MAXIMUM_ITERATIONS

typedef struct {
    double x, y;
    int iters;
    int converged;
} seed_result;

/* three-valued sign used by every heuristic: heuristics.hpp:54, :102, :222, :258, :331 */
static int sgn3(double x) { return (x > 0) - (x < 0); }

/* ------------------------------------------------------------------------------------------
 * Eigen::Matrix2d::colPivHouseholderQr().solve(rhs)           (newton_raphson.hpp:80)
 *
 * Restated from Eigen's ColPivHouseholderQR<Matrix2d>::computeInPlace() and _solve_impl()
 * (Eigen/src/QR/ColPivHouseholderQR.h) with makeHouseholder / applyHouseholderOnTheLeft
 * (Eigen/src/Householder/Householder.h) and the column-major triangular_solve_vector
 * (Eigen/src/Core/products/TriangularSolverVector.h).  The matrix is column major:
 * J = [[a b],[c d]], column 0 = (a,c), column 1 = (b,d).
 * ------------------------------------------------------------------------------------------ */
static void colpiv_householder_qr_solve_2x2(
    double a, double b, double c, double d, double r0, double r1, double* s0, double* s1)
{
    const double eps = DBL_EPSILON;
    double qr00 = a, qr10 = c, qr01 = b, qr11 = d;
    double norm_direct[2], norm_updated[2];

    /* m_colNormsDirect(k) = m_qr.col(k).norm()  == sqrt(squaredNorm()) */
    norm_direct[0] = sqrt(qr00 * qr00 + qr10 * qr10);
    norm_direct[1] = sqrt(qr01 * qr01 + qr11 * qr11);
    norm_updated[0] = norm_direct[0];
    norm_updated[1] = norm_direct[1];

    /* threshold_helper = abs2(maxCoeff * epsilon) / rows ; maxCoeff keeps the first maximum */
    double maxn = norm_updated[0];
    if (norm_updated[1] > maxn) maxn = norm_updated[1];
    double th = maxn * eps;
    const double threshold_helper = (th * th) / 2.0;
    const double norm_downdate_threshold = sqrt(eps);

    int nonzero_pivots = 2; /* size */

    /* ---- k = 0 ---- */
    int big = (norm_updated[1] > norm_updated[0]) ? 1 : 0; /* maxCoeff(&index), first wins */
    double biggest_sq = norm_updated[big] * norm_updated[big];
    if (nonzero_pivots == 2 && biggest_sq < threshold_helper * 2.0 /* rows-k */) nonzero_pivots = 0;
    if (big != 0) {
        double t;
        t = qr00, qr00 = qr01, qr01 = t;
        t = qr10, qr10 = qr11, qr11 = t;
        t = norm_updated[0], norm_updated[0] = norm_updated[1], norm_updated[1] = t;
        t = norm_direct[0], norm_direct[0] = norm_direct[1], norm_direct[1] = t;
    }
    /* makeHouseholderInPlace on (qr00, qr10) */
    double tau0, beta0, v;
    {
        double tail_sq = qr10 * qr10;
        double c0 = qr00;
        if (tail_sq <= DBL_MIN) {
            tau0 = 0.0;
            beta0 = c0;
            v = 0.0;
        } else {
            beta0 = sqrt(c0 * c0 + tail_sq);
            if (c0 >= 0.0) beta0 = -beta0;
            v = qr10 / (c0 - beta0);
            tau0 = (beta0 - c0) / beta0;
        }
    }
    qr10 = v;
    qr00 = beta0;
    /* bottomRightCorner(2,1).applyHouseholderOnTheLeft(essential, tau0, workspace) */
    if (tau0 != 0.0) {
        double tmp = v * qr11;
        tmp += qr01;
        qr01 -= tau0 * tmp;
        qr11 -= tmp * (tau0 * v);
    }
    /* norm down-date of column 1 (LAPACK xGEQPF style) */
    if (norm_updated[1] != 0.0) {
        double temp = fabs(qr01) / norm_updated[1];
        temp = (1.0 + temp) * (1.0 - temp);
        temp = temp < 0.0 ? 0.0 : temp;
        double ratio = norm_updated[1] / norm_direct[1];
        double temp2 = temp * (ratio * ratio);
        if (temp2 <= norm_downdate_threshold) {
            norm_direct[1] = sqrt(qr11 * qr11); /* col(1).tail(1).norm() */
            norm_updated[1] = norm_direct[1];
        } else {
            norm_updated[1] *= sqrt(temp);
        }
    }
    /* ---- k = 1 ---- (1x1 tail: tau1 = 0, beta1 = qr11, nothing to apply) */
    {
        double sq = norm_updated[1] * norm_updated[1];
        if (nonzero_pivots == 2 && sq < threshold_helper * 1.0 /* rows-k */) nonzero_pivots = 1;
    }

    /* ---- _solve_impl ---- */
    if (nonzero_pivots == 0) {
        *s0 = 0.0;
        *s1 = 0.0;
        return;
    }
    double c0 = r0, c1 = r1;
    /* c.applyOnTheLeft(householderQ().setLength(nonzero_pivots).adjoint()) : H0 then (H1 = I) */
    if (tau0 != 0.0) {
        double tmp = v * c1;
        tmp += c0;
        c0 -= tau0 * tmp;
        c1 -= tmp * (tau0 * v);
    }
    if (nonzero_pivots == 2) c1 *= (1.0 - 0.0); /* rows()==1 branch of applyHouseholderOnTheLeft */
    /* topLeftCorner(nz,nz).triangularView<Upper>().solveInPlace(c.topRows(nz)), column major,
     * with the exact-zero skips of triangular_solve_vector */
    if (nonzero_pivots == 2) {
        if (c1 != 0.0) {
            c1 /= qr11;
            c0 -= c1 * qr01;
        }
        if (c0 != 0.0) c0 /= qr00;
    } else {
        if (c0 != 0.0) c0 /= qr00;
        c1 = 0.0; /* dst.row(perm(i)).setZero() for i >= nonzero_pivots */
    }
    /* dst.row(colsPermutation.indices()(i)) = c.row(i) ; indices = [big, 1-big] */
    if (big == 0) {
        *s0 = c0;
        *s1 = c1;
    } else {
        *s1 = c0;
        *s0 = c1;
    }
}

/* ------------------------------------------------------------------------------------------
 * Equation primitives (equation_primitives.hpp), evaluated in the order autodiff v1.1.2's
 * expression templates produce for a `dual`: `a - b` is `a + (-b)`;
 * `expr + number` becomes `number + expr`; assigning `l + r` evaluates r first, then adds l
 * (a nested sum adds its l then its r; adding `-e` subtracts e); assigning `l * r` evaluates r,
 * then multiplies by l; `pow(e, 2)` is aux = pow(e.val, 1); val = aux * val.
 * Each returns the value and writes the two partial derivatives.
 * ------------------------------------------------------------------------------------------ */

/* pointToPointDistance: pow(x - x0, 2) + pow(y - y0, 2) - pow(d, 2)   equation_primitives.hpp:26 */
static double eq_point_to_point(
    double x0, double y0, double d, double x, double y, double* dfx, double* dfy)
{
    double q = d * d;
    double dx = (-x0) + x;
    double dy = (-y0) + y;
    double f = ((-q) + dx * dx) + dy * dy;
    *dfx = 2.0 * dx;
    *dfy = 2.0 * dy;
    return f;
}

/* pointToLineDistance: (xb-xa)*(y-ya) - (yb-ya)*(x-xa) - d*lineLength  equation_primitives.hpp:74 */
static double eq_point_to_line(double xa, double ya, double xb, double yb, double d, double len,
    double x, double y, double* dfx, double* dfy)
{
    double ex = (-xa) + xb;
    double ey = (-ya) + yb;
    double ux = (-xa) + x;
    double uy = (-ya) + y;
    double f = ((-(len * d)) + uy * ex) - ux * ey;
    *dfx = -ey;
    *dfy = ex;
    return f;
}

/* lineNormalSignedDistanceDiff: nx*dX + ny*dY + s1 - s2            equation_primitives.hpp:181-182 */
static double eq_signed_distance_diff(
    double dX, double dY, double s1, double s2, double nx, double ny, double* dfx, double* dfy)
{
    double f = (((-s2) + dX * nx) + dY * ny) + s1;
    *dfx = dX;
    *dfy = dY;
    return f;
}

/* lineNormalAngleConstraint: -ny*fdx + nx*fdy - L*cosA             equation_primitives.hpp:146-147 */
static double eq_normal_angle(
    double fdx, double fdy, double len, double cosA, double nx, double ny, double* dfx, double* dfy)
{
    double f = ((-(cosA * len)) + fdx * (-ny)) + fdy * nx;
    *dfx = fdy;
    *dfy = -fdx;
    return f;
}

/* unitNormalConstraint: nx*nx + ny*ny - 1.0                        equation_primitives.hpp:198 */
static double eq_unit_normal(double nx, double ny, double* dfx, double* dfy)
{
    double f = ((ny * ny) + (nx * nx)) + (-1.0);
    *dfx = nx + nx;
    *dfy = ny + ny;
    return f;
}

/* ------------------------------------------------------------------------------------------
 * One system = kind + its constants.  eval() gives (f, g, J) at (x, y).
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    int kind;
    double k[12];
} system2;

static void eval_system(const system2* s, double x, double y, double* f, double* g, double J[4])
{
    const double* k = s->k;
    switch (s->kind) {
    case GCS_KIND_PP:
        *f = eq_point_to_point(k[0], k[1], k[2], x, y, &J[0], &J[1]);
        *g = eq_point_to_point(k[3], k[4], k[5], x, y, &J[2], &J[3]);
        break;
    case GCS_KIND_SDD: /* k: dX dY s1 s2 */
        *f = eq_signed_distance_diff(k[0], k[1], k[2], k[3], x, y, &J[0], &J[1]);
        *g = eq_unit_normal(x, y, &J[2], &J[3]);
        break;
    case GCS_KIND_PPL: /* k: px py r | xa ya xb yb s L */
        *f = eq_point_to_point(k[0], k[1], k[2], x, y, &J[0], &J[1]);
        *g = eq_point_to_line(k[3], k[4], k[5], k[6], k[7], k[8], x, y, &J[2], &J[3]);
        break;
    case GCS_KIND_PLL: /* k: xa ya xb yb s L | xa ya xb yb s L */
        *f = eq_point_to_line(k[0], k[1], k[2], k[3], k[4], k[5], x, y, &J[0], &J[1]);
        *g = eq_point_to_line(k[6], k[7], k[8], k[9], k[10], k[11], x, y, &J[2], &J[3]);
        break;
    case GCS_KIND_ANG: /* k: fdx fdy L cosA */
        *f = eq_normal_angle(k[0], k[1], k[2], k[3], x, y, &J[0], &J[1]);
        *g = eq_unit_normal(x, y, &J[2], &J[3]);
        break;
    default:
        *f = *g = NAN;
        J[0] = J[1] = J[2] = J[3] = NAN;
    }
}

/* ------------------------------------------------------------------------------------------
 * Equations::solve2D, one guess                                   newton_raphson.hpp:53-99
 *   vars = guess, prevVars = (0,0); for i < 1000: J, -f, -g, step = QR-solve; THEN the
 *   convergence test on |prev - vars| (strict <, both components); prev = vars; vars += step.
 * New observable outputs: iters = i at break (or 1000), converged = left via break.
 * ------------------------------------------------------------------------------------------ */
static seed_result newton2d(const system2* s, double gx, double gy)
{
    double x = gx, y = gy;
    double px = 0.0, py = 0.0;
    seed_result out;
    int i;
    out.converged = 0;
    for (i = 0; i < MAXIT; ++i) {
        double f, g, J[4], s0, s1;
        eval_system(s, x, y, &f, &g, J);
        colpiv_householder_qr_solve_2x2(J[0], J[1], J[2], J[3], -f, -g, &s0, &s1);
        if (fabs(px - x) < TOL && fabs(py - y) < TOL) {
            out.converged = 1;
            break;
        }
        px = x;
        py = y;
        x += s0;
        y += s1;
    }
    out.x = x;
    out.y = y;
    out.iters = i;
    return out;
}

/* The same run, watching how close every convergence decision came to the threshold.  For the
 * update k that led to iterate k+1 (Jacobian J_k), m_k = max(|dx_k|, |dy_k|) is what the reference
 * compares with TOL one iteration later (newton_raphson.hpp:83-88).  The contracted kernels
 * (csrc/newton_relaxed.cuh, G3) vouch for such a decision only if it is further from the threshold
 * than     first level : band  = 2^-36 S + 2^-16 TOL            (integer tests in the hot loop)
 *          second level: sdr / |det J_k| + 2^-40 TOL, sdr = 2^-44 S dr   (careful mode: every update)
 * Returns in out[0] the minimum over k of |m_k - TOL| - min(band, second)/2 and in out[1] the minimum
 * of |m_k - TOL| - second/2: how far the LITERAL trajectory stayed from every decision boundary in
 * excess of half those margins (half: the two arithmetics' own m_k differ by a fraction of the margin).
 * Negative: some decision of this run sat closer to the threshold than any arithmetic but the
 * literal one may decide.  (tests/test_gpu_margins.py) */
static void newton2d_decision_slack(const system2* s, double gx, double gy, double sdr, double band, double out[2])
{
    double x = gx, y = gy, px = 0.0, py = 0.0, det_prev = INFINITY;
    out[0] = out[1] = INFINITY;
    for (int i = 0; i < MAXIT; ++i) {
        double f, g, J[4], s0, s1;
        eval_system(s, x, y, &f, &g, J);
        colpiv_householder_qr_solve_2x2(J[0], J[1], J[2], J[3], -f, -g, &s0, &s1);
        if (i > 0) { /* i == 0 compares the seed itself with (0,0): the same exact numbers in any arithmetic */
            const double m = fmax(fabs(px - x), fabs(py - y));
            const double second = sdr / det_prev + 0x1p-40 * TOL;
            const double gap = fabs(m - TOL);
            const double a = gap - 0.5 * fmin(band, second), b2 = gap - 0.5 * second;
            if (!(a >= out[0])) out[0] = a; /* NaN sticks */
            if (!(b2 >= out[1])) out[1] = b2;
        }
        if (fabs(px - x) < TOL && fabs(py - y) < TOL) break;
        det_prev = fabs(J[0] * J[3] - J[1] * J[2]);
        px = x, py = y;
        x += s0, y += s1;
    }
}

int gcs_oracle_newton2d(int kind, const double* consts, double gx, double gy, double* x, double* y,
    int* iters, int* converged)
{
    system2 s;
    if (kind < 1 || kind > GCS_KIND_COUNT) return GCS_E_INVALID;
    s.kind = kind;
    memcpy(s.k, consts, sizeof(s.k));
    seed_result r = newton2d(&s, gx, gy);
    *x = r.x, *y = r.y, *iters = r.iters, *converged = r.converged;
    return GCS_OK;
}

void gcs_oracle_qr_solve_2x2(const double J[4], const double rhs[2], double step[2])
{
    colpiv_householder_qr_solve_2x2(J[0], J[1], J[2], J[3], rhs[0], rhs[1], &step[0], &step[1]);
}

/* ------------------------------------------------------------------------------------------
 * Heuristic helpers (heuristics.hpp)
 * ------------------------------------------------------------------------------------------ */

/* triangleOrientation                                              heuristics.hpp:22-27 */
static double triangle_orientation(
    double ax, double ay, double bx, double by, double cx, double cy)
{
    return ((bx - ax) * (cy - ay)) - ((by - ay) * (cx - ax));
}

/* perpendicularFoot                                                heuristics.hpp:144-150 */
static void perpendicular_foot(double px, double py, double l1x, double l1y, double l2x,
    double l2y, double* fx, double* fy)
{
    double dx = l2x - l1x, dy = l2y - l1y;
    double t = (dx * (px - l1x) + dy * (py - l1y)) / (dx * dx + dy * dy);
    *fx = l1x + t * dx;
    *fy = l1y + t * dy;
}

/* lineLineIntersection                                             heuristics.hpp:165-181 */
static int line_line_intersection(double l1ax, double l1ay, double l1bx, double l1by,
    double l2ax, double l2ay, double l2bx, double l2by, double* ix, double* iy)
{
    double d1x = l1bx - l1ax, d1y = l1by - l1ay;
    double d2x = l2bx - l2ax, d2y = l2by - l2ay;
    double cross = d1x * d2y - d1y * d2x;
    if (fabs(cross) < GCS_PARALLEL_EPSILON) return 0;
    double dlx = l2ax - l1ax, dly = l2ay - l1ay;
    double t = (dlx * d2y - dly * d2x) / cross;
    *ix = l1ax + t * d1x;
    *iy = l1ay + t * d1y;
    return 1;
}

/* reconstructLineEndpoints         point_line_solvers.cpp:74-106 == line_angle_solvers.cpp:128-161 */
static void reconstruct_line_endpoints(double c1x, double c1y, double c2x, double c2y, double nx,
    double ny, double p, double canvas_len, double out[4])
{
    double sd1 = (nx * c1x + ny * c1y) - p;
    double pr1x = c1x - sd1 * nx, pr1y = c1y - sd1 * ny;
    double sd2 = (nx * c2x + ny * c2y) - p;
    double pr2x = c2x - sd2 * nx, pr2y = c2y - sd2 * ny;
    double dirx = -ny, diry = nx;
    double midx = (pr1x + pr2x) / 2.0, midy = (pr1y + pr2y) / 2.0;
    double span = fabs(dirx * (pr2x - pr1x) + diry * (pr2y - pr1y));
    double mx = (canvas_len < span) ? span : canvas_len; /* std::max(canvas_len, span) */
    double half = mx / 2.0;
    out[0] = midx - half * dirx;
    out[1] = midy - half * diry;
    out[2] = midx + half * dirx;
    out[3] = midy + half * diry;
}

/* default seeds: 0,1 = newton_raphson.hpp:105-107; 2..7 = the multi-start extension
 * (SURVEY.md section 8a): the other two diagonal corners, then the four axis points at radius
 * 20000*sqrt(2). */
static const double k_default_seeds[GCS_MAX_SEEDS][2] = {
    { 20000.0, 20000.0 },
    { -20000.0, -20000.0 },
    { 20000.0, -20000.0 },
    { -20000.0, 20000.0 },
    /* multi-start extension: radius ~20000*sqrt(2) at 22.5 deg + k*90 deg.  Off the axes on
     * purpose: an anchored triangle has both fixed points on y = 0, where the distance-distance
     * Jacobian is singular, and a seed on that line crawls for hundreds of iterations */
    { 26131.0, 10824.0 },
    { -10824.0, 26131.0 },
    { -26131.0, -10824.0 },
    { 10824.0, -26131.0 },
};

/* nearest-to-canvas among the candidates; with two candidates this is
 * `(dist0 <= dist1) ? candidate0 : candidate1` (heuristics.hpp:214-216). */
static int pick_nearest(const seed_result* c, int ns, double cfx, double cfy)
{
    int best = 0;
    double bd = 0.0;
    for (int k = 0; k < ns; ++k) {
        double dx = c[k].x - cfx, dy = c[k].y - cfy;
        double dd = dx * dx + dy * dy;
        if (k == 0) {
            bd = dd;
        } else if (!(bd <= dd)) {
            bd = dd;
            best = k;
        }
    }
    return best;
}

/* pickByTriangleOrientation (heuristics.hpp:46-57), generalised to ns candidates: first whose
 * orientation sign equals the canvas sign, else the LAST one unchecked (for ns = 2 exactly the
 * reference: candidate1 is never tested). */
static int pick_by_orientation(const seed_result* c, int ns, int canvas_sign, double ax,
    double ay, double bx, double by)
{
    for (int k = 0; k + 1 < ns; ++k) {
        double ori = triangle_orientation(ax, ay, bx, by, c[k].x, c[k].y);
        if (canvas_sign == sgn3(ori)) return k;
    }
    return ns - 1;
}

/* one full sub-system: solve2D over all seeds + selection (+ reconstruction) */
/* gcs_oracle_decision_slack: where solve_one leaves each run's slack (NULL: not tracing) */
static const double* g_trace_sdr = NULL;
static const double* g_trace_band = NULL;
static double* g_trace_slack = NULL; /* [2][n_seeds][n] */

static void solve_one(const gcs_b200_batch* b, int64_t i)
{
    const int ns = b->n_seeds;
    const int64_t n = b->n;
    const uint8_t code = b->code ? b->code[i] : (uint8_t)GCS_MAKE_CODE(0, 0, 0);
    const int sign0 = GCS_CODE_SIGN0(code);
    seed_result cand[GCS_MAX_SEEDS];
    system2 s;
    double g[GCS_MAX_SEEDS][2];
    double out[4] = { 0, 0, 0, 0 };
    int root = 0;
    double in[GCS_MAX_IN_COLS];
    const int nin = gcs_oracle_kind_in_cols(b->kind);
    for (int c = 0; c < nin; ++c) in[c] = b->in[c][i];

    s.kind = b->kind;
    memset(s.k, 0, sizeof(s.k));
    for (int k = 0; k < ns; ++k) {
        g[k][0] = k_default_seeds[k][0];
        g[k][1] = k_default_seeds[k][1];
    }

    switch (b->kind) {
    case GCS_KIND_PP:
        memcpy(s.k, in, 6 * sizeof(double));
        break;
    case GCS_KIND_SDD: {
        /* delta = P2.position - P1.position        point_line_solvers.cpp:205, :349 */
        s.k[0] = in[2] - in[0];
        s.k[1] = in[3] - in[1];
        s.k[2] = in[4];
        s.k[3] = in[5];
        /* guesses { canvasNormal, -canvasNormal }   point_line_solvers.cpp:218-219 */
        g[0][0] = in[6], g[0][1] = in[7];
        g[1][0] = -in[6], g[1][1] = -in[7];
        break;
    }
    case GCS_KIND_PPL: {
        memcpy(s.k, in, 8 * sizeof(double));
        /* fixedLineLength = Line::length() = (p2 - p1).norm()   point_line_solvers.cpp:506,
         * src/model/elements.cpp:118-121 */
        double ex = in[5] - in[3], ey = in[6] - in[4];
        s.k[8] = sqrt(ex * ex + ey * ey);
        break;
    }
    case GCS_KIND_PLL: {
        for (int l = 0; l < 2; ++l) {
            const double* q = in + 5 * l;
            double ex = q[2] - q[0], ey = q[3] - q[1];
            s.k[6 * l + 0] = q[0], s.k[6 * l + 1] = q[1], s.k[6 * l + 2] = q[2];
            s.k[6 * l + 3] = q[3], s.k[6 * l + 4] = q[4];
            s.k[6 * l + 5] = sqrt(ex * ex + ey * ey); /* point_line_solvers.cpp:636, :642 */
        }
        break;
    }
    case GCS_KIND_ANG: {
        s.k[0] = in[0], s.k[1] = in[1];
        s.k[2] = sqrt(in[0] * in[0] + in[1] * in[1]); /* fixedLineDirection.norm() line_angle_solvers.cpp:479 */
        s.k[3] = in[2];
        g[0][0] = in[3], g[0][1] = in[4];
        g[1][0] = -in[3], g[1][1] = -in[4];
        break;
    }
    }
    if (b->guesses) {
        for (int k = 0; k < ns; ++k) {
            g[k][0] = b->guesses[((int64_t)k * 2 + 0) * n + i];
            g[k][1] = b->guesses[((int64_t)k * 2 + 1) * n + i];
        }
    }

    for (int k = 0; k < ns; ++k) cand[k] = newton2d(&s, g[k][0], g[k][1]);
    if (g_trace_slack)
        for (int k = 0; k < ns; ++k) {
            double sl[2];
            newton2d_decision_slack(&s, g[k][0], g[k][1], g_trace_sdr[i], g_trace_band[i], sl);
            g_trace_slack[((int64_t)0 * ns + k) * n + i] = sl[0];
            g_trace_slack[((int64_t)1 * ns + k) * n + i] = sl[1];
        }

    switch (b->kind) {
    case GCS_KIND_PP:
        /* point_point_solvers.cpp:68-71, :148-151 */
        root = pick_by_orientation(cand, ns, sign0, in[0], in[1], in[3], in[4]);
        out[0] = cand[root].x, out[1] = cand[root].y;
        break;
    case GCS_KIND_PPL: {
        /* point_line_solvers.cpp:520-529 ; heuristics.hpp:203-224 */
        if (code & GCS_CODE_COLLINEAR) {
            root = pick_nearest(cand, ns, in[8], in[9]);
        } else {
            double fx, fy;
            perpendicular_foot(in[0], in[1], in[3], in[4], in[5], in[6], &fx, &fy);
            root = pick_by_orientation(cand, ns, sign0, in[0], in[1], fx, fy);
        }
        out[0] = cand[root].x, out[1] = cand[root].y;
        break;
    }
    case GCS_KIND_PLL: {
        /* point_line_solvers.cpp:656-682 */
        double ix, iy;
        int has = line_line_intersection(
            in[0], in[1], in[2], in[3], in[5], in[6], in[7], in[8], &ix, &iy);
        if (has && !(code & GCS_CODE_CANVAS_PARALLEL)) {
            if (code & GCS_CODE_COLLINEAR) {
                root = pick_nearest(cand, ns, in[10], in[11]);
            } else {
                /* solverRefPoint = intersection + line1.unitDirection()  (elements.cpp:103-106:
                 * (p2 - p1).normalized(): z = squaredNorm; z > 0 ? d / sqrt(z) : d) */
                double dx = in[2] - in[0], dy = in[3] - in[1];
                double z = dx * dx + dy * dy;
                if (z > 0.0) {
                    double nz = sqrt(z);
                    dx = dx / nz;
                    dy = dy / nz;
                }
                root = pick_by_orientation(cand, ns, sign0, ix, iy, ix + dx, iy + dy);
            }
        } else {
            root = pick_nearest(cand, ns, in[10], in[11]);
        }
        out[0] = cand[root].x, out[1] = cand[root].y;
        break;
    }
    case GCS_KIND_SDD: {
        /* point_line_solvers.cpp:226-246 ; heuristics.hpp:250-277 */
        const int sign1 = GCS_CODE_SIGN1(code);
        double p0 = (cand[0].x * in[0] + cand[0].y * in[1]) - in[4];
        double p1 = (cand[1].x * in[0] + cand[1].y * in[1]) - in[4];
        double d1 = (cand[0].x * in[0] + cand[0].y * in[1]) - p0;
        double d2 = (cand[0].x * in[2] + cand[0].y * in[3]) - p0;
        double nx, ny, p;
        if (sgn3(d1) == sign0 && sgn3(d2) == sign1) {
            root = 0, nx = cand[0].x, ny = cand[0].y, p = p0;
        } else {
            root = 1, nx = cand[1].x, ny = cand[1].y, p = p1;
        }
        reconstruct_line_endpoints(in[0], in[1], in[2], in[3], nx, ny, p, in[8], out);
        break;
    }
    case GCS_KIND_ANG: {
        /* line_angle_solvers.cpp:319-361, :509-557 ; heuristics.hpp:303-335 */
        double fdirx = -cand[0].y, fdiry = cand[0].x; /* candidate 0's free direction */
        double cross0 = (in[5] * fdiry) - (in[6] * fdirx);
        root = (sign0 == sgn3(cross0)) ? 0 : 1;
        double nx = cand[root].x, ny = cand[root].y;
        double p = (nx * in[7] + ny * in[8]) - in[9];
        reconstruct_line_endpoints(in[7], in[8], in[10], in[11], nx, ny, p, in[12], out);
        break;
    }
    }

    const int nout = gcs_oracle_kind_out_cols(b->kind);
    for (int c = 0; c < nout; ++c)
        if (b->out[c]) b->out[c][i] = out[c];
    for (int k = 0; k < ns; ++k) {
        if (b->cand) {
            b->cand[((int64_t)k * 2 + 0) * n + i] = cand[k].x;
            b->cand[((int64_t)k * 2 + 1) * n + i] = cand[k].y;
        }
        if (b->iters) b->iters[(int64_t)k * n + i] = (int16_t)cand[k].iters;
        if (b->converged) b->converged[(int64_t)k * n + i] = (uint8_t)cand[k].converged;
    }
    if (b->root_index) b->root_index[i] = (uint8_t)root;
}

int gcs_oracle_kind_in_cols(int kind)
{
    static const int t[GCS_KIND_COUNT + 1] = { 0, 6, 9, 10, 12, 13 };
    return (kind >= 1 && kind <= GCS_KIND_COUNT) ? t[kind] : 0;
}

int gcs_oracle_kind_out_cols(int kind)
{
    static const int t[GCS_KIND_COUNT + 1] = { 0, 2, 4, 2, 2, 4 };
    return (kind >= 1 && kind <= GCS_KIND_COUNT) ? t[kind] : 0;
}

int gcs_oracle_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

int gcs_oracle_solve(const gcs_b200_batch* b, int threads)
{
    if (!b || b->kind < 1 || b->kind > GCS_KIND_COUNT || b->n < 0) return GCS_E_INVALID;
    if (b->n_seeds < 2 || b->n_seeds > GCS_MAX_SEEDS) return GCS_E_INVALID;
    if ((b->kind == GCS_KIND_SDD || b->kind == GCS_KIND_ANG) && b->n_seeds != 2) return GCS_E_INVALID;
    if (b->mem != GCS_MEM_HOST) return GCS_E_INVALID;
    const int nin = gcs_oracle_kind_in_cols(b->kind);
    for (int c = 0; c < nin; ++c)
        if (!b->in[c] && b->n > 0) return GCS_E_INVALID;
    (void)threads;
#ifdef _OPENMP
    if (threads < 1) threads = omp_get_max_threads();
#pragma omp parallel for schedule(static) num_threads(threads)
#endif
    for (int64_t i = 0; i < b->n; ++i) solve_one(b, i);
    return GCS_OK;
}

int gcs_oracle_decision_slack(const gcs_b200_batch* b, const double* sdr, const double* band, double* slack, int threads)
{
    if (!sdr || !band || !slack) return GCS_E_INVALID;
    g_trace_sdr = sdr, g_trace_band = band, g_trace_slack = slack;
    const int rc = gcs_oracle_solve(b, threads);
    g_trace_sdr = g_trace_band = NULL, g_trace_slack = NULL;
    return rc;
}
