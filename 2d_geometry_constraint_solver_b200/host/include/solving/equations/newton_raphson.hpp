// Equations::solve2D with the reference's signature and constants (reference:
// src/constraint_solver/src/solving/equations/newton_raphson.hpp:17-20, :41-127).
//
// Same contract: two equations, two initial guesses, two candidates back, no status.  The
// iteration itself (Jacobian, 2x2 column-pivoted Householder QR step, the |prev - vars| < 1e-5
// test, the 1000-iteration cap) runs in the sm_100a kernel for the pair's kind, as a batch of
// one; there is no host implementation.  Pairs the reference never forms do not compile.
#pragma once

#include <array>
#include <cstddef>
#include <type_traits>

#include <gcs/export.hpp>
#include <gcs/math/vector2d.hpp>

#include "solving/equations/equation_primitives.hpp"

namespace Gcs::Equations {

static constexpr double CONVERGENCE_THRESHOLD = 0.00001;  // newton_raphson.hpp:17
static constexpr int MAXIMUM_ITERATIONS = 1000;           // newton_raphson.hpp:20

// newton_raphson.hpp:105-107 (the doc comment there says 2000; the code says 20000)
inline const std::array<Eigen::Vector2d, 2> DEFAULT_SPATIAL_GUESSES { Eigen::Vector2d { 20000.0, 20000.0 },
    Eigen::Vector2d { -20000.0, -20000.0 } };

// per-seed outcome the reference computes but does not return
struct Solve2DInfo {
    int iterations[2] = { 0, 0 };
    bool converged[2] = { false, false };
};

namespace detail {
    // One solve2D-equivalent on the device: `cols` are the kind's input columns (gcs_b200.h).
    // Throws std::runtime_error when the CUDA library cannot run.
    GCS_API std::array<Eigen::Vector2d, 2> solveOnDevice(int kind, const double* cols,
        const std::array<Eigen::Vector2d, 2>& guesses, Solve2DInfo* info, int device);
    // the line length as the kernel recomputes it: sqrt(ex*ex + ey*ey)
    GCS_API void requireLength(double ex, double ey, double length, const char* what);
}  // namespace detail

inline std::array<Eigen::Vector2d, 2> solve2D(const PointToPointDistanceEq& f, const PointToPointDistanceEq& g,
    const std::array<Eigen::Vector2d, 2>& initialGuesses, Solve2DInfo* info = nullptr, int device = 0)
{
    const double c[6] = { f.x0, f.y0, f.d, g.x0, g.y0, g.d };
    return detail::solveOnDevice(1, c, initialGuesses, info, device);
}

inline std::array<Eigen::Vector2d, 2> solve2D(const LineNormalSignedDistanceDiffEq& f, const UnitNormalEq&,
    const std::array<Eigen::Vector2d, 2>& initialGuesses, Solve2DInfo* info = nullptr, int device = 0)
{
    // the kernel forms delta = P2 - P1: with P1 = (0,0) that is (dx - 0, dy - 0), exact
    const double c[9] = { 0.0, 0.0, f.dx, f.dy, f.s1, f.s2, initialGuesses[0].x(), initialGuesses[0].y(), 0.0 };
    return detail::solveOnDevice(2, c, initialGuesses, info, device);
}

inline std::array<Eigen::Vector2d, 2> solve2D(const PointToPointDistanceEq& f, const PointToLineDistanceEq& g,
    const std::array<Eigen::Vector2d, 2>& initialGuesses, Solve2DInfo* info = nullptr, int device = 0)
{
    detail::requireLength(g.xb - g.xa, g.yb - g.ya, g.length, "pointToLineDistance");
    const double c[10] = { f.x0, f.y0, f.d, g.xa, g.ya, g.xb, g.yb, g.d, 0.0, 0.0 };
    return detail::solveOnDevice(3, c, initialGuesses, info, device);
}

inline std::array<Eigen::Vector2d, 2> solve2D(const PointToLineDistanceEq& f, const PointToLineDistanceEq& g,
    const std::array<Eigen::Vector2d, 2>& initialGuesses, Solve2DInfo* info = nullptr, int device = 0)
{
    detail::requireLength(f.xb - f.xa, f.yb - f.ya, f.length, "pointToLineDistance");
    detail::requireLength(g.xb - g.xa, g.yb - g.ya, g.length, "pointToLineDistance");
    const double c[12] = { f.xa, f.ya, f.xb, f.yb, f.d, g.xa, g.ya, g.xb, g.yb, g.d, 0.0, 0.0 };
    return detail::solveOnDevice(4, c, initialGuesses, info, device);
}

inline std::array<Eigen::Vector2d, 2> solve2D(const LineNormalAngleEq& f, const UnitNormalEq&,
    const std::array<Eigen::Vector2d, 2>& initialGuesses, Solve2DInfo* info = nullptr, int device = 0)
{
    detail::requireLength(f.fdx, f.fdy, f.length, "lineNormalAngleConstraint");
    const double c[13] = { f.fdx, f.fdy, f.cosAngle, initialGuesses[0].x(), initialGuesses[0].y(), 1.0, 0.0, 0.0, 0.0,
        0.0, 0.0, 0.0, 0.0 };
    return detail::solveOnDevice(5, c, initialGuesses, info, device);
}

// convenience overload with the default spatial guesses (newton_raphson.hpp:122-127)
template <typename FuncF, typename FuncG>
std::array<Eigen::Vector2d, 2> solve2D(FuncF&& f, FuncG&& g)
{
    return solve2D(std::forward<FuncF>(f), std::forward<FuncG>(g), DEFAULT_SPATIAL_GUESSES);
}

}  // namespace Gcs::Equations
