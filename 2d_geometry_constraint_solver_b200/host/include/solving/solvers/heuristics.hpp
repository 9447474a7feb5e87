// Canvas-side geometric helpers of the root-selection heuristics (reference:
// src/solving/solvers/heuristics.hpp:22-181).  The packer uses them to reduce the canvas layout
// to the orientation code of a batch row; the candidate-side halves of the heuristics
// (pickByTriangleOrientation[WithFallback], pickLineBySignedDistances,
// pickLineNormalByAngleOrientation, heuristics.hpp:46-57, :203-335) run inside the CUDA kernels.
#pragma once

#include <cmath>
#include <optional>

#include <gcs/math/vector2d.hpp>

namespace Gcs::Solvers {

// signed area of triangle ABC: > 0 counter-clockwise            heuristics.hpp:22-27
inline double triangleOrientation(const Eigen::Vector2d& a, const Eigen::Vector2d& b, const Eigen::Vector2d& c)
{
    return ((b.x() - a.x()) * (c.y() - a.y())) - ((b.y() - a.y()) * (c.x() - a.x()));
}

// (lineP2 - lineP1) x (point - lineP1) / |lineP2 - lineP1|     heuristics.hpp:113-125
inline double signedDistanceToLine(const Eigen::Vector2d& point, const Eigen::Vector2d& lineP1, const Eigen::Vector2d& lineP2)
{
    const Eigen::Vector2d dir = lineP2 - lineP1;
    const double len = dir.norm();
    const Eigen::Vector2d rel = point - lineP1;
    const double cross = (dir.x() * rel.y()) - (dir.y() * rel.x());
    return cross / len;
}

// foot of the perpendicular from `point` on the infinite line  heuristics.hpp:144-150
inline Eigen::Vector2d perpendicularFoot(const Eigen::Vector2d& point, const Eigen::Vector2d& lineP1, const Eigen::Vector2d& lineP2)
{
    const Eigen::Vector2d dir = lineP2 - lineP1;
    const double t = dir.dot(point - lineP1) / dir.squaredNorm();
    return lineP1 + t * dir;
}

// intersection of two infinite lines, nullopt when |cross| < 1e-10   heuristics.hpp:165-181
inline std::optional<Eigen::Vector2d> lineLineIntersection(const Eigen::Vector2d& l1p1, const Eigen::Vector2d& l1p2,
    const Eigen::Vector2d& l2p1, const Eigen::Vector2d& l2p2)
{
    const Eigen::Vector2d d1 = l1p2 - l1p1;
    const Eigen::Vector2d d2 = l2p2 - l2p1;
    const double cross = d1.x() * d2.y() - d1.y() * d2.x();
    constexpr double PARALLEL_EPSILON = 1e-10;
    if (std::abs(cross) < PARALLEL_EPSILON) return std::nullopt;
    const Eigen::Vector2d delta = l2p1 - l1p1;
    const double t = (delta.x() * d2.y() - delta.y() * d2.x()) / cross;
    return l1p1 + t * d1;
}

}  // namespace Gcs::Solvers
