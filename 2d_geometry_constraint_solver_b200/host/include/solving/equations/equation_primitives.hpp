// Equation primitives with the reference's factory names and argument order (reference:
// src/constraint_solver/src/solving/equations/equation_primitives.hpp:23-199).
//
// The reference returns autodiff lambdas; here each factory returns a small tagged value that
// names the primitive and carries its constants, so that Equations::solve2D(f, g [, guesses])
// (newton_raphson.hpp) can hand the pair to the CUDA kernel specialised on it.  The residual
// formulas and their evaluation order live in csrc/newton_core.cuh (Sys<KIND>::eval).
// pointOnLine / lineToLineAngle exist in the reference but have no caller and no kernel here.
#pragma once

namespace Gcs::Equations {

struct PointToPointDistanceEq { double x0, y0, d; };                 // (x-x0)^2 + (y-y0)^2 - d^2      :23-28
struct PointToLineDistanceEq { double xa, ya, xb, yb, d, length; };  // (xb-xa)(y-ya) - (yb-ya)(x-xa) - d L   :70-76
struct LineNormalAngleEq { double fdx, fdy, length, cosAngle; };     // -ny fdx + nx fdy - L cosA      :141-149
struct LineNormalSignedDistanceDiffEq { double dx, dy, s1, s2; };    // nx dx + ny dy + s1 - s2        :176-184
struct UnitNormalEq { };                                             // nx^2 + ny^2 - 1                :196-199

inline PointToPointDistanceEq pointToPointDistance(double x0, double y0, double d) { return { x0, y0, d }; }
inline PointToLineDistanceEq pointToLineDistance(double xa, double ya, double xb, double yb, double d, double lineLength)
{
    return { xa, ya, xb, yb, d, lineLength };
}
inline LineNormalAngleEq lineNormalAngleConstraint(double fixedDirectionX, double fixedDirectionY, double fixedLineLength, double cosAngle)
{
    return { fixedDirectionX, fixedDirectionY, fixedLineLength, cosAngle };
}
inline LineNormalSignedDistanceDiffEq lineNormalSignedDistanceDiff(double deltaX, double deltaY, double signedDistanceToPoint1, double signedDistanceToPoint2)
{
    return { deltaX, deltaY, signedDistanceToPoint1, signedDistanceToPoint2 };
}
inline UnitNormalEq unitNormalConstraint() { return {}; }

}  // namespace Gcs::Equations
