// The step after the solve: move the solved sketch back over the user's drawing (reference:
// gui/src/constraint_model.cpp:394-501, ConstraintModel::applySolverToCanvasTransform, called at
// the end of ConstraintModel::solveConstraintSystem :362-382).  SURVEY.md section 8f rank 4.
//
// The solver works in its own frame (the first leaf is anchored at the origin); the least-squares
// rigid motion solver -> canvas over all solved POINTS (2-D Procrustes: centroids, 2x2
// cross-covariance, SVD, R = V diag(1, det(V U^T)) U^T, t = c_canvas - R c_solver) is applied to
// every solved point and line, overwriting canvasPosition / canvasP1 / canvasP2.  One solved point:
// translation only.  None: nothing happens.  A reduction over the sketch followed by an
// element-wise map, on host objects: host code (a 100k-point sketch is 3 MB; the copies alone would
// cost more than the arithmetic).
#pragma once

#include <gcs/export.hpp>
#include <gcs/model/gcs_data_structures.hpp>

namespace Gcs::B200 {

GCS_API void applySolverToCanvasTransform(ConstraintGraph& graph);

}  // namespace Gcs::B200
