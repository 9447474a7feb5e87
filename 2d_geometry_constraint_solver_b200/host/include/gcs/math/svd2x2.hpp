// Two-sided Jacobi SVD of a 2x2 matrix: a = u * diag(s) * v^T with s descending (s itself is not
// returned; both callers only need the rotations).  Restates the published algorithm of
// Eigen::JacobiSVD<Matrix2d> (third-party; unpinned in this image - see DESIGN.md section 4).
#pragma once

#include <gcs/export.hpp>
#include <gcs/math/matrix2d.hpp>

namespace Gcs::Math {

GCS_API void jacobiSvd2x2(const Eigen::Matrix2d& a, Eigen::Matrix2d& u, Eigen::Matrix2d& v);

}  // namespace Gcs::Math
