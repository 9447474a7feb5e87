// Pose value types of the bottom-up plan solver (reference:
// src/constraint_solver/src/solving/bottom_up/plan_pose_types.hpp:18-30): a cluster's pose is a
// map element id -> point position or line endpoints.
#pragma once

#include <unordered_map>
#include <variant>

#include <gcs/math/vector2d.hpp>
#include <gcs/model/gcs_data_structures.hpp>

namespace Gcs::Solvers::BottomUp {

struct PointPose {
    Eigen::Vector2d position;
};

struct LinePose {
    Eigen::Vector2d p1;
    Eigen::Vector2d p2;
};

using ElementPose = std::variant<PointPose, LinePose>;
using ClusterPose = std::unordered_map<ConstraintGraph::NodeIdType, ElementPose>;

}  // namespace Gcs::Solvers::BottomUp
