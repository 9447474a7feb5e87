// What a solved cluster looks like to the bottom-up plan solver (reference:
// src/constraint_solver/src/solving/bottom_up/plan_pose_types.hpp:18-30): every element of the
// cluster mapped to where it sits in the cluster's own frame - a position for a point, two
// endpoints for a line.  Plain aggregates with the reference's member names (its code initialises
// them with designated initialisers); the map type and the hash of its key are the reference's too,
// so that iteration order - and with it the order of every floating-point sum over a cluster -
// is the same.
#pragma once

#include <unordered_map>
#include <variant>

#include <gcs/math/vector2d.hpp>
#include <gcs/model/gcs_data_structures.hpp>

namespace Gcs::Solvers::BottomUp {

struct LinePose {
    Eigen::Vector2d p1;
    Eigen::Vector2d p2;
};

struct PointPose {
    Eigen::Vector2d position;
};

using ElementPose = std::variant<PointPose, LinePose>;

using ClusterPose = std::unordered_map<ConstraintGraph::NodeIdType, ElementPose>;

}  // namespace Gcs::Solvers::BottomUp
