// GeometricConstraintSystem: the 3-step driver (reference:
// includes/gcs/orchestration/geometric_constraint_system.hpp:14-31,
// src/orchestration/geometric_constraint_system.cpp:9-26).
#pragma once

#include <memory>
#include <utility>

#include <gcs/decomposition/top_down/stree_top_down_strategy.hpp>
#include <gcs/export.hpp>
#include <gcs/orchestration/solving_strategy.hpp>

namespace Gcs {

// Owns one strategy and runs its three steps on a sketch.  With
// DeficitStreeBasedTopDownStrategy the last step is the batched wave scheduler of
// gcs/b200/leaf_batch.hpp: one kernel launch per equation-pair kind per dependency wave.
class GCS_API GeometricConstraintSystem final {
    std::unique_ptr<GcsSolvingStrategy> m_strategy;

public:
    explicit GeometricConstraintSystem(std::unique_ptr<GcsSolvingStrategy> strategy) : m_strategy(std::move(strategy)) {}

    // check -> (resolve) -> decompose -> solve.  Throws std::runtime_error("Gcs is not
    // well-constrained, ...") when the check fails and resolve() cannot repair the sketch, and
    // whatever the strategy throws (no CUDA device: std::runtime_error; malformed leaf:
    // std::bad_expected_access).  Results land in the Element objects of `gcs`.
    void solveGeometricConstraintSystem(ConstraintGraph& gcs);

    [[nodiscard]] GcsSolvingStrategy& getStrategy() { return *m_strategy; }
    [[nodiscard]] const GcsSolvingStrategy& getStrategy() const { return *m_strategy; }
};

}  // namespace Gcs
