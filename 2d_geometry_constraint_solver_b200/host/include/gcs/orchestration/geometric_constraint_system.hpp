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

class GCS_API GeometricConstraintSystem final {
public:
    explicit GeometricConstraintSystem(std::unique_ptr<GcsSolvingStrategy> strategy) : m_strategy(std::move(strategy)) {}
    // throws std::runtime_error when the graph is not well-constrained and resolve() fails
    void solveGeometricConstraintSystem(ConstraintGraph& gcs);
    [[nodiscard]] const GcsSolvingStrategy& getStrategy() const { return *m_strategy; }
    [[nodiscard]] GcsSolvingStrategy& getStrategy() { return *m_strategy; }

private:
    std::unique_ptr<GcsSolvingStrategy> m_strategy;
};

}  // namespace Gcs
