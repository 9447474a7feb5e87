// The strategy interface GeometricConstraintSystem drives (reference:
// includes/gcs/orchestration/solving_strategy.hpp:13-41).  Kept signature for signature so that a
// strategy written against the reference compiles here; the top-down implementation in this repo
// (gcs/decomposition/top_down/stree_top_down_strategy.hpp) overrides solveGcs with the batched
// wave scheduler.
#pragma once

#include <vector>

#include <gcs/export.hpp>
#include <gcs/model/gcs_data_structures.hpp>

namespace Gcs {

// Outcome of the degree-of-freedom count (2n - 3 against the number of constraints).
enum class Constrainedness {
    UNDER_CONSTRAINED,
    WELL_CONSTRAINED,
    CONSISTENTLY_OVER_CONSTRAINED,
    INCONSISTENTLY_OVER_CONSTRAINED,
};

class GCS_API GcsSolvingStrategy {
public:
    virtual ~GcsSolvingStrategy() = default;

    // Step 1: is the sketch solvable as it stands?
    virtual Constrainedness checkConstraintGraphConstrainedness(const ConstraintGraph&) = 0;
    // Step 1b: try to repair a sketch that is not; false = give up (the driver then throws).
    virtual bool resolve(ConstraintGraph&) = 0;
    // Step 2: split into 3-element leaf components that share Element objects.
    virtual std::vector<ConstraintGraph> decomposeConstraintGraph(ConstraintGraph&) = 0;
    // Step 3: solve the leaves; results are written into the shared elements.
    virtual void solveGcs(std::vector<ConstraintGraph>&) = 0;

    // value semantics as in the reference: copyable and movable
    GcsSolvingStrategy() = default;
    GcsSolvingStrategy(const GcsSolvingStrategy&) = default;
    GcsSolvingStrategy& operator=(const GcsSolvingStrategy&) = default;
    GcsSolvingStrategy(GcsSolvingStrategy&&) = default;
    GcsSolvingStrategy& operator=(GcsSolvingStrategy&&) = default;
};

}  // namespace Gcs
