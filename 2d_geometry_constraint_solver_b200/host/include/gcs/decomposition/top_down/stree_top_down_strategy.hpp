// DeficitStreeBasedTopDownStrategy (reference:
// includes/gcs/decomposition/top_down/stree_top_down_strategy.hpp:15-31,
// src/decomposition/top_down/stree_top_down_strategy.cpp:12-79).
//
// solveGcs is the accelerated entry point: the reference's sequential
// `for_each(leaves, classifyAndSolve)` becomes Gcs::B200::solveLeaves (same final element state,
// one kernel launch per equation kind per dependency wave).  decomposeConstraintGraph applies the
// reference's S-tree split rules to degree-2 separation pairs (gcs/b200/peel_decomposition.hpp;
// OGDF is not available here): Henneberg-style sketches decompose completely, graphs that need
// general separation pairs throw - supply their leaves from the reference's own decomposition.
#pragma once

#include <vector>

#include <gcs/b200/leaf_batch.hpp>
#include <gcs/export.hpp>
#include <gcs/model/gcs_data_structures.hpp>
#include <gcs/orchestration/solving_strategy.hpp>

namespace Gcs {

class GCS_API DeficitStreeBasedTopDownStrategy : public GcsSolvingStrategy {
public:
    Constrainedness checkConstraintGraphConstrainedness(const ConstraintGraph& gcs) override;
    bool resolve(ConstraintGraph& gcs) override;
    std::vector<ConstraintGraph> decomposeConstraintGraph(ConstraintGraph& gcs) override;
    void solveGcs(std::vector<ConstraintGraph>& splitComponents) override;
    ~DeficitStreeBasedTopDownStrategy() override = default;

    // what the last solveGcs did (leaves, waves, launches, per-leaf results)
    const B200::BatchReport& lastReport() const { return m_report; }
    void setDevice(int device) { m_device = device; }
    // spread every wave over the first n devices of gcs_b200_init (B200::solveLeavesOnDevices); 1 = one device
    void setDeviceCount(int n, std::size_t minRowsPerDevice = 16384) { m_devices = n < 1 ? 1 : n, m_minRows = minRowsPerDevice; }

private:
    B200::BatchReport m_report;
    int m_device = 0;
    int m_devices = 1;
    std::size_t m_minRows = 16384;
};

}  // namespace Gcs
