// DeficitStreeBasedTopDownStrategy (reference:
// src/constraint_solver/src/decomposition/top_down/stree_top_down_strategy.cpp:12-45).
#include <stdexcept>

#include <gcs/b200/peel_decomposition.hpp>
#include <gcs/decomposition/top_down/stree_top_down_strategy.hpp>

namespace Gcs {

Constrainedness DeficitStreeBasedTopDownStrategy::checkConstraintGraphConstrainedness(const ConstraintGraph& gcs)
{
    // The reference computes the deficit in std::size_t (stree_top_down_strategy.cpp:16): it can
    // never be negative, an over-constrained graph wraps around and reports UNDER_CONSTRAINED.
    const std::size_t deficit = (2 * gcs.nodeCount() - 3) - gcs.edgeCount();
    if (deficit == 0) return Constrainedness::WELL_CONSTRAINED;
    return Constrainedness::UNDER_CONSTRAINED;
}

bool DeficitStreeBasedTopDownStrategy::resolve(ConstraintGraph& /*gcs*/) { return false; }

std::vector<ConstraintGraph> DeficitStreeBasedTopDownStrategy::decomposeConstraintGraph(ConstraintGraph& gcs)
{
    // A triconnected 3-element graph is its own single leaf (stree_top_down_strategy.cpp:54-57).
    if (gcs.nodeCount() == 3) return { gcs };
    // Otherwise: the reference's split rules applied to degree-2 separation pairs (no OGDF here);
    // throws for graphs that need general separation pairs.
    return B200::decomposeByPeeling(gcs);
}

void DeficitStreeBasedTopDownStrategy::solveGcs(std::vector<ConstraintGraph>& splitComponents)
{
    // reference: std::ranges::for_each(splitComponents, classifyAndSolve), results discarded
    m_report = m_devices > 1 ? B200::solveLeavesOnDevices(splitComponents, m_devices, m_minRows) : B200::solveLeaves(splitComponents, m_device);
}

}  // namespace Gcs
