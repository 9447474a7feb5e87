// The three-step driver: check (and, if needed, repair) constrainedness, decompose, solve the
// leaves (reference: src/constraint_solver/src/orchestration/geometric_constraint_system.cpp:9-26;
// its "Solver Called" print to stderr is not reproduced).
#include <stdexcept>

#include <gcs/orchestration/geometric_constraint_system.hpp>

namespace Gcs {

void GeometricConstraintSystem::solveGeometricConstraintSystem(ConstraintGraph& gcs)
{
    GcsSolvingStrategy& strategy = *m_strategy;
    const bool solvable = strategy.checkConstraintGraphConstrainedness(gcs) == Constrainedness::WELL_CONSTRAINED
        || strategy.resolve(gcs);  // resolve() is only consulted for graphs that fail the check
    if (!solvable) throw std::runtime_error("Gcs is not well-constrained, current algorithms do not support such inputs");
    std::vector<ConstraintGraph> leaves = strategy.decomposeConstraintGraph(gcs);
    strategy.solveGcs(leaves);  // with the top-down strategy: the batched wave scheduler
}

}  // namespace Gcs
