// GeometricConstraintSystem::solveGeometricConstraintSystem (reference:
// src/constraint_solver/src/orchestration/geometric_constraint_system.cpp:9-26).
#include <stdexcept>

#include <gcs/orchestration/geometric_constraint_system.hpp>

namespace Gcs {

void GeometricConstraintSystem::solveGeometricConstraintSystem(ConstraintGraph& gcs)
{
    if (m_strategy->checkConstraintGraphConstrainedness(gcs) != Constrainedness::WELL_CONSTRAINED) {
        if (!m_strategy->resolve(gcs))
            throw std::runtime_error("Gcs is not well-constrained, current algorithms do not support such inputs");
    }
    auto decomposition = m_strategy->decomposeConstraintGraph(gcs);
    m_strategy->solveGcs(decomposition);
}

}  // namespace Gcs
