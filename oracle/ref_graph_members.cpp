// ref_graph_members.cpp — the ConstraintGraph members that the reference defines in
// src/constraint_solver/src/model/gcs_data_structures.cpp.  That translation unit cannot be
// compiled in this image (it includes structures/graph_algorithms.hpp, which uses C++23
// deducing-this, GCC >= 14; and the OGDF bridge).  The solver path only needs the accessors
// below, written here from their declarations and documented behaviour in
// includes/gcs/model/gcs_data_structures.hpp:31-148.  TEST INFRASTRUCTURE ONLY.
#include <stdexcept>

#include <gcs/model/gcs_data_structures.hpp>

namespace Gcs {

ConstraintGraph::EdgeIdType ConstraintGraph::addVirtualEdge(NodeIdType s, NodeIdType t)
{
    auto e = m_constraintGraph.addEdge(s, t);
    if (!e.has_value()) throw std::runtime_error("Failed to insert virtual edge");
    m_virtualEdges.insert(e.value());
    return e.value();
}

ConstraintGraphError ConstraintGraph::addElement(NodeIdType node, std::shared_ptr<Element> element)
{
    if (!m_constraintGraph.hasNode(node)) return ConstraintGraphError::NodeNotFound;
    m_elementNodeMap.set(node, std::move(element));
    return ConstraintGraphError::OK;
}

ConstraintGraphError ConstraintGraph::addConstraint(EdgeIdType edge, std::shared_ptr<Constraint> constraint)
{
    if (!m_constraintGraph.hasEdge(edge)) return ConstraintGraphError::EdgeNotFound;
    if (m_virtualEdges.contains(edge)) throw std::runtime_error("Virtual edges cannot carry constraints.");
    m_constraintEdgeMap.set(edge, std::move(constraint));
    return ConstraintGraphError::OK;
}

std::shared_ptr<Element> ConstraintGraph::getElement(NodeIdType node) const
{
    auto r = m_elementNodeMap.get(node);
    return r.has_value() ? r.value().get() : nullptr;
}

std::shared_ptr<Constraint> ConstraintGraph::getConstraintForEdge(EdgeIdType edge) const
{
    auto r = m_constraintEdgeMap.get(edge);
    return r.has_value() ? r.value().get() : nullptr;
}

int ConstraintGraph::numberOfSolvedElements() const
{
    int n = 0;
    for (const auto& [node, element] : m_elementNodeMap)
        if (element->isElementSet()) ++n;
    return n;
}

}  // namespace Gcs
