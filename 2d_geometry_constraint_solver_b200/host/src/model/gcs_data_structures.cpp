// ConstraintGraph accessors used by the solver path (reference:
// includes/gcs/model/gcs_data_structures.hpp:31-148 and the matching members of
// src/model/gcs_data_structures.cpp).  Graph splitting is not part of this path.
#include <stdexcept>

#include <gcs/model/gcs_data_structures.hpp>

namespace Gcs {

ConstraintGraphError ConstraintGraph::addElement(NodeIdType node, std::shared_ptr<Element> element)
{
    if (!m_constraintGraph.hasNode(node)) return ConstraintGraphError::NodeNotFound;
    m_elementNodeMap.set(node, std::move(element));
    ++m_version;
    return ConstraintGraphError::OK;
}

ConstraintGraphError ConstraintGraph::addConstraint(EdgeIdType edge, std::shared_ptr<Constraint> constraint)
{
    if (!m_constraintGraph.hasEdge(edge)) return ConstraintGraphError::EdgeNotFound;
    if (m_virtualEdges.count(edge) != 0) throw std::runtime_error("Virtual edges cannot carry constraints.");
    m_constraintEdgeMap.set(edge, std::move(constraint));
    ++m_version;
    return ConstraintGraphError::OK;
}

std::expected<ConstraintGraph::EdgeIdType, ConstraintGraphError> ConstraintGraph::getEdgeBetween(NodeIdType s, NodeIdType t) const
{
    const auto edge = m_constraintGraph.getEdgeBetween(s, t);
    if (!edge.has_value()) return std::unexpected(ConstraintGraphError::EdgeNotFound);
    return edge.value();
}

ConstraintGraph::EdgeIdType ConstraintGraph::addVirtualEdge(NodeIdType s, NodeIdType t)
{
    const auto e = m_constraintGraph.addEdge(s, t);
    if (!e.has_value()) throw std::runtime_error("Failed to insert virtual edge");
    m_virtualEdges.insert(e.value());
    ++m_version;
    return e.value();
}

ConstraintGraphError ConstraintGraph::removeVirtualEdge(EdgeIdType virtualEdge)
{
    if (m_virtualEdges.erase(virtualEdge) == 0) return ConstraintGraphError::EdgeNotFound;
    ++m_version;
    if (!m_constraintGraph.removeEdge(virtualEdge).has_value()) return ConstraintGraphError::EdgeNotFound;
    return ConstraintGraphError::OK;
}

ConstraintGraphError ConstraintGraph::removeElement(NodeIdType node)
{
    if (!m_constraintGraph.hasNode(node)) return ConstraintGraphError::NodeNotFound;
    ++m_version;
    for (EdgeIdType e : std::vector<EdgeIdType>(m_constraintGraph.getEdges(node).begin(), m_constraintGraph.getEdges(node).end())) {
        (void)m_constraintEdgeMap.erase(e);
        m_virtualEdges.erase(e);
    }
    (void)m_constraintGraph.removeNode(node);
    (void)m_elementNodeMap.erase(node);
    return ConstraintGraphError::OK;
}

ConstraintGraphError ConstraintGraph::removeConstraintEdge(EdgeIdType edge)
{
    if (!m_constraintGraph.hasEdge(edge)) return ConstraintGraphError::EdgeNotFound;
    ++m_version;
    (void)m_constraintEdgeMap.erase(edge);
    m_virtualEdges.erase(edge);
    (void)m_constraintGraph.removeEdge(edge);
    return ConstraintGraphError::OK;
}

std::shared_ptr<Element> ConstraintGraph::getElement(NodeIdType node) const
{
    const auto r = m_elementNodeMap.get(node);
    return r.has_value() ? r.value().get() : nullptr;
}

std::shared_ptr<Constraint> ConstraintGraph::getConstraintForEdge(EdgeIdType edge) const
{
    const auto r = m_constraintEdgeMap.get(edge);
    return r.has_value() ? r.value().get() : nullptr;
}

std::shared_ptr<Constraint> ConstraintGraph::getConstraintBetweenNodes(NodeIdType s, NodeIdType t) const
{
    // two .value() calls, as in the reference: no edge, or an edge without a constraint (a virtual
    // one), surfaces as std::bad_expected_access
    const EdgeIdType edge = m_constraintGraph.getEdgeBetween(s, t).value();
    return m_constraintEdgeMap.get(edge).value().get();
}

const TriangleDigest& ConstraintGraph::triangleDigest() const
{
    const std::uint64_t stamp = (static_cast<std::uint64_t>(m_constraintGraph.version()) << 32) | m_version;
    if (m_digestStamp == stamp) return m_digest;
    TriangleDigest d;
    d.edgeCount = static_cast<int>(m_constraintGraph.edgeCount());
    NodeIdType node[3];
    int n = 0;
    bool ok = m_constraintGraph.nodeCount() == 3 && m_elementNodeMap.size() == 3 && d.edgeCount <= 3;
    if (ok)
        for (const auto& [nd, el] : m_elementNodeMap) {  // ascending node id
            if (!el) ok = false;
            node[n] = nd;
            d.element[n++] = el.get();
        }
    // at most one edge per node pair: every edge then is THE edge between its two nodes, and the
    // constraint map has nothing the three pairs do not show
    int seen = 0, held = 0;
    for (int a = 0; ok && a < 3; ++a)
        for (int b = a + 1; b < 3; ++b) {
            const auto edge = m_constraintGraph.getEdgeBetween(node[a], node[b]);
            if (!edge.has_value()) continue;
            ++seen;
            if (m_virtualEdges.count(edge.value()) != 0) continue;
            const auto c = m_constraintEdgeMap.get(edge.value());
            if (!c.has_value()) continue;
            if (!c.value().get()) ok = false;  // a null constraint on a real edge: the container walk decides
            d.constraint[a + b - 1] = c.value().get().get();
            ++held;
        }
    ok = ok && seen == d.edgeCount && held == static_cast<int>(m_constraintEdgeMap.size());
    d.simple = ok;
    m_digest = d;
    m_digestStamp = stamp;
    return m_digest;
}

int ConstraintGraph::numberOfSolvedElements() const
{
    int n = 0;
    for (const auto& [node, element] : m_elementNodeMap)
        if (element->isElementSet()) ++n;
    return n;
}

std::vector<std::shared_ptr<Element>> ConstraintGraph::getElements() const
{
    std::vector<std::shared_ptr<Element>> v;
    for (const auto& [node, element] : m_elementNodeMap) v.push_back(element);
    return v;
}

std::vector<std::shared_ptr<Constraint>> ConstraintGraph::getConstraints() const
{
    std::vector<std::shared_ptr<Constraint>> v;
    for (const auto& [edge, constraint] : m_constraintEdgeMap) v.push_back(constraint);
    return v;
}

}  // namespace Gcs
