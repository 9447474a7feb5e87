// ConstraintGraph: graph + element / constraint property maps + virtual-edge set, with the
// accessor surface the sub-problem solvers use (reference:
// includes/gcs/model/gcs_data_structures.hpp:31-148).  Graph splitting (getSeparatingGraphs,
// OGDF triconnectivity) is decomposition work outside the accelerated path and is not provided.
#pragma once

#include <cstddef>
#include <cstdint>
#include <expected>
#include <memory>
#include <vector>

#include <gcs/export.hpp>
#include <gcs/model/constraints.hpp>
#include <gcs/model/elements.hpp>
#include <structures/property_map.hpp>
#include <structures/simple_graph.hpp>

namespace Gcs {

enum class ConstraintGraphError { OK, NodeNotFound, EdgeNotFound };

// What the batched leaf scheduler (gcs/b200/leaf_batch.hpp) reads from a three-element component,
// kept in the graph object itself so that planning 1e5 leaves is one pass over the vector of
// leaves instead of a walk through five small containers per leaf: the elements in ascending node
// id and, for each node pair, the constraint on the edge between them (null: no edge, a virtual
// edge, or an edge without a constraint).  Pointers only - element flags and constraint values are
// read through them when a plan is made.  `simple` is false for anything else than three nodes
// with at most one edge per node pair (such components go through the container walk).
struct TriangleDigest {
    Element* element[3] {};
    const Constraint* constraint[3] {};  // node pairs (0,1), (0,2), (1,2)
    int edgeCount = 0;                   // virtual edges included
    bool simple = false;
};

class GCS_API ConstraintGraph final {
public:
    using Graph = MathUtils::SimpleGraph;
    using NodeIdType = Graph::NodeIdType;
    using EdgeIdType = Graph::EdgeIdType;
    using ElementMap = MathUtils::NodePropertyMap<std::shared_ptr<Element>>;
    using ConstraintMap = MathUtils::EdgePropertyMap<std::shared_ptr<Constraint>>;
    // (an std::unordered_set in the reference: two heap blocks for the one member a leaf has)
    using VirtualEdgeSet = MathUtils::detail::FlatSet<EdgeIdType>;

    // ---- the underlying graph (nodes and edges are created there, then decorated here) ----
    Graph& getGraph() { return m_constraintGraph; }
    const Graph& getGraph() const { return m_constraintGraph; }
    // storage for a graph of this size in one allocation per container (a decomposition builds 1e5
    // three-element leaves); `degree`: edges expected at every node
    void reserve(std::size_t nodes, std::size_t edges, std::size_t degree = 0)
    {
        m_constraintGraph.reserve(nodes, edges, degree);
        m_elementNodeMap.reserve(nodes), m_constraintEdgeMap.reserve(edges), m_virtualEdges.reserve(2);
    }
    std::size_t nodeCount() const { return m_constraintGraph.nodeCount(); }
    std::size_t edgeCount() const { return m_constraintGraph.edgeCount(); }  // virtual edges included
    std::expected<EdgeIdType, ConstraintGraphError> getEdgeBetween(NodeIdType s, NodeIdType t) const;

    // ---- elements on nodes ----
    ConstraintGraphError addElement(NodeIdType node, std::shared_ptr<Element> element);
    ConstraintGraphError removeElement(NodeIdType node);
    std::shared_ptr<Element> getElement(NodeIdType node) const;  // null when absent
    std::vector<std::shared_ptr<Element>> getElements() const;
    const ElementMap& getElementMap() const { return m_elementNodeMap; }  // ascending node id
    int numberOfSolvedElements() const;

    // ---- constraints on edges ----
    ConstraintGraphError addConstraint(EdgeIdType edge, std::shared_ptr<Constraint> constraint);
    ConstraintGraphError removeConstraintEdge(EdgeIdType edge);
    std::shared_ptr<Constraint> getConstraintForEdge(EdgeIdType edge) const;  // null when absent
    // Throws std::bad_expected_access, like the reference, when there is no edge or the edge is a
    // virtual one that carries no constraint (gcs_data_structures.hpp:66-71).
    std::shared_ptr<Constraint> getConstraintBetweenNodes(NodeIdType s, NodeIdType t) const;
    std::vector<std::shared_ptr<Constraint>> getConstraints() const;  // real constraints only
    const ConstraintMap& getConstraintMap() const { return m_constraintEdgeMap; }

    // ---- virtual edges: the separation-pair edge a split leaves behind; no constraint on it ----
    EdgeIdType addVirtualEdge(NodeIdType s, NodeIdType t);
    ConstraintGraphError removeVirtualEdge(EdgeIdType virtualEdge);
    bool hasVirtualEdge() const { return !m_virtualEdges.empty(); }
    bool isVirtualEdge(EdgeIdType edge) const { return m_virtualEdges.count(edge) != 0; }
    const VirtualEdgeSet& getVirtualEdges() const { return m_virtualEdges; }

    // Derived from the containers above on first use and again after any change to them (through
    // this class or through getGraph()); the decomposition asks for it while the leaf it has just
    // built is still in cache.  Not thread safe on one graph object, like the rest of the class.
    const TriangleDigest& triangleDigest() const;

    // 2n - 3 - edges, in int (the reference's strategy computes it in size_t and wraps)
    int getDeficit() const { return (2 * static_cast<int>(nodeCount()) - 3) - static_cast<int>(edgeCount()); }

private:
    Graph m_constraintGraph;
    ElementMap m_elementNodeMap;
    ConstraintMap m_constraintEdgeMap;
    VirtualEdgeSet m_virtualEdges;
    unsigned m_version = 0;  // changes to the property maps / virtual-edge set (the graph counts its own)
    mutable TriangleDigest m_digest;
    mutable std::uint64_t m_digestStamp = ~std::uint64_t { 0 };  // (graph version, m_version) the digest was taken at
};

}  // namespace Gcs
