// MathUtils::NodeId / EdgeId / SimpleGraph with the interface the constraint graph relies on
// (reference: src/structures/include/structures/simple_graph.hpp:23-188).  Only what the solver
// path needs is provided: id-stable node/edge insertion and removal, endpoint and incidence
// queries.  Storage is ordered (std::map) - ids iterate ascending, which is the order the
// reference's flat_map property maps expose and its role assignment depends on.
#pragma once

#include <compare>
#include <cstddef>
#include <expected>
#include <functional>
#include <map>
#include <set>
#include <utility>
#include <vector>

namespace MathUtils {

enum class GraphError { NodeNotFound, EdgeNotFound, InternalError };

struct NodeId {
    int value {};
    bool operator==(const NodeId&) const = default;
    auto operator<=>(const NodeId&) const = default;
};

struct EdgeId {
    int value {};
    bool operator==(const EdgeId&) const = default;
    auto operator<=>(const EdgeId&) const = default;
};

class SimpleGraph {
public:
    using NodeIdType = NodeId;
    using EdgeIdType = EdgeId;

    NodeId addNode()
    {
        const NodeId id { m_nextNode++ };
        m_incident.emplace(id, std::set<EdgeId> {});
        return id;
    }

    std::expected<EdgeId, GraphError> addEdge(NodeId s, NodeId t)
    {
        auto is = m_incident.find(s), it = m_incident.find(t);
        if (is == m_incident.end() || it == m_incident.end()) return std::unexpected(GraphError::NodeNotFound);
        const EdgeId id { m_nextEdge++ };
        m_ends.emplace(id, std::make_pair(s, t));
        is->second.insert(id);
        it->second.insert(id);
        return id;
    }

    std::expected<void, GraphError> removeEdge(EdgeId e)
    {
        auto f = m_ends.find(e);
        if (f == m_ends.end()) return std::unexpected(GraphError::EdgeNotFound);
        m_incident[f->second.first].erase(e);
        m_incident[f->second.second].erase(e);
        m_ends.erase(f);
        return {};
    }

    std::expected<void, GraphError> removeNode(NodeId n)
    {
        auto f = m_incident.find(n);
        if (f == m_incident.end()) return std::unexpected(GraphError::NodeNotFound);
        const std::vector<EdgeId> gone(f->second.begin(), f->second.end());
        for (EdgeId e : gone) removeEdge(e);
        m_incident.erase(n);
        return {};
    }

    std::vector<NodeId> getNodes() const
    {
        std::vector<NodeId> v;
        for (const auto& kv : m_incident) v.push_back(kv.first);
        return v;
    }
    std::vector<EdgeId> getEdges() const
    {
        std::vector<EdgeId> v;
        for (const auto& kv : m_ends) v.push_back(kv.first);
        return v;
    }
    const std::set<EdgeId>& getEdges(NodeId n) const { return m_incident.at(n); }
    std::pair<NodeId, NodeId> getEndpoints(EdgeId e) const { return m_ends.at(e); }
    std::vector<NodeId> getNeighbors(NodeId n) const
    {
        std::vector<NodeId> v;
        for (EdgeId e : getEdges(n)) {
            const auto [a, b] = getEndpoints(e);
            v.push_back(a == n ? b : a);
        }
        return v;
    }

    std::size_t nodeCount() const { return m_incident.size(); }
    std::size_t edgeCount() const { return m_ends.size(); }
    bool hasNode(NodeId n) const { return m_incident.count(n) != 0; }
    bool hasEdge(EdgeId e) const { return m_ends.count(e) != 0; }

    std::expected<EdgeId, GraphError> getEdgeBetween(NodeId s, NodeId t) const
    {
        auto is = m_incident.find(s);
        if (is == m_incident.end() || !hasNode(t)) return std::unexpected(GraphError::EdgeNotFound);
        for (EdgeId e : is->second) {
            const auto& ends = m_ends.at(e);
            if ((ends.first == s && ends.second == t) || (ends.first == t && ends.second == s)) return e;
        }
        return std::unexpected(GraphError::EdgeNotFound);
    }
    bool hasEdgeBetween(NodeId s, NodeId t) const { return getEdgeBetween(s, t).has_value(); }

private:
    int m_nextNode { 0 };
    int m_nextEdge { 0 };
    std::map<NodeId, std::set<EdgeId>> m_incident;
    std::map<EdgeId, std::pair<NodeId, NodeId>> m_ends;
};

}  // namespace MathUtils

template <>
struct std::hash<MathUtils::NodeId> {
    std::size_t operator()(const MathUtils::NodeId& id) const noexcept { return std::hash<int> {}(id.value); }
};
template <>
struct std::hash<MathUtils::EdgeId> {
    std::size_t operator()(const MathUtils::EdgeId& id) const noexcept { return std::hash<int> {}(id.value); }
};
