// MathUtils::NodeId / EdgeId / SimpleGraph with the interface the constraint graph relies on
// (reference: src/structures/include/structures/simple_graph.hpp:23-188).  Only what the solver
// path needs is provided: id-stable node/edge insertion and removal, endpoint and incidence
// queries.
//
// Storage is flat: sorted vectors keyed by id (ids are handed out ascending, so insertion is an
// append).  A leaf component is a 3-node graph and a decomposition creates one per element of the
// sketch; node-based containers cost ~20 heap allocations per leaf there, sorted vectors three.
// Ids iterate ascending, the order the reference's flat_map property maps expose and its role
// assignment depends on.
#pragma once

#include <algorithm>
#include <compare>
#include <cstddef>
#include <expected>
#include <functional>
#include <stdexcept>
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

namespace detail {

// id -> value in a vector sorted by id.  find() is a linear scan while the table is tiny (the
// leaf case) and a binary search beyond; keys at or past the end append.
template <typename Key, typename Value>
class FlatTable {
public:
    using Entry = std::pair<Key, Value>;
    using iterator = typename std::vector<Entry>::iterator;
    using const_iterator = typename std::vector<Entry>::const_iterator;

    iterator begin() { return m_rows.begin(); }
    iterator end() { return m_rows.end(); }
    const_iterator begin() const { return m_rows.begin(); }
    const_iterator end() const { return m_rows.end(); }
    std::size_t size() const { return m_rows.size(); }
    bool empty() const { return m_rows.empty(); }
    void clear() { m_rows.clear(); }
    void reserve(std::size_t n) { m_rows.reserve(n); }

    const_iterator find(const Key& k) const { return locate(m_rows.begin(), m_rows.end(), k); }
    iterator find(const Key& k) { return locate(m_rows.begin(), m_rows.end(), k); }
    bool contains(const Key& k) const { return find(k) != m_rows.end(); }

    // insert or overwrite; returns the entry
    Entry& put(const Key& k, Value v)
    {
        if (m_rows.empty() || m_rows.back().first < k) return m_rows.emplace_back(k, std::move(v));
        auto it = std::lower_bound(m_rows.begin(), m_rows.end(), k, [](const Entry& e, const Key& key) { return e.first < key; });
        if (it != m_rows.end() && it->first == k) {
            it->second = std::move(v);
            return *it;
        }
        return *m_rows.insert(it, Entry(k, std::move(v)));
    }
    bool erase(const Key& k)
    {
        auto it = find(k);
        if (it == m_rows.end()) return false;
        m_rows.erase(it);
        return true;
    }

private:
    template <typename It>
    static It locate(It first, It last, const Key& k)
    {
        if (last - first <= 8) {
            for (It it = first; it != last; ++it)
                if (it->first == k) return it;
            return last;
        }
        // ids handed out in sequence and never erased sit at their own index: one probe
        if (k.value >= 0 && k.value < last - first && (first + k.value)->first == k) return first + k.value;
        It it = std::lower_bound(first, last, k, [](const Entry& e, const Key& key) { return e.first < key; });
        return (it != last && it->first == k) ? it : last;
    }
    std::vector<Entry> m_rows;
};

// A set of ids in a sorted vector (the virtual-edge set of a constraint graph: one or two members).
template <typename Key>
class FlatSet {
public:
    using const_iterator = typename std::vector<Key>::const_iterator;
    const_iterator begin() const { return m_keys.begin(); }
    const_iterator end() const { return m_keys.end(); }
    std::size_t size() const { return m_keys.size(); }
    bool empty() const { return m_keys.empty(); }
    void reserve(std::size_t n) { m_keys.reserve(n); }
    std::size_t count(const Key& k) const { return std::binary_search(m_keys.begin(), m_keys.end(), k) ? 1 : 0; }
    bool contains(const Key& k) const { return count(k) != 0; }
    bool insert(const Key& k)
    {
        if (m_keys.empty() || m_keys.back() < k) {
            m_keys.push_back(k);
            return true;
        }
        const auto it = std::lower_bound(m_keys.begin(), m_keys.end(), k);
        if (it != m_keys.end() && *it == k) return false;
        m_keys.insert(it, k);
        return true;
    }
    std::size_t erase(const Key& k)
    {
        const auto it = std::lower_bound(m_keys.begin(), m_keys.end(), k);
        if (it == m_keys.end() || !(*it == k)) return 0;
        m_keys.erase(it);
        return 1;
    }

private:
    std::vector<Key> m_keys;
};

}  // namespace detail

class SimpleGraph {
public:
    using NodeIdType = NodeId;
    using EdgeIdType = EdgeId;
    using EdgeList = std::vector<EdgeId>;  // ascending

    // room for this many nodes / edges, and for `degree` edges at every node added from now on
    void reserve(std::size_t nodes, std::size_t edges, std::size_t degree = 0)
    {
        m_incident.reserve(nodes), m_ends.reserve(edges);
        m_degreeHint = degree;
    }

    NodeId addNode()
    {
        const NodeId id { m_nextNode++ };
        EdgeList& list = m_incident.put(id, EdgeList {}).second;
        if (m_degreeHint) list.reserve(m_degreeHint);
        ++m_version;
        return id;
    }

    std::expected<EdgeId, GraphError> addEdge(NodeId s, NodeId t)
    {
        auto is = m_incident.find(s), it = m_incident.find(t);
        if (is == m_incident.end() || it == m_incident.end()) return std::unexpected(GraphError::NodeNotFound);
        const EdgeId id { m_nextEdge++ };
        m_ends.put(id, std::make_pair(s, t));
        is->second.push_back(id);  // the newest id is the largest: lists stay ascending
        if (it != is) it->second.push_back(id);
        ++m_version;
        return id;
    }

    std::expected<void, GraphError> removeEdge(EdgeId e)
    {
        auto f = m_ends.find(e);
        if (f == m_ends.end()) return std::unexpected(GraphError::EdgeNotFound);
        for (NodeId n : { f->second.first, f->second.second }) {
            auto in = m_incident.find(n);
            if (in != m_incident.end()) std::erase(in->second, e);
        }
        m_ends.erase(e);
        ++m_version;
        return {};
    }

    std::expected<void, GraphError> removeNode(NodeId n)
    {
        auto f = m_incident.find(n);
        if (f == m_incident.end()) return std::unexpected(GraphError::NodeNotFound);
        const EdgeList gone = f->second;
        for (EdgeId e : gone) removeEdge(e);
        m_incident.erase(n);
        ++m_version;
        return {};
    }

    std::vector<NodeId> getNodes() const
    {
        std::vector<NodeId> v;
        v.reserve(m_incident.size());
        for (const auto& kv : m_incident) v.push_back(kv.first);
        return v;
    }
    std::vector<EdgeId> getEdges() const
    {
        std::vector<EdgeId> v;
        v.reserve(m_ends.size());
        for (const auto& kv : m_ends) v.push_back(kv.first);
        return v;
    }
    const EdgeList& getEdges(NodeId n) const
    {
        const auto it = m_incident.find(n);
        if (it == m_incident.end()) throw std::out_of_range("SimpleGraph::getEdges: unknown node");
        return it->second;
    }
    std::pair<NodeId, NodeId> getEndpoints(EdgeId e) const
    {
        const auto it = m_ends.find(e);
        if (it == m_ends.end()) throw std::out_of_range("SimpleGraph::getEndpoints: unknown edge");
        return it->second;
    }
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
    bool hasNode(NodeId n) const { return m_incident.contains(n); }
    bool hasEdge(EdgeId e) const { return m_ends.contains(e); }

    std::expected<EdgeId, GraphError> getEdgeBetween(NodeId s, NodeId t) const
    {
        const auto is = m_incident.find(s);
        if (is == m_incident.end() || !hasNode(t)) return std::unexpected(GraphError::EdgeNotFound);
        for (EdgeId e : is->second) {
            const auto ends = m_ends.find(e)->second;
            if ((ends.first == s && ends.second == t) || (ends.first == t && ends.second == s)) return e;
        }
        return std::unexpected(GraphError::EdgeNotFound);
    }
    bool hasEdgeBetween(NodeId s, NodeId t) const { return getEdgeBetween(s, t).has_value(); }

    // counts the structural changes made so far (what ConstraintGraph keys its derived digest on)
    unsigned version() const { return m_version; }

private:
    int m_nextNode { 0 };
    int m_nextEdge { 0 };
    unsigned m_version { 0 };
    std::size_t m_degreeHint { 0 };
    detail::FlatTable<NodeId, EdgeList> m_incident;
    detail::FlatTable<EdgeId, std::pair<NodeId, NodeId>> m_ends;
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
