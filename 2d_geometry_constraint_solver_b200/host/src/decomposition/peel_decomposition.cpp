// Degree-2 peeling decomposition (see gcs/b200/peel_decomposition.hpp).
#include <algorithm>
#include <array>
#include <map>
#include <set>
#include <stdexcept>
#include <unordered_map>

#include <gcs/b200/peel_decomposition.hpp>

namespace Gcs::B200 {

namespace {

using NodeId = ConstraintGraph::NodeIdType;
using EdgeId = ConstraintGraph::EdgeIdType;

struct Peel {
    NodeId v, a, b;
    EdgeId va, vb;
};

// one leaf: nodes in ascending original id (the order role assignment sees), the given edges
// with their constraints (or as virtual edges), plus an optional fresh virtual edge
ConstraintGraph makeLeaf(const ConstraintGraph& g, std::array<NodeId, 3> nodes, const std::vector<EdgeId>& edges,
    const std::pair<NodeId, NodeId>* virtualPair)
{
    std::sort(nodes.begin(), nodes.end());
    ConstraintGraph leaf;
    std::array<NodeId, 3> local {};
    for (int i = 0; i < 3; ++i) {
        local[static_cast<std::size_t>(i)] = leaf.getGraph().addNode();
        leaf.addElement(local[static_cast<std::size_t>(i)], g.getElement(nodes[static_cast<std::size_t>(i)]));
    }
    auto toLocal = [&](NodeId n) {
        for (std::size_t i = 0; i < 3; ++i)
            if (nodes[i] == n) return local[i];
        throw std::logic_error("peel decomposition: edge endpoint outside its leaf");
    };
    for (EdgeId e : edges) {
        const auto [s, t] = g.getGraph().getEndpoints(e);
        if (g.isVirtualEdge(e)) {
            leaf.addVirtualEdge(toLocal(s), toLocal(t));
        } else {
            const EdgeId le = leaf.getGraph().addEdge(toLocal(s), toLocal(t)).value();
            if (auto c = g.getConstraintForEdge(e)) leaf.addConstraint(le, c);
        }
    }
    if (virtualPair) leaf.addVirtualEdge(toLocal(virtualPair->first), toLocal(virtualPair->second));
    return leaf;
}

}  // namespace

std::vector<ConstraintGraph> decomposeByPeeling(const ConstraintGraph& gcs, PeelStats* stats)
{
    const auto& graph = gcs.getGraph();
    if (graph.nodeCount() < 3) throw std::runtime_error("decomposeByPeeling: fewer than three elements");

    // live adjacency: node -> incident live edges
    std::map<NodeId, std::set<EdgeId>> incident;
    for (NodeId n : graph.getNodes()) {
        const auto& es = graph.getEdges(n);
        incident.emplace(n, std::set<EdgeId>(es.begin(), es.end()));
    }
    auto other = [&](EdgeId e, NodeId n) {
        const auto [s, t] = graph.getEndpoints(e);
        return s == n ? t : s;
    };
    std::set<NodeId> degreeTwo;
    for (const auto& [n, es] : incident)
        if (es.size() == 2) degreeTwo.insert(n);

    std::vector<Peel> peels;
    peels.reserve(graph.nodeCount());
    while (incident.size() > 3) {
        // smallest-id degree-2 node whose two edges lead to two different neighbours
        NodeId v {};
        bool found = false;
        for (auto it = degreeTwo.begin(); it != degreeTwo.end();) {
            const auto& es = incident.at(*it);
            if (es.size() != 2) {
                it = degreeTwo.erase(it);
                continue;
            }
            const EdgeId e0 = *es.begin(), e1 = *std::next(es.begin());
            if (other(e0, *it) == other(e1, *it)) {  // a double edge, not a separation pair
                ++it;
                continue;
            }
            v = *it;
            found = true;
            break;
        }
        if (!found)
            throw std::runtime_error("decomposeByPeeling: no degree-2 element left with " + std::to_string(incident.size())
                + " elements remaining; general separation pairs need the OGDF-based decomposition of the reference");
        const auto es = incident.at(v);
        const EdgeId e0 = *es.begin(), e1 = *std::next(es.begin());
        const NodeId a = other(e0, v), b = other(e1, v);
        peels.push_back({ v, a, b, e0, e1 });
        incident.at(a).erase(e0);
        incident.at(b).erase(e1);
        incident.erase(v);
        degreeTwo.erase(v);
        for (NodeId n : { a, b }) {
            if (incident.at(n).size() == 2)
                degreeTwo.insert(n);
            else
                degreeTwo.erase(n);
        }
    }

    std::vector<ConstraintGraph> leaves;
    leaves.reserve(peels.size() + 1);
    {  // the base: the three remaining elements with every edge still alive between them
        std::array<NodeId, 3> base {};
        std::set<EdgeId> edges;
        std::size_t i = 0;
        for (const auto& [n, es] : incident) {
            base[i++] = n;
            edges.insert(es.begin(), es.end());
        }
        leaves.push_back(makeLeaf(gcs, base, std::vector<EdgeId>(edges.begin(), edges.end()), nullptr));
    }
    for (auto it = peels.rbegin(); it != peels.rend(); ++it) {
        const std::pair<NodeId, NodeId> pair { it->a, it->b };
        leaves.push_back(makeLeaf(gcs, { it->a, it->b, it->v }, { it->va, it->vb }, &pair));
    }
    if (stats) *stats = { graph.nodeCount(), graph.edgeCount(), leaves.size() };
    return leaves;
}

}  // namespace Gcs::B200
