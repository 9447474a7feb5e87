// Degree-2 peeling decomposition (see gcs/b200/peel_decomposition.hpp).
#include <algorithm>
#include <array>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <exception>
#include <functional>
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
    (void)leaf.triangleDigest();  // what the leaf scheduler will read, taken while the leaf is in this thread's cache
    return leaf;
}

}  // namespace

std::vector<ConstraintGraph> decomposeByPeeling(const ConstraintGraph& gcs, PeelStats* stats)
{
    const auto& graph = gcs.getGraph();
    const auto t0 = std::chrono::steady_clock::now();
    if (graph.nodeCount() < 3) throw std::runtime_error("decomposeByPeeling: fewer than three elements");

    // live adjacency in flat arrays over the dense positions of the nodes in ascending id order:
    // the live incident edges of node slot i are inc[i] (ascending, as SimpleGraph keeps them; a
    // node of degree <= 2 is all the peel ever looks into, higher degrees only lose edges)
    const auto nodeList = graph.getNodes();
    const std::size_t nn = nodeList.size();
    std::vector<NodeId> ids(nodeList.begin(), nodeList.end());
    std::sort(ids.begin(), ids.end());
    auto slotOf = [&](NodeId n) -> std::size_t {
        if (n.value >= 0 && static_cast<std::size_t>(n.value) < nn && ids[static_cast<std::size_t>(n.value)] == n)
            return static_cast<std::size_t>(n.value);
        return static_cast<std::size_t>(std::lower_bound(ids.begin(), ids.end(), n) - ids.begin());
    };
    std::vector<std::vector<EdgeId>> inc(nn);
    std::vector<char> alive(nn, 1);
    for (std::size_t i = 0; i < nn; ++i) {
        const auto& es = graph.getEdges(ids[i]);
        inc[i].assign(es.begin(), es.end());
        std::sort(inc[i].begin(), inc[i].end());
    }
    auto other = [&](EdgeId e, NodeId n) {
        const auto [s, t] = graph.getEndpoints(e);
        return s == n ? t : s;
    };
    auto dropEdge = [&](std::size_t slot, EdgeId e) {
        auto& v = inc[slot];
        v.erase(std::find(v.begin(), v.end(), e));
    };
    // candidates in ascending id order: a min-heap of slots with lazy deletion (a slot is looked
    // at again whenever its degree becomes two)
    std::vector<std::size_t> heap;
    auto push = [&](std::size_t slot) {
        heap.push_back(slot);
        std::push_heap(heap.begin(), heap.end(), std::greater<>());
    };
    for (std::size_t i = 0; i < nn; ++i)
        if (inc[i].size() == 2) push(i);
    std::vector<std::size_t> parked;  // degree-2 slots whose two edges form a double edge (skipped, kept)

    const auto t1 = std::chrono::steady_clock::now();
    std::vector<Peel> peels;
    peels.reserve(nn);
    std::size_t remaining = nn;
    while (remaining > 3) {
        // smallest-id degree-2 node whose two edges lead to two different neighbours
        bool found = false;
        std::size_t vs = 0;
        while (!heap.empty()) {
            std::pop_heap(heap.begin(), heap.end(), std::greater<>());
            const std::size_t c = heap.back();
            heap.pop_back();
            if (!alive[c] || inc[c].size() != 2) continue;  // stale entry
            if (other(inc[c][0], ids[c]) == other(inc[c][1], ids[c])) {  // a double edge, not a separation pair
                parked.push_back(c);
                continue;
            }
            vs = c;
            found = true;
            break;
        }
        for (std::size_t c : parked) push(c);  // they stay candidates for later rounds, in id order
        parked.clear();
        if (!found)
            throw std::runtime_error("decomposeByPeeling: no degree-2 element left with " + std::to_string(remaining)
                + " elements remaining; general separation pairs need the OGDF-based decomposition of the reference");
        const NodeId v = ids[vs];
        const EdgeId e0 = inc[vs][0], e1 = inc[vs][1];
        const NodeId a = other(e0, v), b = other(e1, v);
        peels.push_back({ v, a, b, e0, e1 });
        const std::size_t sa = slotOf(a), sb = slotOf(b);
        dropEdge(sa, e0);
        dropEdge(sb, e1);
        alive[vs] = 0;
        inc[vs].clear();
        --remaining;
        for (std::size_t sl : { sa, sb })
            if (inc[sl].size() == 2) push(sl);
    }
    const auto t2 = std::chrono::steady_clock::now();
    // what the base-leaf code below iterates: the three remaining nodes with their live edges
    std::map<NodeId, std::set<EdgeId>> incident;
    for (std::size_t i = 0; i < nn; ++i)
        if (alive[i]) incident.emplace(ids[i], std::set<EdgeId>(inc[i].begin(), inc[i].end()));

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
    // the peeled leaves, in reverse peel order (= solve order).  Each is a small graph of its own
    // (a dozen allocations, shared_ptr copies of its elements and constraints) built from read-only
    // looks at the sketch: every host thread builds its share.
    const long long np = static_cast<long long>(peels.size());
    leaves.resize(static_cast<std::size_t>(np) + 1);
    std::exception_ptr failure;
#pragma omp parallel for schedule(static) if (np > 2048)
    for (long long k = 0; k < np; ++k) {
        const Peel& p = peels[static_cast<std::size_t>(np - 1 - k)];
        try {
            const std::pair<NodeId, NodeId> pair { p.a, p.b };
            leaves[static_cast<std::size_t>(k) + 1] = makeLeaf(gcs, { p.a, p.b, p.v }, { p.va, p.vb }, &pair);
        } catch (...) {
#pragma omp critical
            if (!failure) failure = std::current_exception();
        }
    }
    if (failure) std::rethrow_exception(failure);
    if (std::getenv("GCS_HOST_TRACE")) {
        const auto t3 = std::chrono::steady_clock::now();
        auto ms = [](auto a, auto b) { return std::chrono::duration<double>(b - a).count() * 1e3; };
        std::fprintf(stderr, "[host] peel: adjacency %.1f ms, peel order %.1f ms, leaf graphs %.1f ms, %zu leaves\n", ms(t0, t1), ms(t1, t2),
            ms(t2, t3), leaves.size());
    }
    if (stats) *stats = { graph.nodeCount(), graph.edgeCount(), leaves.size() };
    return leaves;
}

}  // namespace Gcs::B200
