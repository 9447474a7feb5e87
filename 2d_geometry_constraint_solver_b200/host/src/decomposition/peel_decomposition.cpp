// Degree-2 peeling decomposition (see gcs/b200/peel_decomposition.hpp).
#include <algorithm>
#include <array>
#include <chrono>
#include <cstdint>
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

// one peeled node with its two neighbours and the two edges to them (positions in the flat arrays below)
struct Peel {
    std::uint32_t v, a, b;
    std::uint32_t va, vb;
};

// The sketch as the leaf builder reads it: element / constraint handles and edge ends by position,
// gathered in one pass over the sketch's own containers (which 1e5 leaf builders would otherwise
// each search).
struct SketchView {
    std::vector<NodeId> ids;                                       // ascending
    std::vector<EdgeId> eids;                                      // ascending
    std::vector<std::uint32_t> ea, eb;                             // edge ends, as node positions
    std::vector<const std::shared_ptr<Element>*> element;         // per node position (null: none)
    std::vector<const std::shared_ptr<Constraint>*> constraint;   // per edge position (null: none)
    std::vector<char> isVirtual;                                   // per edge position
};

// A set of indices below n with insert / erase / smallest member in a handful of operations: a
// bitmap with two summary levels above it (64-way), all of it a few KB.
class MinSet {
public:
    explicit MinSet(std::size_t n)
    {
        std::size_t words = (n + 63) / 64;
        for (int l = 0; l < 3; ++l) {
            m_level[l].assign(std::max<std::size_t>(words, 1), 0);
            words = (words + 63) / 64;
        }
        if (m_level[2].size() > 1) throw std::length_error("MinSet: more than 2^18 * 64 indices");
    }
    bool empty() const { return m_level[2][0] == 0; }
    void insert(std::size_t i)
    {
        for (int l = 0; l < 3; ++l) {
            m_level[l][i >> 6] |= std::uint64_t { 1 } << (i & 63);
            i >>= 6;
        }
    }
    void erase(std::size_t i)
    {
        for (int l = 0; l < 3; ++l) {
            std::uint64_t& w = m_level[l][i >> 6];
            w &= ~(std::uint64_t { 1 } << (i & 63));
            if (w != 0) return;  // the word still has members: the summaries stand
            i >>= 6;
        }
    }
    std::size_t min() const  // not empty
    {
        std::size_t i = 0;
        for (int l = 2; l >= 0; --l) i = (i << 6) | static_cast<std::size_t>(__builtin_ctzll(m_level[l][i]));
        return i;
    }

private:
    std::vector<std::uint64_t> m_level[3];
};

// one leaf: nodes in ascending original id (the order role assignment sees), the given edges
// with their constraints (or as virtual edges), plus an optional fresh virtual edge
ConstraintGraph makeLeaf(const SketchView& g, std::array<std::uint32_t, 3> nodes, const std::uint32_t* edges, std::size_t edgeCount,
    const std::pair<std::uint32_t, std::uint32_t>* virtualPair)
{
    std::sort(nodes.begin(), nodes.end());  // positions ascend with ids
    ConstraintGraph leaf;
    leaf.reserve(3, edgeCount + (virtualPair ? 1 : 0), 2);
    std::array<NodeId, 3> local {};
    for (std::size_t i = 0; i < 3; ++i) {
        local[i] = leaf.getGraph().addNode();
        const auto* el = g.element[nodes[i]];
        leaf.addElement(local[i], el ? *el : std::shared_ptr<Element> {});
    }
    auto toLocal = [&](std::uint32_t n) {
        for (std::size_t i = 0; i < 3; ++i)
            if (nodes[i] == n) return local[i];
        throw std::logic_error("peel decomposition: edge endpoint outside its leaf");
    };
    for (std::size_t k = 0; k < edgeCount; ++k) {
        const std::uint32_t e = edges[k];
        if (g.isVirtual[e]) {
            leaf.addVirtualEdge(toLocal(g.ea[e]), toLocal(g.eb[e]));
        } else {
            const EdgeId le = leaf.getGraph().addEdge(toLocal(g.ea[e]), toLocal(g.eb[e])).value();
            if (g.constraint[e] && *g.constraint[e]) leaf.addConstraint(le, *g.constraint[e]);
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

    // The sketch as flat arrays: node slots = positions in ascending id order, edge slots likewise;
    // adjacency in compressed rows (edge slots ascending within a row, as SimpleGraph keeps its
    // incidence lists).  Peeling only ever removes: an edge dies (alive flag), its endpoints lose a
    // degree; the two live edges of a degree-2 node are found by scanning its row once.
    SketchView view;
    view.ids = graph.getNodes();   // ascending
    view.eids = graph.getEdges();  // ascending
    const auto& ids = view.ids;
    const auto& eids = view.eids;
    const std::size_t nn = ids.size(), ne = eids.size();
    auto slotOf = [&](NodeId n) -> std::size_t {
        if (n.value >= 0 && static_cast<std::size_t>(n.value) < nn && ids[static_cast<std::size_t>(n.value)] == n)
            return static_cast<std::size_t>(n.value);
        return static_cast<std::size_t>(std::lower_bound(ids.begin(), ids.end(), n) - ids.begin());
    };
    auto edgeSlotOf = [&](EdgeId e) -> std::size_t {
        if (e.value >= 0 && static_cast<std::size_t>(e.value) < ne && eids[static_cast<std::size_t>(e.value)] == e)
            return static_cast<std::size_t>(e.value);
        return static_cast<std::size_t>(std::lower_bound(eids.begin(), eids.end(), e) - eids.begin());
    };
    view.ea.resize(ne), view.eb.resize(ne);
    auto& ea = view.ea;  // endpoints of every edge, as node slots
    auto& eb = view.eb;
    view.element.assign(nn, nullptr), view.constraint.assign(ne, nullptr), view.isVirtual.assign(ne, 0);
    for (const auto& [node, el] : gcs.getElementMap()) {
        const std::size_t slot = slotOf(node);
        if (slot < nn && ids[slot] == node) view.element[slot] = &el;
    }
    for (const auto& [edge, con] : gcs.getConstraintMap()) {
        const std::size_t slot = edgeSlotOf(edge);
        if (slot < ne && eids[slot] == edge) view.constraint[slot] = &con;
    }
    for (EdgeId edge : gcs.getVirtualEdges()) {
        const std::size_t slot = edgeSlotOf(edge);
        if (slot < ne && eids[slot] == edge) view.isVirtual[slot] = 1;
    }
    std::vector<std::uint32_t> rowStart(nn + 1, 0);
    for (std::size_t k = 0; k < ne; ++k) {
        const auto [s, t] = graph.getEndpoints(eids[k]);
        ea[k] = static_cast<std::uint32_t>(slotOf(s)), eb[k] = static_cast<std::uint32_t>(slotOf(t));
        ++rowStart[ea[k] + 1];
        if (eb[k] != ea[k]) ++rowStart[eb[k] + 1];
    }
    for (std::size_t i = 0; i < nn; ++i) rowStart[i + 1] += rowStart[i];
    std::vector<std::uint32_t> row(rowStart[nn]), fill(rowStart.begin(), rowStart.end() - 1);
    for (std::size_t k = 0; k < ne; ++k) {
        row[fill[ea[k]]++] = static_cast<std::uint32_t>(k);
        if (eb[k] != ea[k]) row[fill[eb[k]]++] = static_cast<std::uint32_t>(k);
    }
    std::vector<std::uint32_t> degree(nn);
    for (std::size_t i = 0; i < nn; ++i) degree[i] = rowStart[i + 1] - rowStart[i];
    std::vector<char> alive(nn, 1), edgeAlive(ne, 1);
    auto otherEnd = [&](std::uint32_t e, std::size_t slot) { return ea[e] == slot ? eb[e] : ea[e]; };
    // the two live edges of a degree-2 node, ascending
    auto liveEdges = [&](std::size_t slot, std::uint32_t out[2]) {
        int found = 0;
        for (std::uint32_t q = rowStart[slot]; q < rowStart[slot + 1] && found < 2; ++q)
            if (edgeAlive[row[q]]) out[found++] = row[q];
    };
    // candidates (live nodes of degree two) in a MinSet: the peel always takes the smallest id
    MinSet candidates(nn);
    for (std::size_t i = 0; i < nn; ++i)
        if (degree[i] == 2) candidates.insert(i);
    std::vector<std::size_t> parked;  // degree-2 slots whose two edges form a double edge (skipped, kept)

    const auto t1 = std::chrono::steady_clock::now();
    std::vector<Peel> peels;
    peels.reserve(nn);
    std::size_t remaining = nn;
    while (remaining > 3) {
        // smallest-id degree-2 node whose two edges lead to two different neighbours
        bool found = false;
        std::size_t vs = 0;
        std::uint32_t e2[2] = {};
        while (!candidates.empty()) {
            const std::size_t c = candidates.min();
            candidates.erase(c);
            liveEdges(c, e2);
            if (otherEnd(e2[0], c) == otherEnd(e2[1], c)) {  // a double edge, not a separation pair
                parked.push_back(c);
                continue;
            }
            vs = c;
            found = true;
            break;
        }
        for (std::size_t c : parked) candidates.insert(c);  // they stay candidates for later rounds
        parked.clear();
        if (!found)
            throw std::runtime_error("decomposeByPeeling: no degree-2 element left with " + std::to_string(remaining)
                + " elements remaining; general separation pairs need the OGDF-based decomposition of the reference");
        const std::size_t sa = otherEnd(e2[0], vs), sb = otherEnd(e2[1], vs);
        peels.push_back({ static_cast<std::uint32_t>(vs), static_cast<std::uint32_t>(sa), static_cast<std::uint32_t>(sb), e2[0], e2[1] });
        edgeAlive[e2[0]] = edgeAlive[e2[1]] = 0;
        alive[vs] = 0;
        degree[vs] = 0;
        --remaining;
        for (std::size_t sl : { sa, sb }) {
            if (degree[sl] == 2) candidates.erase(sl);
            if (--degree[sl] == 2) candidates.insert(sl);
        }
    }
    const auto t2 = std::chrono::steady_clock::now();
    std::vector<ConstraintGraph> leaves(peels.size() + 1);
    {  // the base: the three remaining elements with every edge still alive between them, in id order
        std::array<std::uint32_t, 3> base {};
        std::set<std::uint32_t> edges;
        std::size_t found = 0;
        for (std::size_t i = 0; i < nn; ++i) {
            if (!alive[i]) continue;
            if (found < 3) base[found] = static_cast<std::uint32_t>(i);
            ++found;
            for (std::uint32_t q = rowStart[i]; q < rowStart[i + 1]; ++q)
                if (edgeAlive[row[q]]) edges.insert(row[q]);
        }
        if (found != 3) throw std::logic_error("decomposeByPeeling: the peel did not end at three elements");
        const std::vector<std::uint32_t> list(edges.begin(), edges.end());
        leaves[0] = makeLeaf(view, base, list.data(), list.size(), nullptr);
    }
    // the peeled leaves, in reverse peel order (= solve order).  Each is a small graph of its own
    // (eight allocations, shared_ptr copies of its elements and constraints) built from read-only
    // looks at the sketch: every host thread builds its share.  (Prefetching the elements and
    // constraints a few leaves ahead was measured and made this loop slower.)
    const long long np = static_cast<long long>(peels.size());
    std::exception_ptr failure;
#pragma omp parallel for schedule(static) if (np > 2048)
    for (long long k = 0; k < np; ++k) {
        const Peel& p = peels[static_cast<std::size_t>(np - 1 - k)];
        try {
            const std::pair<std::uint32_t, std::uint32_t> pair { p.a, p.b };
            const std::uint32_t edges[2] = { p.va, p.vb };
            leaves[static_cast<std::size_t>(k) + 1] = makeLeaf(view, { p.a, p.b, p.v }, edges, 2, &pair);
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
