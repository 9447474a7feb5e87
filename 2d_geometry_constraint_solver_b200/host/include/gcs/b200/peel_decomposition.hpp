// peel_decomposition.hpp — S-tree leaves of Henneberg-I style constraint graphs without OGDF.
//
// The reference decomposes top-down (stree_top_down_strategy.cpp:47-79): find a separation pair
// with OGDF, split into two graphs, give the side with the larger deficit a virtual edge between
// the pair and solve it AFTER the other side (right-first post-order), recurse.  OGDF is not in
// this image and which pair it reports first cannot be restated, so this is NOT a transliteration:
// it applies the reference's rules to the one family of separation pairs that can be found in
// linear time - the two neighbours {a, b} of a degree-2 node v.  There the split is
// {a, b, v} | rest; the triangle side holds the two real edges (a,v), (b,v), has deficit 1, gets the
// virtual edge (a,b) and is solved last; a real edge (a,b), if present, stays with the rest
// (gcs_data_structures.cpp:233-277: it must appear on exactly one side).  Graphs built by
// repeatedly hanging a new element on two placed ones (2-trees: BASELINE config 4's "rigidly
// well-constrained linkage") reduce completely this way; anything else throws.
// The leaf ORDER differs from the reference's wherever OGDF would have picked another pair
// first; the leaf SET and the virtual/real edge placement follow the same rules, and the wave
// scheduler only needs a valid order.
#pragma once

#include <vector>

#include <gcs/export.hpp>
#include <gcs/model/gcs_data_structures.hpp>

namespace Gcs::B200 {

struct PeelStats {
    std::size_t nodes = 0, edges = 0, leaves = 0;
};

// Leaves in solve order (base triangle first).  Throws std::runtime_error when fewer than 3
// nodes are present or no degree-2 node is left before the graph is down to 3 nodes.
GCS_API std::vector<ConstraintGraph> decomposeByPeeling(const ConstraintGraph& gcs, PeelStats* stats = nullptr);

}  // namespace Gcs::B200
