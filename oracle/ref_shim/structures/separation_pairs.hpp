// Stand-in for <structures/separation_pairs.hpp>: the real header needs OGDF (absent).  The
// solver path never asks for separation pairs; isTriconnected is declared so that
// gcs_data_structures.hpp compiles and throws if anything calls it.
#pragma once
#include <expected>
#include <stdexcept>
#include <structures/graph.hpp>
#include <structures/graph_errors.hpp>
namespace MathUtils {
template <typename G>
std::expected<bool, GraphError> isTriconnected(const G&)
{
    throw std::runtime_error("isTriconnected: OGDF is not available in the reference-shim build");
}
}  // namespace MathUtils
