// Pass 2 of a Merge3 enumeration: every candidate is placed, merged and scored independently of the
// others, and the reference keeps the smallest score, the first one on ties (`score < bestScore` in
// enumeration order: merge3_ppp_solver.cpp:188, merge3_pll_solver.cpp:172, ...).  The per-candidate
// work - two rigid fits through a 2x2 SVD, a copy of the reference cluster, the score over every
// element - is ~4 us on one core and was all that was left of a merge once the Newton solves had
// moved into one launch; candidates are scored on every host thread here, and only the winner's
// pose is built a second time.  Same arithmetic per candidate, same winner: a first-minimum scan
// over the stored scores is the sequential loop's decision (a NaN score never wins in either).
#pragma once

#include <cstddef>
#include <exception>
#include <limits>
#include <optional>
#include <vector>

#include "solving/bottom_up/merge3_solver_common.hpp"

namespace Gcs::B200::detail {

// build(i): the merged pose of candidate i, or std::nullopt where the reference's loop `continue`s
template <class Build>
std::optional<Solvers::BottomUp::ClusterPose> pickBestMergedPose(
    const ConstraintGraph& sourceGraph, std::size_t n, Build&& build, std::size_t& scored, double& bestScore)
{
    std::vector<double> score(n, std::numeric_limits<double>::quiet_NaN());
    std::vector<unsigned char> placed(n, 0);
    std::exception_ptr failure;
#pragma omp parallel for schedule(dynamic, 8) if (n >= 64)
    for (long long i = 0; i < static_cast<long long>(n); ++i) {
        try {
            const auto merged = build(static_cast<std::size_t>(i));
            if (!merged) continue;
            placed[static_cast<std::size_t>(i)] = 1;
            score[static_cast<std::size_t>(i)] = Solvers::BottomUp::scoreMergedPose(sourceGraph, *merged);
        } catch (...) {
#pragma omp critical(gcs_merge3_pass2)
            if (!failure) failure = std::current_exception();
        }
    }
    if (failure) std::rethrow_exception(failure);
    scored = 0;
    bestScore = std::numeric_limits<double>::infinity();
    std::size_t best = n;
    for (std::size_t i = 0; i < n; ++i) {
        if (!placed[i]) continue;
        ++scored;
        if (score[i] < bestScore) bestScore = score[i], best = i;
    }
    if (best == n) return std::nullopt;
    return build(best);
}

}  // namespace Gcs::B200::detail
