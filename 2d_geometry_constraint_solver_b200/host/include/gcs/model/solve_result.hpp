// Outcome of a sub-problem solve (reference: includes/gcs/model/solve_result.hpp:14-54).
#pragma once

#include <string>
#include <utility>

#include <gcs/export.hpp>

namespace Gcs {

enum class SolveStatus { Success, Unsupported, Failed };

struct GCS_API SolveResult {
    SolveStatus status;
    std::string message;

    static SolveResult success() { return { SolveStatus::Success, {} }; }
    static SolveResult unsupported(std::string msg) { return { SolveStatus::Unsupported, std::move(msg) }; }
    static SolveResult failed(std::string msg) { return { SolveStatus::Failed, std::move(msg) }; }
};

}  // namespace Gcs
