// Equations::solve2D on the device: a batch of one through the C ABI (reference:
// src/constraint_solver/src/solving/equations/newton_raphson.hpp:41-102).
#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <string>

#include "gcs_b200.h"
#include "solving/equations/newton_raphson.hpp"

namespace Gcs::Equations::detail {

std::array<Eigen::Vector2d, 2> solveOnDevice(int kind, const double* cols, const std::array<Eigen::Vector2d, 2>& guesses,
    Solve2DInfo* info, int device)
{
    gcs_b200_batch b {};
    b.kind = kind;
    b.n_seeds = 2;
    b.n = 1;
    b.mem = GCS_MEM_HOST;
    b.variant = GCS_VARIANT_STATIC;
    const int nin = gcs_b200_kind_in_cols(kind), nout = gcs_b200_kind_out_cols(kind);
    for (int c = 0; c < nin; ++c) b.in[c] = cols + c;
    const std::uint8_t code = GCS_MAKE_CODE(0, 0, 0);
    b.code = &code;
    const double g[4] = { guesses[0].x(), guesses[0].y(), guesses[1].x(), guesses[1].y() };  // [seed][xy][n=1]
    b.guesses = g;
    double out[GCS_MAX_OUT_COLS] = {};
    for (int c = 0; c < nout; ++c) b.out[c] = out + c;
    double cand[4] = {};
    std::int16_t iters[2] = {};
    std::uint8_t conv[2] = {}, root = 0;
    b.cand = cand, b.iters = iters, b.converged = conv, b.root_index = &root;
    const int rc = gcs_b200_solve_host(&b, device);
    if (rc != GCS_OK)
        throw std::runtime_error(std::string("Equations::solve2D: gcs_b200_solve_host failed (") + std::to_string(rc)
            + "): " + gcs_b200_last_error());
    if (info) {
        for (int s = 0; s < 2; ++s) info->iterations[s] = iters[s], info->converged[s] = conv[s] != 0;
    }
    return { Eigen::Vector2d { cand[0], cand[1] }, Eigen::Vector2d { cand[2], cand[3] } };
}

void requireLength(double ex, double ey, double length, const char* what)
{
    const double l = std::sqrt(ex * ex + ey * ey);
    if (!(l == length))
        throw std::invalid_argument(std::string(what) + ": the length argument must be the Euclidean length of the "
            "direction it belongs to (every reference call site passes exactly that); the kernel recomputes it");
}

}  // namespace Gcs::Equations::detail
