// newton_kernels.cuh — the kernel variants of the batched Newton-Raphson path (sm_100a).
//
//  newton_static_kernel : one lane per (sub-system, seed); the seeds of a sub-system sit in
//      adjacent lanes, candidates are exchanged with warp shuffles and the group leader applies
//      the orientation test and writes the result.  Handles every option of the ABI
//      (explicit guesses, 2 or 8 seeds, ragged sizes, unaligned columns).
//
//  newton_refill_kernel : persistent, warp-autonomous.  Each warp pulls chunks of CH
//      sub-systems from a global ticket counter; the chunk's input columns are brought into the
//      warp's shared-memory slab by TMA bulk copies (cp.async.bulk + mbarrier, double buffered:
//      chunk c+1 is in flight while chunk c is being iterated).  The CH*NS Newton runs of a
//      chunk are handed to lanes dynamically: a lane whose run converged takes the next run in
//      the same iteration (register-uniform counter + ballot, no atomics), so lanes do not idle
//      on the spread of iteration counts (11..18 for K1 from the +-20000 guesses).  Candidates
//      are parked in shared memory; after the chunk's runs drain, each lane selects the root for
//      its sub-systems and writes all outputs with coalesced stores.
//
//  newton_sorted_kernel : one CTA per tile; after three updates every run's remaining updates are
//      predicted from its step lengths, the tile's runs are sorted by the prediction and each
//      lane finishes one run, so the lanes of a warp stop (almost) together.  See the kernel.
//
// All produce bit-identical results (same device functions, same operation order per run).
//
//  The static and sorted kernels also exist in a contracted instantiation (template flag RLX):
//  closed-form updates on fused multiply-adds with guards (newton_relaxed.cuh) - iteration counts,
//  flags and root indices identical to the kernels above, coordinates within 1e-9 relative.
#pragma once

#include <cuda_runtime.h>

#include <cstring>

#include "newton_core.cuh"
#include "newton_relaxed.cuh"

namespace gcsk {

struct BatchDev {
    const double* in[GCS_MAX_IN_COLS];  // never null here: the ABI's NULL anchor columns arrive as the device's zero column
    const uint8_t* code;
    const double* guesses;  // [NS][2][n] or null
    double* out[GCS_MAX_OUT_COLS];
    double* cand;       // [NS][2][n] or null
    int16_t* iters;     // [NS][n] or null
    uint8_t* converged; // [NS][n] or null
    uint8_t* root;      // [n] or null
    uint8_t* path;      // [NS][n] or null (test hook, contracted static kernels): how each run was decided, kPath*
    long long n;
    long long stride;  // distance between per-seed planes (== n unless this launch is a slice of a larger batch)
    long long pf;      // sequential kernel: sub-systems one wave of resident CTAs covers (0: no look-ahead)
};

// Look-ahead of the sequential kernel: a CTA asks L2 for the columns the CTA one wave behind it will
// load (the 126 MB L2 holds whole batches), so that DRAM keeps streaming while the resident CTAs
// iterate instead of idling between one wave's load phase and the next one's.  Fire and forget: no
// register, no scoreboard.  Measured on K4 (two updates per seed: the memory-leaning kind): 30.7 us
// against 32.8 per 2^19, 170 against 184 per 2^22; nothing on the static kernel (K1, K2, K5: the
// loads of 8 resident CTAs already overlap the other CTAs' arithmetic), where it is not used.
__device__ __forceinline__ void prefetch_l2(const void* q)
{
    asm volatile("prefetch.global.L2 [%0];" ::"l"(q));
}

constexpr unsigned kFull = 0xffffffffu;

// BatchDev::path values (gcs_b200_debug_path_buffer)
enum : int {
    kPathFirstLevel = 0,   // closed form; every decision by the integer first-level tests
    kPathSecondLevel = 1,  // closed form; at least one decision by the second-level margin test
    kPathCareful = 2,      // closed form; careful mode (every late decision by the margin test)
    kPathLiteralRun = 3,   // redone by the literal code: a run-level guard fired
    kPathLiteralSel = 4    // redone by the literal code: the selection guard fired
};

__device__ __forceinline__ unsigned lanemask_lt()
{
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// ------------------------------------------------------------------------------------------
// static variant
// ------------------------------------------------------------------------------------------
#ifndef GCS_STATIC_MINB
#define GCS_STATIC_MINB 1
#endif
// RLX: closed-form updates with guards (newton_relaxed.cuh); a run or a selection the guards do
// not vouch for is redone with the literal device functions.
// RLX: held to 64 registers (8 CTAs of 128 lanes per SM).  The closed-form loop needs ~45; what is
// spilled belongs to the inlined literal re-run, which one run in a thousand takes.  Measured per
// 2^19 sub-systems, 5 / 6 / 8 / 10 / 12 CTAs per SM: K1 53.2 / 51.2 / 51.2 / 51.2 / 53.3 us,
// K5 51.2 / 45.1 / 44.0 / 43.0 / 45.0 us, K3 75.8 / 69.6 / 67.6 / 69.6 / 73.7 us.
#ifndef GCS_STATIC_RLX_MINB
#define GCS_STATIC_RLX_MINB 8
#endif
template <int KIND, int NS, bool RLX = false>
__global__ void __launch_bounds__(128, RLX ? GCS_STATIC_RLX_MINB : GCS_STATIC_MINB) newton_static_kernel(const BatchDev p)
{
    using S = Sys<KIND>;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long sub = t / NS;
    const int seed = (int)(t % NS);
    const bool valid = sub < p.n;
    const long long i = valid ? sub : p.n - 1;  // clamp: every lane takes part in the shuffles

    // (Loading the columns only the root selection reads - 8 of K5's 13 - late and in the leader lane
    // alone, with an L2 prefetch up front, was measured: K5 45.1 us against 43.0, K4 38.9 against 34.8.
    // The late loads sit on the critical path of a short-lived CTA; all columns are read up front.)
    double k[S::kCols];
#pragma unroll
    for (int c = 0; c < S::kCols; ++c) k[c] = __ldg(p.in[c] + i);
    const uint8_t code = p.code ? __ldg(p.code + i) : (uint8_t)GCS_MAKE_CODE(0, 0, 0);

    S sys;
    sys.load(k);
    double x, y;
    if (p.guesses) {
        x = __ldg(p.guesses + ((long long)seed * 2 + 0) * p.stride + i);
        y = __ldg(p.guesses + ((long long)seed * 2 + 1) * p.stride + i);
    } else if constexpr (S::kGuessFromCols) {
        column_seed<KIND>(k, seed, x, y);
    } else {
        default_seed(seed, x, y);
    }
    const double runtime_zero = (double)(p.n >> 62);  // 0.0 for every valid n, opaque to the compiler
    int it, conv;
    bool literal = !RLX;
    int pathv = 0;
    if constexpr (RLX) {
        RelaxedSystem<KIND> sysr;
        sysr.load(k);
        RelaxGuard g = sysr.g0;
        it = 0;
        int state = kRlxConverged;
        // iteration 0 compares the guess with prev = (0, 0)
        if (!(fabs(0.0 - x) < kTol && fabs(0.0 - y) < kTol)) {
            state = sysr.run(g, x, y, it, p.path ? &pathv : nullptr);
            if (state == kRlxWantCareful) {  // ill conditioned, above the floor: replay with every decision margin-tested
                run_seed<KIND>(p.guesses, p.stride, i, k, seed, x, y);
                const CarefulOut o = relaxed_careful<KIND>(sysr.rs, g, x, y);
                x = o.x, y = o.y, it = o.it, state = o.state, pathv |= o.trace;
            }
        }
        conv = 1;
        pathv = (pathv & 2) ? kPathCareful : (pathv & 1) ? kPathSecondLevel : kPathFirstLevel;
        if (state != kRlxConverged) {
            literal_rerun<KIND>(p.guesses, p.stride, i, k, seed, runtime_zero, x, y, it, conv, state - kRlxUncertain);
            literal = true;
            pathv = kPathLiteralRun;
        }
    } else {
        FastConsts fc;
        fc.init(runtime_zero);
        newton_run<KIND>(sys, fc, x, y, it, conv);
    }

    // exchange candidates inside the NS-lane group
    const int lane = threadIdx.x & 31;
    const int lead = lane & ~(NS - 1);
    double cx[NS], cy[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        cx[s] = __shfl_sync(kFull, x, lead + s);
        cy[s] = __shfl_sync(kFull, y, lead + s);
    }
    if constexpr (RLX) {
        // (G5) a selection the margins do not vouch for: the whole group goes literal
        int redo = (seed == 0 && !selection_is_robust<KIND, NS>(k, code, cx, cy)) ? 1 : 0;
        redo = __shfl_sync(kFull, redo, lead);
        if (__any_sync(kFull, redo)) {
            if (redo && !literal) {
                literal_rerun<KIND>(p.guesses, p.stride, i, k, seed, runtime_zero, x, y, it, conv, kWhySelection);
                pathv = kPathLiteralSel;
            }
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                cx[s] = __shfl_sync(kFull, x, lead + s);
                cy[s] = __shfl_sync(kFull, y, lead + s);
            }
        }
    }
    if (valid) {
        if constexpr (RLX)
            if (p.path) p.path[(long long)seed * p.stride + sub] = (uint8_t)pathv;
        if (p.iters) p.iters[(long long)seed * p.stride + sub] = (int16_t)it;
        if (p.converged) p.converged[(long long)seed * p.stride + sub] = (uint8_t)conv;
        if (p.cand) {
            p.cand[((long long)seed * 2 + 0) * p.stride + sub] = x;
            p.cand[((long long)seed * 2 + 1) * p.stride + sub] = y;
        }
        if (seed == 0) {
            double out[4];
            const int root = select_and_finish<KIND, NS>(k, code, cx, cy, out);
#pragma unroll
            for (int c = 0; c < S::kOut; ++c) p.out[c][sub] = out[c];
            if (p.root) p.root[sub] = (uint8_t)root;
        }
    }
}

// ------------------------------------------------------------------------------------------
// sequential variant: one lane per sub-system, its seeds one after the other
//
// The static mapping gives every (sub-system, seed) its own lane: the columns of a sub-system are
// requested by NS lanes, only one lane in NS selects the root, and a warp lasts as long as the
// longest of its 32 runs.  Here a lane owns a whole sub-system: it loads the columns once (32
// consecutive doubles per column and warp: full 256-byte requests), iterates seed 0, then seed 1
// (..., seed 7), selects and stores - every lane busy in every phase, no shuffles.  A warp now
// lasts as long as the largest SUM of iteration counts among its 32 sub-systems, and sums spread
// less than single runs do (K1: runs of 11..18 updates, sums of 24..30), so fewer lane-slots idle;
// for the short kinds (K4: two updates per seed; K2 / K5: five or six) the selection - a third to
// a half of the work - no longer runs on half-empty warps.  The price: half as many threads per
// batch (fewer CTAs to fill the last wave with) and twice the latency of a CTA.
// The literal re-run of the contracted class is an out-of-line call here (it reloads the columns
// itself), which keeps the closed-form path free of its registers.
// ------------------------------------------------------------------------------------------
struct RunOut {
    double x, y;
    int it, conv;
};

template <int KIND>
static __device__ __noinline__ RunOut literal_run_from_global(const BatchDev& p, long long i, int seed, int why)
{
    using S = Sys<KIND>;
    double k[S::kCols];
#pragma unroll
    for (int c = 0; c < S::kCols; ++c) k[c] = __ldg(p.in[c] + i);
    RunOut r;
    literal_rerun<KIND>(p.guesses, p.stride, i, k, seed, (double)(p.n >> 62), r.x, r.y, r.it, r.conv, why);
    return r;
}

// contracted: held to 64 registers (8 CTAs per SM).  With the Cramer form in the loop 80 registers
// (6 CTAs) were the faster choice (8 / 6 / 5 CTAs: K1 55.3 / 51.7 / 50.7 us, K3 80.4 / 73.7 / 74.8 us per
// 2^19); the line form's loop carries a fraction of that state, and the kernel - one dependent chain
// per lane - wants the warps: 8 / 6 CTAs: K1 x 8 seeds 228 / 235 us per 2^20, K1 225.8 / 244.6 us, K5
// 218.1 / 225.3 us per 2^22 with 2 seeds (10 CTAs = 48 registers: slower again)
#ifndef GCS_SEQ_MINB
#define GCS_SEQ_MINB 8
#endif
#ifndef GCS_SEQ_LIT_MINB
#define GCS_SEQ_LIT_MINB 5
#endif
template <int KIND, int NS, bool RLX>
__global__ void __launch_bounds__(128, RLX ? GCS_SEQ_MINB : GCS_SEQ_LIT_MINB) newton_seq_kernel(const __grid_constant__ BatchDev p)
{
    using S = Sys<KIND>;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n) return;
    double k[S::kCols];
#pragma unroll
    for (int c = 0; c < S::kCols; ++c) k[c] = __ldg(p.in[c] + i);
    const uint8_t code = p.code ? __ldg(p.code + i) : (uint8_t)GCS_MAKE_CODE(0, 0, 0);
    const double runtime_zero = (double)(p.n >> 62);
    if (p.pf > 0 && i + p.pf < p.n) {
#pragma unroll
        for (int c = 0; c < S::kCols; ++c) prefetch_l2(p.in[c] + i + p.pf);
    }
    double cx[NS], cy[NS];
    // contracted: what the seeds of a sub-system share (equations, guard, line constants) is set once
    std::conditional_t<RLX, RelaxedSystem<KIND>, int> sysr;
    if constexpr (RLX) sysr.load(k);

    auto one_seed = [&](int s) {
        double x, y;
        run_seed<KIND, (NS > 2)>(p.guesses, p.stride, i, k, s, x, y);
        int it = 0, conv = 1, pathv = 0;
        if constexpr (RLX) {
            RelaxGuard g = sysr.g0;
            int state = kRlxConverged;
            // iteration 0 compares the guess with prev = (0, 0)
            if (!(fabs(0.0 - x) < kTol && fabs(0.0 - y) < kTol)) {
                state = sysr.run(g, x, y, it, p.path ? &pathv : nullptr);
                if (state == kRlxWantCareful) {
                    run_seed<KIND, (NS > 2)>(p.guesses, p.stride, i, k, s, x, y);
                    const CarefulOut o = relaxed_careful<KIND>(sysr.rs, g, x, y);
                    x = o.x, y = o.y, it = o.it, state = o.state, pathv |= o.trace;
                }
            }
            pathv = (pathv & 2) ? kPathCareful : (pathv & 1) ? kPathSecondLevel : kPathFirstLevel;
            if (state != kRlxConverged) {
                const RunOut r = literal_run_from_global<KIND>(p, i, s, state - kRlxUncertain);
                x = r.x, y = r.y, it = r.it, conv = r.conv;
                pathv = kPathLiteralRun;
            }
            if (p.path) p.path[(long long)s * p.stride + i] = (uint8_t)pathv;
        } else {
            S sys;
            sys.load(k);
            FastConsts fc;
            fc.init(runtime_zero);
            newton_run<KIND>(sys, fc, x, y, it, conv);
        }
        if (p.iters) p.iters[(long long)s * p.stride + i] = (int16_t)it;
        if (p.converged) p.converged[(long long)s * p.stride + i] = (uint8_t)conv;
        cx[s] = x, cy[s] = y;
    };
    if constexpr (NS == 2) {
        one_seed(0);
        one_seed(1);
    } else {
#pragma unroll 1
        for (int s = 0; s < NS; ++s) one_seed(s);
    }
    if constexpr (RLX) {
        // (G5) a selection the margins do not vouch for: every seed of the sub-system goes literal
        if (!selection_is_robust<KIND, NS>(k, code, cx, cy)) {
#pragma unroll 1
            for (int s = 0; s < NS; ++s) {
                const RunOut r = literal_run_from_global<KIND>(p, i, s, kWhySelection);
                cx[s] = r.x, cy[s] = r.y;
                if (p.iters) p.iters[(long long)s * p.stride + i] = (int16_t)r.it;
                if (p.converged) p.converged[(long long)s * p.stride + i] = (uint8_t)r.conv;
                if (p.path) p.path[(long long)s * p.stride + i] = (uint8_t)kPathLiteralSel;
            }
        }
    }
    if (p.cand) {
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            p.cand[((long long)s * 2 + 0) * p.stride + i] = cx[s];
            p.cand[((long long)s * 2 + 1) * p.stride + i] = cy[s];
        }
    }
    double out[4];
    const int root = select_and_finish<KIND, NS>(k, code, cx, cy, out);
#pragma unroll
    for (int c = 0; c < S::kOut; ++c) p.out[c][i] = out[c];
    if (p.root) p.root[i] = (uint8_t)root;
}

// ------------------------------------------------------------------------------------------
// linear variant (K4, contracted class): one lane per sub-system, no iteration at all
//
// K4 is two point-to-line equations: a LINEAR pair with a constant Jacobian.  Newton's method lands
// on the solution P with its first update from any seed; the second update is the landing's own
// rounding error and ends the run: iters = 2, converged, candidate = P for EVERY seed.  Reproducing
// that literally costs two Householder-QR updates per seed plus the line-intersection frame of the
// selection (three IEEE divisions, a square root): ~900 warp instructions per 32 sub-systems, which
// makes the kind the survey expected to bind on HBM bind on instruction issue instead.
// The contract (iteration counts, flags, root identical; coordinates to 1e-9) needs none of it:
//   * P in closed form (Cramer on fused multiply-adds, origin moved to the first line's anchor);
//   * per seed, the two decisions certified instead of computed:
//       - the update from the seed is longer than the threshold: m1 = |P - seed|_max > 2 tol;
//       - the second update is shorter: it is the first update's rounding error, at most
//         ~ eps cond (m1 + S) in any backward-stable arithmetic (cond = |J|_F^2 / |det|, S the
//         coordinate scale; the contracted guards' carry / w1 terms budget 32 eps cond m1 for the
//         DIFFERENCE of two arithmetics there) - required: 2^-46 cond (m1 + S) < tol / 4, i.e. a
//         factor 512 between the bound and the threshold;
//       - a seed inside the iteration-0 box, or within 2 tol of P, is not certified;
//   * the root: every candidate is P and the orientation the reference tests is, exactly, the first
//     equation's signed distance s1 (P lies at signed distance s1 from line 1, and the frame's origin
//     is on line 1) - the reference computes it through the lines' intersection A, with an error of a
//     few ulp of |A| + |P|: required |s1| > 2^-24 (2 |A| + 2 + |P|), the margin of selection_is_robust;
//     nearest-to-canvas sub-systems (collinear / parallel codes, |cross| under the reference's absolute
//     epsilon) are not certified: their rule compares two noise-level distances;
//   * (G1) |det| >= 2^-10 |e1| |e2| as everywhere in the class, and cond < 2^16 (coordinates: the two
//     arithmetics' P differ by a few eps cond S); NaN / inf fail every comparison.
// A sub-system that is not certified - every seed of it - goes to the literal code (out of line, as
// in the sequential kernel), so degenerate inputs come out bit for bit.
// ------------------------------------------------------------------------------------------
// every seed of sub-system i by the literal code, then the literal selection (out of line: reloads the
// columns, keeps the certified path free of its registers and stack)
template <int NS>
static __device__ __noinline__ void linear_fallback(const BatchDev& p, long long i, uint8_t code)
{
    constexpr int KIND = GCS_KIND_PLL;
    using S = Sys<KIND>;
    double cx[NS], cy[NS];
#pragma unroll 1
    for (int s = 0; s < NS; ++s) {
        const RunOut r = literal_run_from_global<KIND>(p, i, s, kWhyCond);
        cx[s] = r.x, cy[s] = r.y;
        if (p.iters) p.iters[(long long)s * p.stride + i] = (int16_t)r.it;
        if (p.converged) p.converged[(long long)s * p.stride + i] = (uint8_t)r.conv;
        if (p.path) p.path[(long long)s * p.stride + i] = (uint8_t)kPathLiteralRun;
        if (p.cand) {
            p.cand[((long long)s * 2 + 0) * p.stride + i] = r.x;
            p.cand[((long long)s * 2 + 1) * p.stride + i] = r.y;
        }
    }
    double k[S::kCols];
#pragma unroll
    for (int c = 0; c < S::kCols; ++c) k[c] = __ldg(p.in[c] + i);
    double out[4];
    const int root = select_and_finish<KIND, NS>(k, code, cx, cy, out);
    p.out[0][i] = out[0], p.out[1][i] = out[1];
    if (p.root) p.root[i] = (uint8_t)root;
}

#ifndef GCS_LINEAR_MINB
#define GCS_LINEAR_MINB 8
#endif
template <int NS>
__global__ void __launch_bounds__(128, GCS_LINEAR_MINB) newton_linear_kernel(const __grid_constant__ BatchDev p)
{
    constexpr int KIND = GCS_KIND_PLL;
    using S = Sys<KIND>;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n) return;
    double k[S::kCols];
#pragma unroll
    for (int c = 0; c < S::kCols; ++c) k[c] = __ldg(p.in[c] + i);
    const uint8_t code = p.code ? __ldg(p.code + i) : (uint8_t)GCS_MAKE_CODE(0, 0, 0);
    if (p.pf > 0 && i + p.pf < p.n) {
#pragma unroll
        for (int c = 0; c < S::kCols; ++c) prefetch_l2(p.in[c] + i + p.pf);
    }
    const double e1x = k[2] - k[0], e1y = k[3] - k[1];
    const double e2x = k[7] - k[5], e2y = k[8] - k[6];
    const double q1 = __fma_rn(e1x, e1x, e1y * e1y), q2 = __fma_rn(e2x, e2x, e2y * e2y);
    const double r1 = rsqrt_relaxed(q1), r2 = rsqrt_relaxed(q2);
    const double len1 = q1 * r1, len2 = q2 * r2;
    const double cross_lit = e1x * e2y - e1y * e2x;  // heuristics.hpp:171-173, the reference's roundings
    const double det = __fma_rn(e1x, e2y, -(e1y * e2x));
    const double rdet = rcp_relaxed(det);
    // P - A1 from  -e1y x' + e1x y' = len1 s1,   -e2y x' + e2x y' = len2 s2 + (-e2y dlx + e2x dly),  dl = A2 - A1
    const double dlx = k[5] - k[0], dly = k[6] - k[1];
    const double c1 = len1 * k[4];
    const double c2 = __fma_rn(len2, k[9], __fma_rn(e2x, dly, -(e2y * dlx)));
    const double px = __fma_rn(__fma_rn(c1, e2x, -(e1x * c2)), rdet, k[0]);
    const double py = __fma_rn(__fma_rn(c1, e2y, -(e1y * c2)), rdet, k[1]);
    // the frame's origin A = A1 + t e1 (heuristics.hpp:165-181), for the margin of the orientation only
    const double t = __fma_rn(dlx, e2y, -(dly * e2x)) * rdet;
    const double sa = 2.0 * (fabs(__fma_rn(t, e1x, k[0])) + fabs(__fma_rn(t, e1y, k[1]))) + 2.0;
    const double cond = (q1 + q2) * fabs(rdet);
    const double scale = fabs(k[0]) + fabs(k[1]) + fabs(k[5]) + fabs(k[6]) + fabs(k[4]) + fabs(k[9]);
    bool ok = !(code & (GCS_CODE_COLLINEAR | GCS_CODE_CANVAS_PARALLEL)) && !(fabs(cross_lit) < GCS_PARALLEL_EPSILON)
        && fabs(det) >= 0x1p-10 * (len1 * len2) && fabs(k[4]) > 0x1p-24 * (sa + fabs(px) + fabs(py)) && cond < 0x1p16;
    const double nb = 0x1p-46 * cond;
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        double gx, gy;
        if (p.guesses) {
            gx = __ldg(p.guesses + ((long long)s * 2 + 0) * p.stride + i);
            gy = __ldg(p.guesses + ((long long)s * 2 + 1) * p.stride + i);
        } else {
            default_seed(s, gx, gy);
        }
        const double m1 = fmax(fabs(px - gx), fabs(py - gy));
        ok = ok && !(fabs(gx) < 2.0 * kTol && fabs(gy) < 2.0 * kTol) && m1 > 2.0 * kTol && nb * (m1 + scale) < 0.25 * kTol;
    }
    if (ok) {
        const int root = (GCS_CODE_SIGN0(code) == sgn3(k[4])) ? 0 : NS - 1;
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            if (p.iters) p.iters[(long long)s * p.stride + i] = (int16_t)2;
            if (p.converged) p.converged[(long long)s * p.stride + i] = (uint8_t)1;
            if (p.path) p.path[(long long)s * p.stride + i] = (uint8_t)kPathFirstLevel;
            if (p.cand) {
                p.cand[((long long)s * 2 + 0) * p.stride + i] = px;
                p.cand[((long long)s * 2 + 1) * p.stride + i] = py;
            }
        }
        p.out[0][i] = px, p.out[1][i] = py;
        if (p.root) p.root[i] = (uint8_t)root;
        return;
    }
    linear_fallback<NS>(p, i, code);  // not certified: the reference's arithmetic for every seed of the sub-system
}

// ------------------------------------------------------------------------------------------
// pair variant: one lane per sub-system, its two seeds iterated in lockstep (newton_run2)
// ------------------------------------------------------------------------------------------
#ifndef GCS_PAIR_MINB
#define GCS_PAIR_MINB 1
#endif
template <int KIND>
__global__ void __launch_bounds__(128, GCS_PAIR_MINB) newton_pair_kernel(const BatchDev p)
{
    using S = Sys<KIND>;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n) return;
    double k[S::kCols];
#pragma unroll
    for (int c = 0; c < S::kCols; ++c) k[c] = __ldg(p.in[c] + i);
    const uint8_t code = p.code ? __ldg(p.code + i) : (uint8_t)GCS_MAKE_CODE(0, 0, 0);
    S sys;
    sys.load(k);
    double cx[2], cy[2];
    if (p.guesses) {
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            cx[s] = __ldg(p.guesses + ((long long)s * 2 + 0) * p.stride + i);
            cy[s] = __ldg(p.guesses + ((long long)s * 2 + 1) * p.stride + i);
        }
    } else if constexpr (S::kGuessFromCols) {
        column_seed<KIND>(k, 0, cx[0], cy[0]);
        column_seed<KIND>(k, 1, cx[1], cy[1]);
    } else {
        default_seed(0, cx[0], cy[0]);
        default_seed(1, cx[1], cy[1]);
    }
    FastConsts fc;
    fc.init((double)(p.n >> 62));
    // iteration 0 compares the guess with prev = (0,0) (newton_raphson.hpp:58, :83-88)
    bool cva = fabs(0.0 - cx[0]) < fc.tol && fabs(0.0 - cy[0]) < fc.tol;
    bool cvb = fabs(0.0 - cx[1]) < fc.tol && fabs(0.0 - cy[1]) < fc.tol;
    int ita = 0, itb = 0;
    newton_run2<KIND>(sys, sys, fc, cx[0], cy[0], ita, cva, !cva, cx[1], cy[1], itb, cvb, !cvb);
    if (p.iters) p.iters[i] = (int16_t)ita, p.iters[p.stride + i] = (int16_t)itb;
    if (p.converged) p.converged[i] = (uint8_t)cva, p.converged[p.stride + i] = (uint8_t)cvb;
    if (p.cand) {
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            p.cand[((long long)s * 2 + 0) * p.stride + i] = cx[s];
            p.cand[((long long)s * 2 + 1) * p.stride + i] = cy[s];
        }
    }
    double out[4];
    const int root = select_and_finish<KIND, 2>(k, code, cx, cy, out);
#pragma unroll
    for (int c = 0; c < S::kOut; ++c) p.out[c][i] = out[c];
    if (p.root) p.root[i] = (uint8_t)root;
}

// ------------------------------------------------------------------------------------------
// sorted variant
//
// The static kernel loses lanes to the spread of iteration counts: a warp lasts as long as its
// slowest run (K1 from the default guesses: mean 12.5 updates, mean warp maximum 16.2; K5: 5.3 vs
// 10.3).  Re-packing the live runs after every update (tried: run state in shared memory, one
// barrier per update) costs more than it recovers in an issue-bound kernel.  This kernel instead
// PREDICTS each run's remaining updates once, sorts the runs of a CTA tile by the prediction and
// lets every lane finish one run in registers, so the 32 runs of a warp need (almost) the same
// number of updates and nothing is paid per update.
//
// Prediction.  Every equation pair of this path is one quadratic (a circle) plus an equation that
// is linear (K2-K5) or becomes linear after subtracting the two (K1).  Newton's method is affine
// covariant, so the first update lands on that line and from then on the run is the scalar
// iteration w <- (w^2 + h^2) / (2w) for the coordinate w along the line, measured from the foot
// point, with +-h the two roots.  In t = w/h = coth(theta) this is theta <- 2 theta: the step
// after j more updates is h / sinh(2^(j+1) theta).  Two consecutive step lengths identify the
// state: with s2, s3 the lengths of updates 2 and 3,  w2 = s2^2 / (2 s3),  h^2 = w2^2 - s2^2,
// w3 = w2 - s3,  theta = ln((w3 + h) / s3),  and the run needs
//         ceil(log2( ln(2h / tol) / theta ))
// more updates.  Measured on the bench inputs: exact for 99.6 % of the runs, off by one for the
// rest.  A wrong prediction costs time only: every run still iterates to the reference's own
// convergence test, so results are bit-identical to the other variants.
//
// Phases, per CTA tile: (A) lane = run does the iteration-0 test and up to three updates (hardly
// any run converges earlier, so nothing idles) and predicts; (B) counting sort of the live runs by
// predicted updates (32 bins, warp-aggregated shared-memory atomics); (C) every lane takes runs in
// sorted order and iterates each to the end in registers; (D) selection + write-back, coalesced.
// ------------------------------------------------------------------------------------------
constexpr int kSortBins = 32;

// Remaining updates of a run from the squared lengths of its 2nd and 3rd update (see above).
// d2 - 4 d3 is taken in FP64 (far from the root s3 ~ s2/2 and the difference carries h); the rest
// is FP32 on the SFU.  Garbage in (a run that is not on its line yet, overflow) gives some bin:
// harmless.
__device__ __forceinline__ int predict_remaining(double d2, double d3)
{
    const float t = (float)__fma_rn(-4.0, d3, d2);
    const float f2 = (float)d2, f3 = (float)d3;
    const float r3 = rsqrtf(f3);
    const float s3 = f3 * r3;         // |update 3|
    const float w2 = 0.5f * f2 * r3;  // s2^2 / (2 s3)
    const float hsq = 0.25f * (f2 * t) * (r3 * r3);
    const float h = hsq * rsqrtf(hsq);
    const float w3 = w2 - s3;
    const float theta = __log2f((w3 + h) * r3);
    const float span = __log2f(h * (float)(2.0 / kTol));
    const float k = ceilf(__log2f(__fdividef(span, theta)));
    int key = (k >= 1.0f) ? (int)fminf(k, (float)(kSortBins - 1)) : 1;
    if (!(hsq > 0.0f)) key = kSortBins - 1;  // no real root in sight: put it with the long runs
    return key;
}

#ifndef GCS_SORTED_THREADS
#define GCS_SORTED_THREADS 128
#endif
// CTAs per SM the register allocation is held to: 6 (80 registers) for the point kinds, 7 (72)
// for the line kinds K2/K5, whose runs are short (mean 5-6 updates) so that a larger share of a
// CTA's life is spent at barriers and in the load / selection phases, where more resident CTAs
// help (measured: K5 68.6 vs 71.7 us, K2 72.7 vs 75.2 us; K1 and K3 lose 2 % at 7).
#ifdef GCS_SORTED_MINB
template <int KIND>
constexpr int kSortedMinBlocks = GCS_SORTED_MINB;
#else
template <int KIND>
constexpr int kSortedMinBlocks = (KIND == GCS_KIND_SDD || KIND == GCS_KIND_ANG) ? 7 : 6;
#endif

// One CTA of THREADS lanes owns a tile of TILE sub-systems = RUNS = TILE*NS = 2*THREADS runs: two
// runs per lane in every phase, so a CTA is as short-lived as two static warps and the grid stays
// fine grained (many CTAs per SM slot), while the sort still sees 2*THREADS runs.  Inputs are not
// staged: phase A reads the columns coalesced, phases C and D read them again through L1/L2 (C: a
// gather inside the tile's 8*TILE bytes per column), which keeps shared memory at 22 bytes per run
// and the occupancy where the register file puts it.  In phase C the sorted runs form 2*W blocks
// of 32, longest first; warp w takes blocks w and 2W-1-w (a long one and a short one), so the
// warps of the CTA reach the last barrier at about the same time.
// (RLX: 8 CTAs per SM at 64 registers measured the same as 6-7 at 72-80 - the contracted kernels
// are bound by instruction issue, not by latency - so the allocation of the literal kernels is kept.)
template <int KIND, int NS, int TILE, int THREADS, bool RLX = false>
__global__ void __launch_bounds__(THREADS, kSortedMinBlocks<KIND>) newton_sorted_kernel(const BatchDev p)
{
    using S = Sys<KIND>;
    constexpr int RUNS = TILE * NS;
    constexpr int W = THREADS / 32;
    static_assert(RUNS == 2 * THREADS && THREADS % TILE == 0 && TILE % 32 == 0, "tile shape");
    __shared__ double s_x[RUNS], s_y[RUNS];  // run state, then the candidates; run = seed*TILE + sub
    __shared__ short s_it[RUNS];
    __shared__ unsigned short s_order[RUNS];
    __shared__ unsigned char s_cv[RUNS];
    __shared__ double s_carry[RLX ? RUNS : 1];  // RLX: each run's carry term (RelaxGuard::add_carry) at the hand-off
    // running minimum of hi(|det|) of each run at the hand-off (guard G2).  The growth value itself needs
    // no slot: growth inside phase A's stretch is judged when that stretch ends (relaxed_updates turns a
    // stretch with grow > kBounce into "uncertain"), and growth in phase C is measured against this
    // carried minimum, i.e. against the smallest |det| of the whole run so far.
    __shared__ int s_dmin[RLX ? RUNS : 1];
    __shared__ int s_bin[kSortBins];   // live runs per sort key
    __shared__ int s_fill[kSortBins];  // slots handed out per sort key during the scatter

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const long long base = (long long)blockIdx.x * TILE;
    const int cnt = (int)((p.n - base < TILE) ? (p.n - base) : TILE);
    FastConsts fc;
    fc.init((double)(p.n >> 62));  // 0.0 for every valid n, opaque to the compiler
    if (tid < kSortBins) s_bin[tid] = 0, s_fill[tid] = 0;
    __syncthreads();

    // ---- (A) iteration-0 test, three updates, prediction; this lane's runs: tid and tid + THREADS,
    //      two seeds of one sub-system ----
    int key[2] = { -1, -1 };  // sort key = kSortBins-1 - predicted updates; -1: finished / no run
    {
        const int sub = tid & (TILE - 1);
        if (sub < cnt) {
            const long long gi = base + sub;
            double k[S::kCols];
#pragma unroll
            for (int c = 0; c < S::kCols; ++c) k[c] = __ldg(p.in[c] + gi);
            if constexpr (RLX) {
                Rsys<KIND> rs;
                RelaxGuard g;
                rs.load(k, g);
#pragma unroll 1
                for (int q = 0; q < 2; ++q) {
                    const int r = tid + q * THREADS;
                    RelaxGuard gr = g;  // the run's own copy: its first update adds the carry term
                    double x, y;
                    run_seed<KIND>(p.guesses, p.stride, gi, k, r / TILE, x, y);
                    int it = 0, state = kRlxConverged;
                    double d2 = 0.0, d3 = 0.0;
                    int dmin = 0x7fffffff;
                    if (!(fabs(0.0 - x) < kTol && fabs(0.0 - y) < kTol))
                        state = relaxed_updates<KIND, true>(rs, gr, x, y, it, 3, d2, d3, &dmin);
                    s_dmin[r] = dmin;
                    s_carry[r] = gr.carry;
                    s_x[r] = x, s_y[r] = y, s_it[r] = (short)it;
                    s_cv[r] = (unsigned char)state;  // >= kRlxUncertain: phase C redoes the run literally
                    if (state != kRlxConverged) {
                        const int kk = rlx_uncertain(state) ? 0 : kSortBins - 1 - predict_remaining(d2, d3);
                        atomicAdd(&s_bin[kk], 1);
                        if (q == 0) key[0] = kk; else key[1] = kk;
                    }
                }
            } else {
                S sys;
                sys.load(k);
#pragma unroll 1
                for (int q = 0; q < 2; ++q) {
                    const int r = tid + q * THREADS;
                    const int seed = r / TILE;
                    double x, y;
                    if (p.guesses) {
                        x = __ldg(p.guesses + ((long long)seed * 2 + 0) * p.stride + gi);
                        y = __ldg(p.guesses + ((long long)seed * 2 + 1) * p.stride + gi);
                    } else if constexpr (S::kGuessFromCols) {
                        column_seed<KIND>(k, seed, x, y);
                    } else {
                        default_seed(seed, x, y);
                    }
                    // iteration 0 compares the guess with prev = (0,0) (newton_raphson.hpp:58, :83-88)
                    bool conv = fabs(0.0 - x) < fc.tol && fabs(0.0 - y) < fc.tol;
                    int it = 0;
                    double d2 = 0.0, d3 = 0.0;
#pragma unroll 1
                    for (int j = 0; j < 3 && !conv && it < kMaxIt; ++j) {
                        double nx, ny;
                        newton_update<KIND>(sys, fc, x, y, nx, ny, it);
                        const double ex = x - nx, ey = y - ny;
                        conv = fabs(ex) < fc.tol && fabs(ey) < fc.tol;
                        d2 = d3;
                        d3 = ex * ex + ey * ey;
                        x = nx, y = ny;
                    }
                    s_x[r] = x, s_y[r] = y, s_it[r] = (short)it;
                    s_cv[r] = (conv && it < kMaxIt) ? 1 : 0;
                    if (!conv && it < kMaxIt) {
                        const int kk = kSortBins - 1 - predict_remaining(d2, d3);
                        atomicAdd(&s_bin[kk], 1);
                        if (q == 0) key[0] = kk; else key[1] = kk;
                    }
                }
            }
        }
    }
    __syncthreads();
    // ---- (B) counting sort of the live runs, longest predicted first.  Every warp scans the 32
    //      bin counts for itself (lane k ends up with the first slot of key k), so no barrier is
    //      needed between the scan and the scatter; slots inside a key come from s_fill ----
    int n_live;
    {
        const int c = s_bin[lane];
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(kFull, incl, o);
            if (lane >= o) incl += v;
        }
        n_live = __shfl_sync(kFull, incl, 31);
        const int first = incl - c;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int kk = key[q];
            const int slot0 = __shfl_sync(kFull, first, kk & 31);
            const unsigned live = __ballot_sync(kFull, kk >= 0);
            if (kk >= 0) {
                const unsigned peers = __match_any_sync(live, kk);
                const int leader = __ffs(peers) - 1;
                int o = 0;
                if (lane == leader) o = atomicAdd(&s_fill[kk], __popc(peers));
                o = __shfl_sync(peers, o, leader);
                s_order[slot0 + o + __popc(peers & lanemask_lt())] = (unsigned short)(tid + q * THREADS);
            }
        }
    }
    __syncthreads();

    // ---- (C) the rest of every live run: warp w takes sorted blocks w and 2W-1-w ----
    {
        const int w = tid >> 5;
#pragma unroll 1
        for (int q = 0; q < 2; ++q) {
            const int j = (q == 0 ? w : 2 * W - 1 - w) * 32 + lane;
            if (j < n_live) {
                const int r = s_order[j];
                const int sub = r & (TILE - 1);
                double k[S::kCols];
#pragma unroll
                for (int c = 0; c < S::kCols; ++c) k[c] = __ldg(p.in[c] + base + sub);  // only what load() reads survives
                if constexpr (RLX) {
                    double x = s_x[r], y = s_y[r];
                    int it = s_it[r], conv = 1;
                    int state = s_cv[r];
                    if (!rlx_uncertain(state)) {
                        Rsys<KIND> rs;
                        RelaxGuard g;
                        rs.load(k, g);
                        g.add_carry(s_carry[r], Rsys<KIND>::kStepScale);
                        double u0, u1;
                        int dmin = s_dmin[r];
                        state = relaxed_updates<KIND, false>(rs, g, x, y, it, kRelaxCap, u0, u1, &dmin);
                        if (state == kRlxWantCareful) {  // replayed from the seed, whatever phase A did with its first updates
                            run_seed<KIND>(p.guesses, p.stride, base + sub, k, r / TILE, x, y);
                            const CarefulOut o = relaxed_careful<KIND>(rs, g, x, y);
                            x = o.x, y = o.y, it = o.it, state = o.state;
                        }
                    }
                    if (state != kRlxConverged)
                        literal_rerun<KIND>(p.guesses, p.stride, base + sub, k, r / TILE, (double)(p.n >> 62), x, y, it, conv, state - kRlxUncertain);
                    s_x[r] = x, s_y[r] = y, s_it[r] = (short)it;
                    s_cv[r] = (unsigned char)conv;
                } else {
                    S sys;
                    sys.load(k);
                    double x = s_x[r], y = s_y[r];
                    int it = s_it[r];
                    bool conv = false;
#pragma unroll 1
                    while (!conv && it < kMaxIt) {
                        double nx, ny;
                        newton_update<KIND>(sys, fc, x, y, nx, ny, it);
                        conv = fabs(x - nx) < fc.tol && fabs(y - ny) < fc.tol;
                        x = nx, y = ny;
                    }
                    s_x[r] = x, s_y[r] = y, s_it[r] = (short)it;
                    s_cv[r] = (conv && it < kMaxIt) ? 1 : 0;
                }
            }
        }
    }
    // the columns of this lane's phase-D sub-system are requested before the barrier, so that
    // their latency overlaps the wait for the slower warps of the CTA (TILE <= THREADS: one
    // sub-system per lane)
    static_assert(TILE <= THREADS, "phase D handles one sub-system per lane");
    double kd[S::kCols];
    uint8_t coded = (uint8_t)GCS_MAKE_CODE(0, 0, 0);
    if (tid < cnt) {
#pragma unroll
        for (int c = 0; c < S::kCols; ++c) kd[c] = __ldg(p.in[c] + base + tid);
        if (p.code) coded = __ldg(p.code + base + tid);
    }
    __syncthreads();

    // ---- (D) selection + write-back, coalesced, every lane busy ----
    if (tid < cnt) {
        const int sub = tid;
        const long long gi = base + sub;
        double* k = kd;
        const uint8_t code = coded;
        double cx[NS], cy[NS];
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            cx[s] = s_x[s * TILE + sub];
            cy[s] = s_y[s * TILE + sub];
        }
        if constexpr (RLX) {
            // (G5) a selection the margins do not vouch for: every seed of the sub-system goes literal
            if (!selection_is_robust<KIND, NS>(k, code, cx, cy)) {
#pragma unroll 1
                for (int s = 0; s < NS; ++s) {
                    int it, conv;
                    double x, y;
                    literal_rerun<KIND>(p.guesses, p.stride, gi, k, s, (double)(p.n >> 62), x, y, it, conv, kWhySelection);
                    s_x[s * TILE + sub] = x, s_y[s * TILE + sub] = y;
                    s_it[s * TILE + sub] = (short)it;
                    s_cv[s * TILE + sub] = (unsigned char)conv;
                }
#pragma unroll
                for (int s = 0; s < NS; ++s) {
                    cx[s] = s_x[s * TILE + sub];
                    cy[s] = s_y[s * TILE + sub];
                }
            }
        }
        double out[4];
        const int root = select_and_finish<KIND, NS>(k, code, cx, cy, out);
#pragma unroll
        for (int c = 0; c < S::kOut; ++c) p.out[c][gi] = out[c];
        if (p.root) p.root[gi] = (uint8_t)root;
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            const long long pi = (long long)s * p.stride + gi;
            if (p.iters) p.iters[pi] = s_it[s * TILE + sub];
            if (p.converged) p.converged[pi] = s_cv[s * TILE + sub];
            if (p.cand) {
                p.cand[((long long)s * 2 + 0) * p.stride + gi] = cx[s];
                p.cand[((long long)s * 2 + 1) * p.stride + gi] = cy[s];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + TMA bulk copy (global -> shared::cta)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ------------------------------------------------------------------------------------------
// refill variant
// ------------------------------------------------------------------------------------------
template <int KIND, int NS, int CH>
struct RefillSlab {
    // one warp's shared-memory slab
    double in[2][Sys<KIND>::kCols][CH];  // double-buffered input columns (TMA destination)
    double cx[NS][CH];
    double cy[NS][CH];
    int16_t it[NS][CH];
    uint8_t cv[NS][CH];
    uint8_t code[2][CH];
    uint64_t bar[2];
};

template <int KIND, int NS, int CH>
constexpr size_t refill_smem_bytes(int warps)
{
    return sizeof(RefillSlab<KIND, NS, CH>) * (size_t)warps + 16;
}

// tickets[0] = next chunk, tickets[1] = warps finished (the last one resets both)
template <int KIND, int NS, int CH, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
    newton_refill_kernel(const BatchDev p, unsigned* __restrict__ tickets, int aligned16)
{
    using S = Sys<KIND>;
    using Slab = RefillSlab<KIND, NS, CH>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    Slab& sl = *reinterpret_cast<Slab*>(smem_raw + sizeof(Slab) * warp);

    const long long nchunks = (p.n + CH - 1) / CH;
    const unsigned total_warps = gridDim.x * WARPS;

    if (lane == 0) {
        mbar_init(&sl.bar[0], 1);
        mbar_init(&sl.bar[1], 1);
    }
    // make the barrier initialisation visible to the async proxy
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();

    // issue the loads of one chunk into buffer `buf`
    auto issue = [&](long long chunk, int buf) {
        const long long base = chunk * CH;
        const int cnt = (int)((p.n - base < CH) ? (p.n - base) : CH);
        const bool bulk = aligned16 && (cnt == CH);
        if (bulk) {
            if (lane == 0) {
                fence_proxy_async();  // earlier generic reads of this buffer precede the async writes
                mbar_expect_tx(&sl.bar[buf], (uint32_t)(S::kCols * CH * 8 + CH));
#pragma unroll
                for (int c = 0; c < S::kCols; ++c)
                    tma_bulk_g2s(&sl.in[buf][c][0], p.in[c] + base, CH * 8, &sl.bar[buf]);
                tma_bulk_g2s(&sl.code[buf][0], p.code + base, CH, &sl.bar[buf]);
            }
        } else {
            for (int j = lane; j < cnt; j += 32) {
#pragma unroll
                for (int c = 0; c < S::kCols; ++c) sl.in[buf][c][j] = __ldg(p.in[c] + base + j);
                sl.code[buf][j] = p.code ? __ldg(p.code + base + j) : (uint8_t)GCS_MAKE_CODE(0, 0, 0);
            }
        }
        return bulk;
    };

    // ticket for the first chunk, and prefetch it
    long long cur = 0, nxt = 0;
    if (lane == 0) cur = atomicAdd(&tickets[0], 1u);
    cur = __shfl_sync(kFull, cur, 0);
    int buf = 0;
    unsigned phase[2] = { 0u, 0u };
    bool cur_bulk = false;
    if (cur < nchunks) cur_bulk = issue(cur, 0);

    while (cur < nchunks) {
        // next ticket + prefetch into the other buffer
        if (lane == 0) nxt = atomicAdd(&tickets[0], 1u);
        nxt = __shfl_sync(kFull, nxt, 0);
        bool nxt_bulk = false;
        if (nxt < nchunks) nxt_bulk = issue(nxt, buf ^ 1);

        const long long base = cur * CH;
        const int cnt = (int)((p.n - base < CH) ? (p.n - base) : CH);
        const int runs = cnt * NS;
        if (cur_bulk) {
            mbar_wait(&sl.bar[buf], phase[buf]);
            phase[buf] ^= 1u;
        }
        __syncwarp();

        // ---- Newton phase with lane refill ----
        int next = 0;  // warp-uniform
        FastConsts fc;
        fc.init((double)(p.n >> 62));
        bool active = false, cv = false;
        int slot_seed = 0, slot_sub = 0, it = 0;
        double x = 1.0, y = 1.0;
        S sys;
        {
            double z[S::kCols];
#pragma unroll
            for (int c = 0; c < S::kCols; ++c) z[c] = 1.0 + c;
            sys.load(z);
        }
#pragma unroll 1
        for (;;) {
            if (active && (cv || it >= kMaxIt)) {
                sl.cx[slot_seed][slot_sub] = x;
                sl.cy[slot_seed][slot_sub] = y;
                sl.it[slot_seed][slot_sub] = (int16_t)it;
                sl.cv[slot_seed][slot_sub] = (uint8_t)(cv && it < kMaxIt);
                active = false;
            }
            const unsigned need = __ballot_sync(kFull, !active);
            if (need) {
                if (next < runs) {
                    const int r = next + __popc(need & lanemask_lt());
                    next += __popc(need);
                    if (!active && r < runs) {
                        slot_seed = r / cnt;
                        slot_sub = r - slot_seed * cnt;
                        double k[S::kCols];
#pragma unroll
                        for (int c = 0; c < S::kCols; ++c) k[c] = sl.in[buf][c][slot_sub];
                        sys.load(k);
                        if (p.guesses) {
                            x = __ldg(p.guesses + ((long long)slot_seed * 2 + 0) * p.stride + base + slot_sub);
                            y = __ldg(p.guesses + ((long long)slot_seed * 2 + 1) * p.stride + base + slot_sub);
                        } else if constexpr (S::kGuessFromCols) {
                            column_seed<KIND>(k, slot_seed, x, y);
                        } else {
                            default_seed(slot_seed, x, y);
                        }
                        it = 0;
                        // iteration 0 of the reference loop compares the guess with prev = (0,0)
                        // (newton_raphson.hpp:58, :83-88): a guess inside the tolerance box leaves
                        // at once, unchanged
                        cv = fabs(0.0 - x) < kTol && fabs(0.0 - y) < kTol;
                        if (cv) {
                            sl.cx[slot_seed][slot_sub] = x;
                            sl.cy[slot_seed][slot_sub] = y;
                            sl.it[slot_seed][slot_sub] = 0;
                            sl.cv[slot_seed][slot_sub] = 1;
                        } else {
                            active = true;
                        }
                    }
                }
                if (!__any_sync(kFull, active)) {
                    if (next >= runs) break;
                    continue;  // every fresh run left at iteration 0; hand out the next ones
                }
            }
            double nx, ny;
            int it_next = it;
            newton_update<KIND>(sys, fc, x, y, nx, ny, it_next);
            if (active) {
                cv = fabs(x - nx) < kTol && fabs(y - ny) < kTol;
                x = nx, y = ny;
                it = it_next;
            }
        }
        __syncwarp();

        // ---- selection + write-back, coalesced ----
        for (int j = lane; j < cnt; j += 32) {
            double k[S::kCols];
#pragma unroll
            for (int c = 0; c < S::kCols; ++c) k[c] = sl.in[buf][c][j];
            const uint8_t code = sl.code[buf][j];
            double cx[NS], cy[NS];
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                cx[s] = sl.cx[s][j];
                cy[s] = sl.cy[s][j];
            }
            double out[4];
            const int root = select_and_finish<KIND, NS>(k, code, cx, cy, out);
            const long long gi = base + j;
#pragma unroll
            for (int c = 0; c < S::kOut; ++c) p.out[c][gi] = out[c];
            if (p.root) p.root[gi] = (uint8_t)root;
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                if (p.iters) p.iters[(long long)s * p.stride + gi] = sl.it[s][j];
                if (p.converged) p.converged[(long long)s * p.stride + gi] = sl.cv[s][j];
                if (p.cand) {
                    p.cand[((long long)s * 2 + 0) * p.stride + gi] = cx[s];
                    p.cand[((long long)s * 2 + 1) * p.stride + gi] = cy[s];
                }
            }
        }
        __syncwarp();

        cur = nxt;
        cur_bulk = nxt_bulk;
        buf ^= 1;
    }

    // self-cleaning tickets: the last warp to leave resets both counters for the next launch
    if (lane == 0) {
        __threadfence();
        const unsigned done = atomicAdd(&tickets[1], 1u);
        if (done == total_warps - 1) {
            tickets[0] = 0u;
            tickets[1] = 0u;
            __threadfence();
        }
    }
}

}  // namespace gcsk
