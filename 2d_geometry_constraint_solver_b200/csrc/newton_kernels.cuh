// newton_kernels.cuh — the two kernel variants of the batched Newton-Raphson path (sm_100a).
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
// Both produce bit-identical results (same device functions, same operation order per run).
#pragma once

#include <cuda_runtime.h>

#include "newton_core.cuh"

namespace gcsk {

struct BatchDev {
    const double* in[GCS_MAX_IN_COLS];
    const uint8_t* code;
    const double* guesses;  // [NS][2][n] or null
    double* out[GCS_MAX_OUT_COLS];
    double* cand;       // [NS][2][n] or null
    int16_t* iters;     // [NS][n] or null
    uint8_t* converged; // [NS][n] or null
    uint8_t* root;      // [n] or null
    long long n;
    long long stride;  // distance between per-seed planes (== n unless this launch is a slice of a larger batch)
};

constexpr unsigned kFull = 0xffffffffu;

// ------------------------------------------------------------------------------------------
// static variant
// ------------------------------------------------------------------------------------------
#ifndef GCS_STATIC_MINB
#define GCS_STATIC_MINB 1
#endif
template <int KIND, int NS>
__global__ void __launch_bounds__(128, GCS_STATIC_MINB) newton_static_kernel(const BatchDev p)
{
    using S = Sys<KIND>;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long sub = t / NS;
    const int seed = (int)(t % NS);
    const bool valid = sub < p.n;
    const long long i = valid ? sub : p.n - 1;  // clamp: every lane takes part in the shuffles

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
    FastConsts fc;
    fc.init((double)(p.n >> 62));  // 0.0 for every valid n, opaque to the compiler
    int it, conv;
    newton_run<KIND>(sys, fc, x, y, it, conv);

    // exchange candidates inside the NS-lane group
    const int lane = threadIdx.x & 31;
    const int lead = lane & ~(NS - 1);
    double cx[NS], cy[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        cx[s] = __shfl_sync(kFull, x, lead + s);
        cy[s] = __shfl_sync(kFull, y, lead + s);
    }
    if (valid) {
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
__device__ __forceinline__ unsigned lanemask_lt()
{
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
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
