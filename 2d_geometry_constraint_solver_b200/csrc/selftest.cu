// selftest.cu — on-device self checks of the hand-written arithmetic (test hook of the C ABI).
//
// gcs_b200_selftest generates operand patterns on the device (splitmix64) and counts, entirely
// in-kernel, how often
//   [0] fast_div took its fast path           [1] ... and differed from the built-in a / b
//   [2] fast_sqrt took its fast path          [3] ... and differed from the built-in sqrt(a)
//   [4] qr_solve_fast accepted a 2x2 system   [5] ... and differed from the generic solver
//   [6] systems tried                         [7] operand pairs tried
// Counters [1], [3], [5] must be zero: the fast path is only ever allowed to return the bits the
// generic IEEE path returns.
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/gcs_b200.h"
#include "newton_core.cuh"

using namespace gcsk;

namespace {

__device__ __forceinline__ uint64_t mix64(uint64_t z)
{
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__device__ __forceinline__ bool same_bits(double a, double b)
{
    return __double_as_longlong(a) == __double_as_longlong(b) || (a != a && b != b);
}

// operand generator: class 0 = raw 64-bit patterns (all exponents, NaN, inf, subnormals),
// class 1 = moderate exponents with random mantissas, class 2 = mantissas of all ones / all
// zeros / single bits (the hard cases of correctly rounded division), class 3 = near 1 ulp pairs
__device__ double gen(uint64_t seed, uint64_t idx, int which)
{
    const uint64_t z = mix64(seed + (idx * 4 + which + 1) * 0x9E3779B97F4A7C15ull);
    const int cls = (int)(idx & 3);
    if (cls == 0) return __longlong_as_double((long long)z);
    const uint64_t sign = z & 0x8000000000000000ull;
    const int e = 1023 + (int)((z >> 52) % 1201) - 600;  // 2^-600 .. 2^600
    uint64_t m = z & 0x000fffffffffffffull;
    if (cls == 2) {
        const int pick = (int)((z >> 40) & 7);
        if (pick == 0) m = 0x000fffffffffffffull;
        else if (pick == 1) m = 0;
        else if (pick == 2) m = 1;
        else if (pick == 3) m = 0x000ffffffffffffeull;
        else if (pick == 4) m = 1ull << ((z >> 20) % 52);
        else if (pick == 5) m = 0x0008000000000000ull;
        else if (pick == 6) m = 0x000fffffffffffffull ^ (1ull << ((z >> 20) % 52));
    }
    return __longlong_as_double((long long)(sign | ((uint64_t)e << 52) | m));
}

__global__ void __launch_bounds__(256) selftest_kernel(uint64_t seed, long long n, unsigned long long* counts)
{
    unsigned long long c[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const double a = gen(seed, (uint64_t)i, 0), b = gen(seed, (uint64_t)i, 1);
        ++c[7];
        {
            bool ok = true;
            const double q = fast_div(a, b, ok);
            if (ok) {
                ++c[0];
                if (!same_bits(q, a / b)) ++c[1];
            }
        }
        {
            bool ok = true;
            const double s = fast_sqrt(fabs(a), ok);
            if (ok) {
                ++c[2];
                if (!same_bits(s, sqrt(fabs(a)))) ++c[3];
            }
        }
        {
            // 2x2 systems: entries drawn with a common random scale, sometimes nearly singular
            const double sc = gen(seed ^ 0x5a5a, (uint64_t)i | 1, 2);
            double m[6];
            for (int k = 0; k < 6; ++k) {
                const uint64_t z = mix64(seed + ((uint64_t)i * 8 + k + 17) * 0x9E3779B97F4A7C15ull);
                m[k] = ((double)(long long)(z >> 11) * (1.0 / 9007199254740992.0) - 0.5) * sc;
            }
            if ((i & 15) == 3) m[1] = m[0] * (1.0 + 1e-9), m[3] = m[2] * (1.0 + 1e-9);  // near-parallel columns
            if ((i & 15) == 7) m[2] = 0.0;                                            // triangular
            if ((i & 15) == 11) m[4] = 0.0;                                           // zero rhs entry
            double f0, f1, g0, g1;
            ++c[6];
            if (qr_solve_fast(m[0], m[1], m[2], m[3], m[4], m[5], f0, f1)) {
                ++c[4];
                const double2 g = qr_solve_generic(m[0], m[1], m[2], m[3], m[4], m[5]);
                g0 = g.x, g1 = g.y;
                if (!same_bits(f0, g0) || !same_bits(f1, g1)) ++c[5];
            }
        }
    }
    for (int k = 0; k < 8; ++k) {
        unsigned long long v = c[k];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(&counts[k], v);
    }
}

}  // namespace

extern "C" int gcs_b200_selftest_launch(uint64_t seed, long long n, unsigned long long* dev_counts, int sm_count)
{
    selftest_kernel<<<sm_count * 8, 256>>>(seed, n, dev_counts);
    return cudaGetLastError() == cudaSuccess ? 0 : -3;
}
