// synth.cu — on-device generator of synthetic K1 (point-point-point triangle cluster) instances
// for BASELINE.json configs 2, 3 and 5.  Mirrors 2d_geometry_constraint_solver_b200/synth.py
// `make_pp` operation by operation (splitmix64 counter RNG, only + - * / sqrt, --fmad=false), so
// host and device streams agree bit for bit; tests/test_synth.py checks that.
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/gcs_b200.h"

namespace {

__device__ __forceinline__ uint64_t mix64(uint64_t z)
{
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__device__ __forceinline__ double uni(uint64_t seed, uint64_t idx, int field)
{
    const uint64_t ctr = idx * 32ull + (uint64_t)(field + 1);
    const uint64_t z = mix64(seed + ctr * 0x9E3779B97F4A7C15ull);
    return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}

struct Frame {
    double c, s, tx, ty;
};

__device__ __forceinline__ Frame frame(uint64_t seed, uint64_t idx, int f0)
{
    double t = 2.0 * uni(seed, idx, f0) - 1.0;
    t = t / (1.0 - fabs(t) * 0.999);
    const double den = 1.0 + t * t;
    Frame f;
    f.c = (1.0 - t * t) / den;
    f.s = (2.0 * t) / den;
    f.tx = 1000.0 * uni(seed, idx, f0 + 1);
    f.ty = 1000.0 * uni(seed, idx, f0 + 2);
    return f;
}

__device__ __forceinline__ void apply(const Frame& f, double x, double y, double& ox, double& oy)
{
    ox = (f.c * x - f.s * y) + f.tx;
    oy = (f.s * x + f.c * y) + f.ty;
}

__global__ void __launch_bounds__(256) synth_pp_kernel(uint64_t seed, long long first, long long n,
    int perturb_of, double* ax_, double* ay_, double* ra_, double* bx_, double* by_, double* rb_,
    uint8_t* code_)
{
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const uint64_t idx = (uint64_t)(first + j);
    const uint64_t base = perturb_of ? idx % (uint64_t)perturb_of : idx;
    double d = (10.0 + 490.0 * uni(seed, base, 0)) * 1.0;
    const double px = (-300.0 + 1100.0 * uni(seed, base, 1)) * 1.0;
    const double pym = (5.0 + 495.0 * uni(seed, base, 2)) * 1.0;
    const double side = uni(seed, base, 3) < 0.5 ? 1.0 : -1.0;
    const double py = side * pym;
    double ra = sqrt(px * px + py * py);
    double rb = sqrt((px - d) * (px - d) + py * py);
    if (perturb_of) {
        const uint64_t ps = seed ^ 0xABCDEFull;
        for (int attempt = 0; attempt < 8; ++attempt) {
            const int f = 4 + 3 * attempt;
            const double pa = ra * (1.0 + 0.05 * (2.0 * uni(ps, idx, f) - 1.0));
            const double pb = rb * (1.0 + 0.05 * (2.0 * uni(ps, idx, f + 1) - 1.0));
            const double pd = d * (1.0 + 0.05 * (2.0 * uni(ps, idx, f + 2) - 1.0));
            if ((pa + pb > pd * 1.001) && (fabs(pa - pb) < pd * 0.999)) {
                ra = pa, rb = pb, d = pd;
                break;
            }
        }
    }
    // canvas layout: the true triple moved rigidly -> orientation sign
    const Frame cf = frame(seed, base, 28);
    double cax, cay, cbx, cby, cpx, cpy;
    apply(cf, 0.0, 0.0, cax, cay);
    apply(cf, d, 0.0, cbx, cby);
    apply(cf, px, py, cpx, cpy);
    const double ori = ((cbx - cax) * (cpy - cay)) - ((cby - cay) * (cpx - cax));
    const int sign = (ori > 0.0) - (ori < 0.0);
    // odd instances: general fixed positions
    double ax = 0.0, ay = 0.0, bx = d, by = 0.0;
    if (base & 1ull) {
        const Frame sf = frame(seed, base, 24);
        apply(sf, 0.0, 0.0, ax, ay);
        apply(sf, d, 0.0, bx, by);
    }
    ax_[j] = ax, ay_[j] = ay, ra_[j] = ra;
    bx_[j] = bx, by_[j] = by, rb_[j] = rb;
    code_[j] = GCS_MAKE_CODE(sign, 0, 0);
}

}  // namespace

extern "C" int gcs_b200_synth_pp_launch(void* cuda_stream, uint64_t seed, int64_t first, int64_t n,
    int perturb_of, double* const cols[6], uint8_t* code)
{
    if (n <= 0) return 0;
    const int block = 256;
    const long long grid = (n + block - 1) / block;
    synth_pp_kernel<<<(unsigned)grid, block, 0, static_cast<cudaStream_t>(cuda_stream)>>>(seed,
        (long long)first, (long long)n, perturb_of, cols[0], cols[1], cols[2], cols[3], cols[4],
        cols[5], code);
    return cudaGetLastError() == cudaSuccess ? 0 : -3;
}
