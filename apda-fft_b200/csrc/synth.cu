// Device generator of the synthetic fleet windows (host twin: apda-fft_b200/synth.py).
// Bench / scale-test input only: parity tests always copy back the very buffer the kernels read.
#include "common.cuh"

namespace {

__device__ __forceinline__ uint64_t splitmix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

constexpr int kSeg = 32;  // consecutive samples per thread (one LCG jump-ahead each)

template <typename T>
__global__ void __launch_bounds__(256) synth_kernel(int64_t first, int64_t count, int N, uint64_t seed, int on_bin,
                                                    T *__restrict__ out) {
    const int segs = N / kSeg > 0 ? N / kSeg : 1;
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= count * segs) return;
    const int64_t wl = gid / segs;
    const int seg = (int)(gid % segs);
    const uint64_t w = (uint64_t)(first + wl);
    const uint64_t G = 0x9E3779B97F4A7C15ull;
    const uint64_t z0 = seed * G + w;
    double U[10];
#pragma unroll
    for (int q = 0; q < 10; ++q) U[q] = (double)(splitmix64(z0 + (uint64_t)q * G) >> 11) / 9007199254740992.0;
    double c[3] = {(0.025 + 0.010 * U[0]) * N, (0.055 + 0.015 * U[1]) * N, (0.095 + 0.020 * U[2]) * N};
    if (on_bin) {
        for (int t = 0; t < 3; ++t) c[t] = rint(c[t]);
    }
    const double a[3] = {0.5 * (0.9 + 0.2 * U[3]), 0.3 * (0.9 + 0.2 * U[4]), 0.2 * (0.9 + 0.2 * U[5])};
    const double two_pi = 2.0 * 3.141592653589793;
    const double ph[3] = {two_pi * U[6], two_pi * U[7], two_pi * U[8]};
    uint64_t s = (uint64_t)(U[9] * 9007199254740992.0) | 1ull;
    // jump the LCG ahead by start = seg*kSeg steps: s <- A^start * s + C*(A^start - 1)/(A - 1), by squaring
    const int start = seg * kSeg;
    {
        uint64_t accA = 1, accC = 0, curA = 6364136223846793005ull, curC = 1442695040888963407ull;
        for (int e = start; e; e >>= 1) {
            if (e & 1) {
                accA *= curA;
                accC = accC * curA + curC;
            }
            curC = (curA + 1) * curC;
            curA *= curA;
        }
        s = accA * s + accC;
    }
    const int end = min(N, start + kSeg);
    T *dst = out + wl * (int64_t)N;
    for (int i = start; i < end; ++i) {
        s = s * 6364136223846793005ull + 1442695040888963407ull;
        const double u = (double)(s >> 11) / 9007199254740992.0 * 2.0 - 1.0;
        double x = 0.0;
#pragma unroll
        for (int t = 0; t < 3; ++t) x = x + a[t] * sin(two_pi * c[t] * (double)i / (double)N + ph[t]);
        x = x + 0.01 * u;
        dst[i] = (T)(rint(x * 1e6) / 1e6);
    }
}

}  // namespace

template <typename T>
int launch_synth(apda_ctx *ctx, cudaStream_t st, int64_t first, int64_t count, int64_t N, uint64_t seed, int on_bin,
                 T *d_out) {
    if (count == 0) return APDA_OK;
    const int64_t segs = N / kSeg > 0 ? N / kSeg : 1;
    const int64_t threads = count * segs;
    const int64_t blocks = (threads + 255) / 256;
    if (blocks > 0x7fffffff) {
        apda_set_error("synth: too many windows for one launch");
        return APDA_ERR_INVALID;
    }
    synth_kernel<T><<<(unsigned)blocks, 256, 0, st>>>(first, count, (int)N, seed, on_bin, d_out);
    ctx->launches++;
    APDA_CUDA(cudaGetLastError());
    return APDA_OK;
}
template int launch_synth<double>(apda_ctx *, cudaStream_t, int64_t, int64_t, int64_t, uint64_t, int, double *);
template int launch_synth<float>(apda_ctx *, cudaStream_t, int64_t, int64_t, int64_t, uint64_t, int, float *);
