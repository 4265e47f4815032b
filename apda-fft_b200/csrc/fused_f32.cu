// Fused window -> record kernel (fp32, N = 1024..8192): SURVEY 8(f) rank 1, the "B_min" variant.
//
// K1-fast's three register passes, then the split step keeps only the half spectrum the pickers read: bins k < N/2 go
// straight from registers to magnitudes in shared memory and the window's Sum / Sum-of-squares are reduced across the
// CTA; the picker tail then runs on ALL the CTA's warps (k3_tail_mw in peaks_fast.cuh: a 16-bin chunk per thread, the
// candidates' prominence walks dealt out to the warps), so no FFT warp idles while the record is being made.  HBM traffic per window is
// s*N bytes in and 128 bytes out (16 512 B at N = 4096) instead of the pipeline's 4*s*N + 128: the spectrum never exists
// in memory.  This is a throughput variant: the drop-in start_fft contract (N bins materialised) stays with the
// pipeline kernels, and results are reported under their own byte accounting (never mixed with B_alg numbers).
#include <mutex>

#include "fft_f32_fast.cuh"
#include "peaks_fast.cuh"

int fft_f32_fast_get_tables(apda_ctx *ctx, int64_t N, const float2 **tw1, const float2 **twu);

namespace {

template <int N>
struct FusedLayout {
    using P = Plan<N>;
    static constexpr int M = N / 2, T = M / 16;
    static constexpr int FFT_BYTES = (int)sizeof(float2) * P::R1 * (P::R2 * P::R3 + 16 / P::R1);
    static constexpr int MAG_BYTES = K3M<M>::MAGW * (int)sizeof(float);
    static constexpr int BYTES = FFT_BYTES + MAG_BYTES + 128;
    static constexpr int SLOT_CAP = M / 5 + 8;  // bins above mean + 2 sigma are < 20 % of all bins: never overflows
    // the FFT buffer is free once the magnitudes exist: chunk summaries and the slot list live there
    static_assert(2 * T * (int)sizeof(float) + SLOT_CAP * (int)sizeof(Slot) <= FFT_BYTES, "tail scratch must fit the FFT buffer");
};

#ifndef APDA_FUSED_MINB
#define APDA_FUSED_MINB 8  // resident CTAs per SM the N = 4096 instance is compiled for
#endif
template <int N, int CENTER, bool FULL, bool FLEX>
__global__ void __launch_bounds__(N / 32, (N == 4096 ? APDA_FUSED_MINB : N < 4096 ? 1024 / (N / 32) / 2 : 2))
fused_f32_kernel(const float *__restrict__ samples, int n_samples, int64_t ld, int64_t batch,
                 const float2 *__restrict__ tw1, const float2 *__restrict__ twu, double df_all,
                 const double *__restrict__ d_fs, int k, unsigned char *__restrict__ recs) {
    using L = FusedLayout<N>;
    using Q = K3M<N / 2>;
    constexpr int M = N / 2, T = M / 16;
    extern __shared__ __align__(16) unsigned char dyn_smem[];  // FFT buffer, magnitudes, record
    float2 *s = reinterpret_cast<float2 *>(dyn_smem);
    float *mags = reinterpret_cast<float *>(dyn_smem + L::FFT_BYTES);
    unsigned char *rec_s = dyn_smem + L::FFT_BYTES + L::MAG_BYTES;
    __shared__ uint32_t sel[64];
    __shared__ float red[8];
    __shared__ double stat[2 * (T / 32 > 0 ? T / 32 : 1)];
    __shared__ int ctl[2];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int64_t win = blockIdx.x;
    if (win >= batch) return;
    if (t == 0) ctl[0] = ctl[1] = 0;  // slot count, tie flag (published by the barriers inside k1_forward)

    k1_forward<N, CENTER, FULL, (N == 4096 ? APDA_FUSED_MINB : N < 4096 ? 1024 / (N / 32) / 2 : 2)>(
        samples, n_samples, ld, win, tw1, s, sel, red, 0, t, batch);

    // split step, lower half only: X[k] = S/2 + Wt*D for k = 2p, 2p+1 -> magnitudes in the tail's shared-memory layout
    if (t < 16) reinterpret_cast<uint64_t *>(rec_s)[t] = (t % 3 == 1) ? 0x00000000ffffffffull : 0ull;
    const float4 *s4 = reinterpret_cast<const float4 *>(s);
    const float4 *twu4 = reinterpret_cast<const float4 *>(twu);
    float sum = 0.f, sumsq = 0.f;
#pragma unroll
    for (int g = 0; g < (M / 2) / T; ++g) {
        const int p = t + T * g;
        const float4 zk = s4[p];
        const float2 za = s[(M - 2 * p) & (M - 1)];
        const float2 zb = s[M - 2 * p - 1];
        const float4 w = __ldg(twu4 + p);
        const float2 cj = make_float2(1.f, -1.f), ncj = make_float2(-1.f, 1.f), hf = make_float2(0.5f, 0.5f);
        const float2 zk0 = make_float2(zk.x, zk.y), zk1 = make_float2(zk.z, zk.w);
        const float2 s0 = pfma(za, cj, zk0), d0 = pfma(za, ncj, zk0);
        const float2 s1 = pfma(zb, cj, zk1), d1 = pfma(zb, ncj, zk1);
        const float2 t0 = cmul(d0, make_float2(w.x, w.y)), t1 = cmul(d1, make_float2(w.z, w.w));
        float2 x0 = pfma(s0, hf, t0);
        const float2 x1 = pfma(s1, hf, t1);
        if (p == 0) x0 = make_float2(0.f, 0.f);  // reference: res[0] = 0
        const float p0 = fmaf(x0.x, x0.x, x0.y * x0.y), p1 = fmaf(x1.x, x1.x, x1.y * x1.y);
        const float m0 = sqrt_fast(p0), m1 = sqrt_fast(p1);
        sum += m0 + m1;
        sumsq += p0 + p1;
        *reinterpret_cast<float2 *>(mags + Q::addr(2 * p)) = make_float2(m0, m1);
    }
    double S = (double)sum, Qs = (double)sumsq;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        S += __shfl_xor_sync(0xffffffffu, S, o);
        Qs += __shfl_xor_sync(0xffffffffu, Qs, o);
    }
    constexpr int NW = T / 32 > 0 ? T / 32 : 1;
    if (lane == 0) {
        stat[warp] = S;
        stat[NW + warp] = Qs;
    }
    group_sync<T>(0);  // magnitudes and partial sums complete; the FFT buffer is free
    S = 0.0;
    Qs = 0.0;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
        S += stat[w];
        Qs += stat[NW + w];
    }
    const double nn = (double)M;
    const double mean = S / nn;
    const double var = (Qs - S * S / nn) / (nn - 1.0);
    const double sd = var > 0.0 ? sqrt(var) : 0.0;
    const double thr = mean + 2.0 * sd;
    const float thr_f = __double2float_rd(thr);
    const double df = d_fs ? div_rn(d_fs[win], (double)N) : df_all;
    float *cmaxs = reinterpret_cast<float *>(s), *cmins = cmaxs + T;
    Slot *slots = reinterpret_cast<Slot *>(cmins + T);
    k3_tail_mw<M, FLEX, T>(mags, cmaxs, cmins, slots, L::SLOT_CAP, ctl, rec_s, sd, thr_f, df, k, t, win, recs);
}

// this translation unit owns its own copy of the __constant__ pass-2 twiddles (anonymous namespace in the header)
// (constant memory is per device: uploaded once per device the library is used on)
template <int N>
int upload_tw2(int device) {
    static std::vector<int> done;
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    for (int d : done)
        if (d == device) return APDA_OK;
    using P = Plan<N>;
    constexpr int R2 = P::R2, R3 = P::R3;
    const double two_pi = 6.283185307179586476925286766559;
    std::vector<float2> h2((size_t)R2 * R3);
    for (int k2 = 0; k2 < R2; ++k2)
        for (int n3 = 0; n3 < R3; ++n3) {
            const double a = -two_pi * (double)(k2 * n3) / (double)(R2 * R3);
            h2[(size_t)k2 * R3 + n3] = make_float2((float)cos(a), (float)sin(a));
        }
    const void *sym = N == 1024 ? (const void *)c_tw2_1024 : N == 2048 ? (const void *)c_tw2_2048
                    : N == 4096 ? (const void *)c_tw2_4096 : (const void *)c_tw2_8192;
    APDA_CUDA(cudaMemcpyToSymbol(sym, h2.data(), h2.size() * sizeof(float2)));
    done.push_back(device);
    return APDA_OK;
}

template <int N>
int launch_fused_n(apda_ctx *ctx, cudaStream_t st, const float *d_samples, int64_t n_samples, int64_t ld, int64_t batch,
                   int flags, int flexible, double fs, const double *d_fs, int k, void *d_rec) {
    const float2 *tw1, *twu;
    APDA_TRY(fft_f32_fast_get_tables(ctx, N, &tw1, &twu));
    APDA_TRY(upload_tw2<N>(ctx->device));
    const bool full = n_samples == N && (reinterpret_cast<uintptr_t>(d_samples) & 7u) == 0 && (ld & 1) == 0;
    const bool med = flags == APDA_CENTER_MEDIAN;
    void (*kern)(const float *, int, int64_t, int64_t, const float2 *, const float2 *, double, const double *, int,
                 unsigned char *);
#define PICK(C, F, X) fused_f32_kernel<N, C, F, X>
    if (flexible) {
        kern = med ? (full ? PICK(APDA_CENTER_MEDIAN, true, true) : PICK(APDA_CENTER_MEDIAN, false, true))
                   : (full ? PICK(APDA_CENTER_MEAN, true, true) : PICK(APDA_CENTER_MEAN, false, true));
    } else {
        kern = med ? (full ? PICK(APDA_CENTER_MEDIAN, true, false) : PICK(APDA_CENTER_MEDIAN, false, false))
                   : (full ? PICK(APDA_CENTER_MEAN, true, false) : PICK(APDA_CENTER_MEAN, false, false));
    }
#undef PICK
    const size_t smem = FusedLayout<N>::BYTES;
    APDA_FUNC_SMEM(ctx, kern, smem);
    kern<<<(unsigned)batch, N / 32, smem, st>>>(d_samples, (int)n_samples, ld, batch, tw1, twu, fs / (double)N, d_fs, k,
                                            reinterpret_cast<unsigned char *>(d_rec));
    ctx->launches++;
    APDA_CUDA(cudaGetLastError());
    return APDA_OK;
}

}  // namespace

int launch_fused_f32(apda_ctx *ctx, cudaStream_t st, const float *d_samples, int64_t n_samples, int64_t ld, int64_t batch,
                     int64_t N, int flags, int flexible, double fs, const double *d_fs, int k, void *d_rec) {
    switch (N) {
        case 1024: return launch_fused_n<1024>(ctx, st, d_samples, n_samples, ld, batch, flags, flexible, fs, d_fs, k, d_rec);
        case 2048: return launch_fused_n<2048>(ctx, st, d_samples, n_samples, ld, batch, flags, flexible, fs, d_fs, k, d_rec);
        case 4096: return launch_fused_n<4096>(ctx, st, d_samples, n_samples, ld, batch, flags, flexible, fs, d_fs, k, d_rec);
        case 8192: return launch_fused_n<8192>(ctx, st, d_samples, n_samples, ld, batch, flags, flexible, fs, d_fs, k, d_rec);
    }
    apda_set_error("fused_f32: unsupported N=%lld", (long long)N);
    return APDA_ERR_UNSUPPORTED;
}
