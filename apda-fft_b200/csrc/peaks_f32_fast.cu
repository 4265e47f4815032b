// K3 (fp32 fast form): the pipeline kernel (spectrum in HBM -> records).  The algorithm lives in peaks_fast.cuh.
#include "peaks_fast.cuh"

namespace {

#ifndef APDA_K3_BATCH
#define APDA_K3_BATCH 8
#endif
constexpr int kK3Batch = APDA_K3_BATCH;
#ifndef APDA_K3_PF_CTAS
#define APDA_K3_PF_CTAS 6  // prefetch distance in CTAs per SM (11 are resident at N = 4096; sweep: 6 -> 2.49 ns, 11 -> 2.52, 20 -> 2.93)
#endif
#ifndef APDA_K3_PF_CTAS_LONG
#define APDA_K3_PF_CTAS_LONG 6  // the same at N = 8192 (6 resident; sweep: 6 -> 5.85 ns, 11 -> 6.84, 20 -> 9.33, none 6.97)
#endif  // 128-bit loads in flight per lane in phase 1

template <int HALF, bool FLEX>
__global__ void __launch_bounds__(32 * kWPC)
peaks_f32_fast_kernel(const float2 *__restrict__ spec, int64_t batch, double df_all, const double *__restrict__ d_fs,
                      int k, unsigned char *__restrict__ recs, int *__restrict__ repair) {
    using P = K3<float, HALF>;
    constexpr int N = 2 * HALF;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int nslot_s[kWPC];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t win = (int64_t)blockIdx.x * kWPC + warp;
    if (win >= batch) return;
    unsigned char *base = smem_raw + warp * P::BYTES;
    float *mags = reinterpret_cast<float *>(base);
    Slot *slots = reinterpret_cast<Slot *>(base + P::MAGW * 4);
    unsigned char *rec_s = base + P::REC_OFF;
    if (lane == 0) nslot_s[warp] = 0;
    if (lane < 16)  // empty record: count/status 0, every peak {idx -1, width 0, mag 0, prominence 0}
        reinterpret_cast<uint64_t *>(rec_s)[lane] = (lane % 3 == 1) ? 0x00000000ffffffffull : 0ull;

    // ---- phase 1: stream the half spectrum, magnitudes -> shared memory, statistics in registers -------------------
    const float4 *src = reinterpret_cast<const float4 *>(spec + win * (int64_t)N);
    float sum = 0.f, sumsq = 0.f;
    constexpr int ROWS = HALF / 64, BATCH = ROWS < kK3Batch ? ROWS : kK3Batch;
    // shared-memory word of bin b = 64*R + 2*lane is  row_off(R) + lane_off  (4-word pad per chunk of C bins)
    const int lane_off = P::addr(2 * lane);
#pragma unroll 1
    for (int r0 = 0; r0 < ROWS; r0 += BATCH) {
        float4 z[BATCH];
#pragma unroll
        for (int u = 0; u < BATCH; ++u) z[u] = ldg_stream(src + (r0 + u) * 32 + lane);
        float *dst = mags + lane_off + P::row_off(r0);
#pragma unroll
        for (int u = 0; u < BATCH; ++u) {
            const float p0 = fmaf(z[u].x, z[u].x, z[u].y * z[u].y), p1 = fmaf(z[u].z, z[u].z, z[u].w * z[u].w);
            const float m0 = sqrt_fast(p0), m1 = sqrt_fast(p1);
            sum += m0 + m1;
            sumsq += p0 + p1;
            *reinterpret_cast<float2 *>(dst + P::row_off(u)) = make_float2(m0, m1);
        }
    }
    double S = (double)sum, Q = (double)sumsq;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        S += __shfl_xor_sync(0xffffffffu, S, o);
        Q += __shfl_xor_sync(0xffffffffu, Q, o);
    }
    const double nn = (double)HALF;
    const double mean = S / nn;
    const double var = (Q - S * S / nn) / (nn - 1.0);
    const double sd = var > 0.0 ? sqrt(var) : 0.0;
    const double thr = mean + 2.0 * sd;
    const float thr_f = __double2float_rd(thr);  // for floats m:  m > thr  <=>  m > thr_f
    const double df = d_fs ? div_rn(d_fs[win], (double)N) : df_all;  // fs / n (host-side division when fs is shared)
    __syncwarp();
    // The tail below issues no global loads, and two thirds of a warp's life is tail: the window a later warp of this
    // slot will stream (11 CTAs of kWPC windows per SM are resident) is requested into L2 now, so phase 1 of that warp
    // waits an L2 instead of a DRAM round trip.  N = 4096 fp32: 2.92 -> 2.49 ns per window, N = 8192: 6.97 -> 5.85; short windows do not gain.
    if (APDA_L2_PREFETCH && HALF >= 2048) {
        const int64_t wn = win + (int64_t)sm_count_reg() * ((HALF >= 4096 ? APDA_K3_PF_CTAS_LONG : APDA_K3_PF_CTAS) * kWPC);
        if (wn < batch) l2_prefetch_span(spec + wn * (int64_t)N, HALF * (int)sizeof(float2), lane, 32);
    }
    k3_tail<float, HALF, FLEX>(mags, slots, P::SLOTS, rec_s, &nslot_s[warp], sd, thr_f, df, k, lane, win, recs, repair);
}


template <int HALF>
int launch_half(apda_ctx *ctx, cudaStream_t st, const float *d_spec, int64_t batch, double fs, const double *d_fs,
                int k, int flexible, void *d_rec) {
    const int smem = kWPC * K3<float, HALF>::BYTES;
    auto kern = flexible ? peaks_f32_fast_kernel<HALF, true> : peaks_f32_fast_kernel<HALF, false>;
    APDA_FUNC_SMEM(ctx, kern, smem);
    const int64_t blocks = (batch + kWPC - 1) / kWPC;
    // repair list: [0] = count, [1..] = windows whose candidate list did not fit on chip
    int *repair = nullptr;
    APDA_TRY(apda_repair_list(ctx, st, batch, &repair));
    kern<<<(unsigned)blocks, 32 * kWPC, smem, st>>>(reinterpret_cast<const float2 *>(d_spec), batch,
                                                    fs / (double)(2 * HALF), d_fs, k,
                                                    reinterpret_cast<unsigned char *>(d_rec), repair);
    ctx->launches++;
    APDA_CUDA(cudaGetLastError());
    // the general kernel re-does the listed windows (grid-stride over the device-side count; exits at once if empty)
    return launch_peaks_general_listed(ctx, st, d_spec, 2 * HALF, batch, fs, d_fs, k, 5, flexible, d_rec, repair);
}

}  // namespace

bool peaks_f32_fast_supports(int64_t n, int k, int rec_cap) {
    return (n == 1024 || n == 2048 || n == 4096 || n == 8192) && rec_cap == 5 && k >= 1 && k <= 5;
}

int launch_peaks_f32_fast(apda_ctx *ctx, cudaStream_t st, const float *d_spec, int64_t n, int64_t batch, double fs,
                          const double *d_fs, int k, int flexible, void *d_rec) {
    switch (n) {
        case 1024: return launch_half<512>(ctx, st, d_spec, batch, fs, d_fs, k, flexible, d_rec);
        case 2048: return launch_half<1024>(ctx, st, d_spec, batch, fs, d_fs, k, flexible, d_rec);
        case 4096: return launch_half<2048>(ctx, st, d_spec, batch, fs, d_fs, k, flexible, d_rec);
        case 8192: return launch_half<4096>(ctx, st, d_spec, batch, fs, d_fs, k, flexible, d_rec);
    }
    apda_set_error("peaks_f32_fast: unsupported n=%lld", (long long)n);
    return APDA_ERR_UNSUPPORTED;
}
