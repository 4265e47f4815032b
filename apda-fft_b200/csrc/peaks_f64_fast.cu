// K3 (fp64 fast form): spectrum in HBM -> records, n = 1024 / 2048 / 4096 / 8192, k <= 5.  One warp per window; the
// picker tail is the type-generic code of peaks_fast.cuh, so decisions are those of the general kernel (peaks.cu) and
// the records are byte-identical to it (tests: apda_ctx_set_generic_only on/off).
//
// This kernel is bound by the fp64 pipe (64 lanes per SM per clock), not by HBM: every magnitude is glibc's hypot
// operation sequence (one correctly rounded square root, one correctly rounded division, ~15 further individually
// rounded operations: abs(complex) of the reference, utils/get_peak_prominence.py:159) and the window statistics are
// accumulated in double-double (statistics.mean / stdev return correctly rounded exact values, :163-164).  So the
// phase-1 loop is written for fp64-pipe efficiency: range checks on the integer pipe (one test on the high words sends
// zero / huge / tiny operands to the full-range routine), eight independent magnitudes in flight per lane, and the
// cheapest error-free accumulation that is still exact to ~2^-94 (all summands are non-negative: no cancellation).
#include "peaks_fast.cuh"
#include "peaks_common.cuh"

namespace {

constexpr int kWPC64 = 2;  // windows per CTA
template <int HALF>
constexpr int kWPWof = HALF >= 4096 ? 4 : 2;  // warps per window during phase 1
constexpr int kBatch = 4;  // rows (independent magnitudes per lane) in flight

// butterfly sum of non-negative (hi, lo) pairs (lo: accumulated rounding errors, not normalised): TwoSum on the high
// words, plain adds on the low ones - exact to ~2^-100 relative because nothing cancels; returns a normalised pair
__device__ __forceinline__ dd warp_sum_pos(double hi, double lo) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ohi = __shfl_xor_sync(0xffffffffu, hi, o), olo = __shfl_xor_sync(0xffffffffu, lo, o);
        const dd s = two_sum(hi, ohi);
        hi = s.hi;
        lo = add_rn(add_rn(lo, olo), s.lo);
    }
    return two_sum(hi, lo);
}

__device__ __forceinline__ double2 ldg_stream2(const double2 *p) {
    double2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}

// kWPW warps share one window during phase 1 (the fp64-pipe-bound part: ~80 % of the instructions) - shared memory per
// window (the fp64 magnitudes) is what limits residency, so this doubles the warps per SM that feed the fp64 pipe.
// The picker tail is a one-warp job: the window's other warps retire after handing over their partial sums.
template <int HALF, bool FLEX>
__global__ void __launch_bounds__(32 * kWPWof<HALF> * kWPC64)
peaks_f64_fast_kernel(const double2 *__restrict__ spec, int64_t batch, double df_all, const double *__restrict__ d_fs,
                      int k, unsigned char *__restrict__ recs, int *__restrict__ repair) {
    using P = K3<double, HALF>;
    constexpr int N = 2 * HALF, kWPW = kWPWof<HALF>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int nslot_s[kWPC64];
    __shared__ dd part[kWPC64][kWPW][2];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wslot = warp / kWPW, sub = warp % kWPW;  // window slot of the CTA, this warp's share of the window
    const int64_t win = (int64_t)blockIdx.x * kWPC64 + wslot;
    if (win >= batch) return;
    unsigned char *base = smem_raw + wslot * P::BYTES;
    double *mags = reinterpret_cast<double *>(base);
    SlotT<double> *slots = reinterpret_cast<SlotT<double> *>(base + P::MAGW * 8);
    unsigned char *rec_s = base + P::REC_OFF;
    if (sub == 0) {
        if (lane == 0) nslot_s[wslot] = 0;
        if (lane < 16)  // empty record: count/status 0, every peak {idx -1, width 0, mag 0, prominence 0}
            reinterpret_cast<uint64_t *>(rec_s)[lane] = (lane % 3 == 1) ? 0x00000000ffffffffull : 0ull;
    }

    // the loop below asks for four rows, then spends ~250 fp64 instructions on them: the whole half spectrum is requested
    // into L2 up front, so every batch after the first waits an L2 instead of a DRAM round trip
    if (APDA_L2_PREFETCH) l2_prefetch_span(spec + win * (int64_t)N, HALF * (int)sizeof(double2), sub * 32 + lane, 32 * kWPW);
    // ---- phase 1: stream the half spectrum (one bin per lane and row), magnitudes -> shared memory, double-double sums ----
    const double2 *src = spec + win * (int64_t)N + lane;
    double sx_hi = 0.0, sx_lo = 0.0, sq_hi = 0.0, sq_lo = 0.0;
    constexpr int ROWS = HALF / 32, BATCH = kBatch, NB = ROWS / BATCH;
    static_assert(NB % kWPW == 0, "rows must split evenly over the window's warps");
#pragma unroll 1
    for (int i = sub; i < NB; i += kWPW) {
        double2 z[BATCH];
#pragma unroll
        for (int u = 0; u < BATCH; ++u) z[u] = ldg_stream2(src + (i * BATCH + u) * 32);
        double m[BATCH];
        unsigned bad = 0;
#pragma unroll
        for (int u = 0; u < BATCH; ++u) {
            bool ok;
            m[u] = magnitude_mid(z[u].x, z[u].y, ok);
            bad |= ok ? 0u : 1u << u;
        }
        if (bad) {  // zero (bin 0), huge or tiny operands: the full-range routine
#pragma unroll
            for (int u = 0; u < BATCH; ++u)
                if (bad >> u & 1u) m[u] = magnitude_full_range(z[u].x, z[u].y);
        }
        const int r0 = i * BATCH;
#pragma unroll
        for (int u = 0; u < BATCH; ++u) {
            mags[P::addr((r0 + u) * 32 + lane)] = m[u];
            const dd s = two_sum(sx_hi, m[u]);
            sx_hi = s.hi;
            sx_lo = add_rn(sx_lo, s.lo);
            const dd p = two_prod(m[u], m[u]);
            const dd q = two_sum(sq_hi, p.hi);
            sq_hi = q.hi;
            sq_lo = add_rn(sq_lo, add_rn(q.lo, p.lo));
        }
    }
    dd a = warp_sum_pos(sx_hi, sx_lo);
    dd b = warp_sum_pos(sq_hi, sq_lo);
    if (kWPW > 1) {
        if (lane == 0) {
            part[wslot][sub][0] = a;
            part[wslot][sub][1] = b;
        }
        // this window's warps only; literal barrier ids so that the CTA reserves kWPC64 + 1 barriers, not all 16
        if (wslot == 0) asm volatile("bar.sync 1, %0;" ::"n"(32 * kWPW) : "memory");
        else asm volatile("bar.sync 2, %0;" ::"n"(32 * kWPW) : "memory");
        static_assert(kWPC64 <= 2, "one named barrier per window slot");
        if (sub != 0) return;
        a = part[wslot][0][0];
        b = part[wslot][0][1];
#pragma unroll
        for (int w = 1; w < kWPW; ++w) {
            a = dd_add(a, part[wslot][w][0]);
            b = dd_add(b, part[wslot][w][1]);
        }
    }
    // mean / sample sigma / threshold: the values block_stats (peaks_common.cuh) computes.  The bin count is a power of
    // two, so sum / n and sum^2 / n are exact scalings of the double-double pairs; only the division by n - 1 is real.
    const double nn = (double)HALF, inv_n = 1.0 / (double)HALF;
    const dd mean_dd = {mul_rn(a.hi, inv_n), mul_rn(a.lo, inv_n)};
    const dd a2 = dd_mul(a, a);
    const dd ss = dd_add(b, dd{-mul_rn(a2.hi, inv_n), -mul_rn(a2.lo, inv_n)});  // sxx - sx^2/n  (cancellation: accurate add)
    const dd var = dd_div_d(ss, nn - 1.0);
    const double mean = add_rn(mean_dd.hi, mean_dd.lo);
    const double sd = dd_sqrt_to_double(var);
    const double thr = add_rn(mean, mul_rn(2.0, sd));
    const double df = d_fs ? div_rn(d_fs[win], (double)N) : df_all;
    __syncwarp();
    k3_tail<double, HALF, FLEX>(mags, slots, P::SLOTS, rec_s, &nslot_s[wslot], sd, thr, df, k, lane, win, recs, repair);
}

template <int HALF>
int launch_half64(apda_ctx *ctx, cudaStream_t st, const double *d_spec, int64_t batch, double fs, const double *d_fs,
                  int k, int flexible, void *d_rec) {
    const int smem = kWPC64 * K3<double, HALF>::BYTES;
    auto kern = flexible ? peaks_f64_fast_kernel<HALF, true> : peaks_f64_fast_kernel<HALF, false>;
    {
        static std::mutex mu;  // carve-out preference: once per (device, kernel), like the dynamic shared memory size
        static std::map<std::pair<int, const void *>, bool> carved;
        std::lock_guard<std::mutex> lock(mu);
        bool &done = carved[std::make_pair(ctx->device, reinterpret_cast<const void *>(kern))];
        if (!done) {
            APDA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
            done = true;
        }
    }
    APDA_FUNC_SMEM(ctx, kern, smem);
    const int64_t blocks = (batch + kWPC64 - 1) / kWPC64;
    int *repair = nullptr;  // repair list, see peaks_f32_fast.cu
    APDA_TRY(apda_repair_list(ctx, st, batch, &repair));
    kern<<<(unsigned)blocks, 32 * kWPWof<HALF> * kWPC64, smem, st>>>(reinterpret_cast<const double2 *>(d_spec), batch,
                                                      fs / (double)(2 * HALF), d_fs, k,
                                                      reinterpret_cast<unsigned char *>(d_rec), repair);
    ctx->launches++;
    APDA_CUDA(cudaGetLastError());
    return launch_peaks_general_listed<double>(ctx, st, d_spec, 2 * HALF, batch, fs, d_fs, k, 5, flexible, d_rec,
                                               repair);
}

}  // namespace

bool peaks_f64_fast_supports(int64_t n, int k, int rec_cap) {
    return (n == 1024 || n == 2048 || n == 4096 || n == 8192) && rec_cap == 5 && k >= 1 && k <= 5;
}

int launch_peaks_f64_fast(apda_ctx *ctx, cudaStream_t st, const double *d_spec, int64_t n, int64_t batch, double fs,
                          const double *d_fs, int k, int flexible, void *d_rec) {
    switch (n) {
        case 1024: return launch_half64<512>(ctx, st, d_spec, batch, fs, d_fs, k, flexible, d_rec);
        case 2048: return launch_half64<1024>(ctx, st, d_spec, batch, fs, d_fs, k, flexible, d_rec);
        case 4096: return launch_half64<2048>(ctx, st, d_spec, batch, fs, d_fs, k, flexible, d_rec);
        case 8192: return launch_half64<4096>(ctx, st, d_spec, batch, fs, d_fs, k, flexible, d_rec);
    }
    apda_set_error("peaks_f64_fast: unsupported n=%lld", (long long)n);
    return APDA_ERR_UNSUPPORTED;
}
