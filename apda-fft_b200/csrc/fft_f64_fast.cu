// K1 (fp64 bit-faithful, register blocked) for N = 1024 / 2048 / 4096 / 8192.
//
// Same dataflow graph, same recurrence twiddle table and the same individually rounded operations as the general
// kernel (fft_smem.cu) - so the spectrum stays bit-identical to the reference (metrics/fft_iterativa.py:38-70) - but the
// log2(N) radix-2 stages are executed 4 (or 3) at a time on 16 register-resident values per thread: only
// ceil(log2 N / 4) - 1 shared-memory exchanges instead of log2(N) block-wide stages.
//
//   load      n_samples reals -> shared memory (coalesced), each thread gathers its 16 bit-reversed inputs
//   median    exact counting selection on the register-resident values (statistics.median semantics)
//   pass 0    stages 1..4 in registers (twiddles: 15 table entries shared by all threads), results -> shared memory
//   pass p    stages s0+1..s0+q: 2^q values per work item with stride 2^s0, twiddles T_s[(r mod 2^(t-1))*2^s0 + lo]
//             read coalesced through L1; in place in shared memory (index padded by idx>>4: conflict free)
//   last pass results go straight from registers to HBM with coalesced 128-bit stores; bin 0 := 0.
#include "common.cuh"

#include <mutex>
#include <set>

namespace {

__device__ __forceinline__ void bfly(double2 &u, double2 &v, const double2 w) {
    const double vr = sub_rn(mul_rn(v.x, w.x), mul_rn(v.y, w.y));
    const double vi = add_rn(mul_rn(v.x, w.y), mul_rn(v.y, w.x));
    const double2 a = make_double2(add_rn(u.x, vr), add_rn(u.y, vi));
    const double2 b = make_double2(sub_rn(u.x, vr), sub_rn(u.y, vi));
    u = a;
    v = b;
}

template <int LOGN>
struct Plan64 {
    static constexpr int N = 1 << LOGN;
    static constexpr int T = N / 16;  // threads per window
    static constexpr int NP = (LOGN + 3) / 4;
    // stages per pass: first pass always 4, the rest split the remainder as evenly as possible (4s first)
    __host__ __device__ static constexpr int q(int p) {
        return p == 0 ? 4 : ((LOGN - 4) / (NP - 1) + ((p - 1) < (LOGN - 4) % (NP - 1) ? 1 : 0));
    }
    __host__ __device__ static constexpr int s0(int p) { return p == 0 ? 0 : s0(p - 1) + q(p - 1); }
    static constexpr int SM_ELEMS = N + N / 16;  // complex elements incl. padding
};

__device__ __forceinline__ int pad16(int idx) { return idx + (idx >> 4); }

// smallest / largest of the warp's values inside [lo, hi); out of line on purpose (see bracket_minmax in fft_f32_fast.cuh)
__device__ __noinline__ double2 bracket_minmax_f64(double a0, double a1, double a2, double a3, double a4, double a5, double a6,
                                                   double a7, double a8, double a9, double a10, double a11, double a12,
                                                   double a13, double a14, double a15, double lo, double hi) {
    const double a[16] = {a0, a1, a2, a3, a4, a5, a6, a7, a8, a9, a10, a11, a12, a13, a14, a15};
    double vmin = CUDART_INF, vmax = -CUDART_INF;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        if (a[i] >= lo && a[i] < hi) {
            vmin = fmin(vmin, a[i]);
            vmax = fmax(vmax, a[i]);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        vmin = fmin(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
        vmax = fmax(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
    }
    return make_double2(vmin, vmax);
}

// exact statistics.median of the window; every thread holds 16 of its values (invalid slots: +inf)
template <int T>
__device__ double select_median_f64(const double (&val)[16], int n_valid, double *shd /* 72 doubles */, int tid) {
    constexpr int NW = T / 32 > 0 ? T / 32 : 1;
    const int lane = tid & 31, warp = tid >> 5;
    const int r_lo = (n_valid - 1) >> 1, r_hi = n_valid >> 1;
    unsigned *shu = reinterpret_cast<unsigned *>(shd + 48);
    // mean / standard deviation only steer the first two pivots: every thread contributes 4 of its 16 values, every warp
    // reduces its share and all threads combine the partial sums after one barrier (no warp idles during the estimate).
    // shd: [2, 34) gather list (first: partial sums / even-n temporaries), 40 result, 41 gather counter, [48, 64) the two
    // banks of per-warp round counts, [64, 72) the warps' valid-sample counts of the prologue
    unsigned *gather_cnt = reinterpret_cast<unsigned *>(shd + 41);
    unsigned *cnt_part = reinterpret_cast<unsigned *>(shd + 64);
    {
        double s1 = 0.0, s2 = 0.0;
        int cv = 0;
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
            const bool ok = val[i] < CUDART_INF;
            const double x = ok ? val[i] : 0.0;
            s1 += x;
            s2 += x * x;
            cv += ok;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        }
        cv = __reduce_add_sync(0xffffffffu, cv);
        if (lane == 0) {
            shd[2 + warp] = s1;
            shd[2 + 16 + warp] = s2;
            cnt_part[warp] = (unsigned)cv;
        }
        if (tid == 0) *gather_cnt = 0;
    }
    __syncthreads();
    double mean, sd;
    {
        double s1 = 0.0, s2 = 0.0;
        int cv = 0;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            s1 += shd[2 + w];
            s2 += shd[2 + 16 + w];
            cv += (int)cnt_part[w];
        }
        // steering values only (the counts decide): single-precision reciprocals / roots instead of fp64 divisions and
        // a square root on the pipe the transform needs; data outside the float range takes the exact forms
        const double inv_nv = (double)rcp_approx((float)max(cv, 1));
        mean = s1 * inv_nv;
        const double var = fmax(s2 * inv_nv - mean * mean, 0.0);
        const float var_f = (float)var;
        sd = (var_f > 1e-30f && var_f < 1e30f) ? (double)sqrt_approx(var_f) : sqrt(var);
    }
    // 1 / (values per unit near the centre)
    const double inv_density = fmax(2.5 * sd, 1e-300) * (double)rcp_approx((float)n_valid);
    double lo = -CUDART_INF, hi = CUDART_INF;  // bracket [lo, hi): c_lo = #(v < lo) <= r_lo, c_hi = #(v < hi) > r_hi
    int c_lo = 0, c_hi = n_valid;
    double pivot = mean;
    for (int round = 0;; ++round) {
        if (round > 0) {
            if (c_hi - c_lo <= 32) break;
            if (round >= 4) {
                // quantised samples (see select_median in fft_f32_fast.cuh): cut the bracket to the smallest / largest value
                // inside it; equal -> that value is the median, else every further pivot removes at least one level
                const double2 mm = bracket_minmax_f64(val[0], val[1], val[2], val[3], val[4], val[5], val[6], val[7], val[8],
                                                      val[9], val[10], val[11], val[12], val[13], val[14], val[15], lo, hi);
                if (lane == 0) {
                    shd[2 + warp] = mm.x;
                    shd[2 + 16 + warp] = mm.y;
                }
                __syncthreads();
                double vmin = shd[2], vmax = shd[2 + 16];
#pragma unroll
                for (int w = 1; w < NW; ++w) {
                    vmin = fmin(vmin, shd[2 + w]);
                    vmax = fmax(vmax, shd[2 + 16 + w]);
                }
                if (vmin == vmax) {
                    __syncthreads();
                    return vmin;
                }
                lo = vmin;
                hi = key_value(ordered_key(vmax) + 1ull, 0.0);
                if (!(hi > vmax)) hi = key_value(ordered_key(vmax) + 2ull, 0.0);
            }
            double lo_next = -1.7976931348623157e308;
            if (lo > -CUDART_INF) {
                lo_next = key_value(ordered_key(lo) + 1ull, 0.0);
                if (!(lo_next > lo)) lo_next = key_value(ordered_key(lo) + 2ull, 0.0);
            }
            if (!(lo_next < hi)) break;
            const double want = (double)r_lo + 0.5;
            if (lo == -CUDART_INF) {
                pivot = hi - 1.5 * fmax((double)c_hi - want, 1.0) * inv_density * (double)(1 << min(round - 1, 30));
            } else if (hi == CUDART_INF) {
                pivot = lo + 1.5 * fmax(want - (double)c_lo, 1.0) * inv_density * (double)(1 << min(round - 1, 30));
            } else if (round % 4 == 0) {  // key-space bisection bounds the worst case
                const uint64_t a = ordered_key(lo), b = ordered_key(hi);
                pivot = key_value(a + ((b - a) >> 1), 0.0);
            } else {
                pivot = lo + (hi - lo) * (double)__fdividef((float)r_lo + 0.5f - (float)c_lo, (float)(c_hi - c_lo));
            }
            if (!(pivot > lo)) pivot = lo_next;
            if (!(pivot < hi)) pivot = lo_next;
        }
        int cnt = 0;
#pragma unroll
        for (int i = 0; i < 16; ++i) cnt += (val[i] < pivot);
        cnt = __reduce_add_sync(0xffffffffu, cnt);
        unsigned *cslot = shu + 16 * (round & 1);  // alternating banks: one barrier per round
        if (lane == 0) cslot[warp] = (unsigned)cnt;
        __syncthreads();
        int tot = 0;
#pragma unroll
        for (int w = 0; w < NW; ++w) tot += (int)cslot[w];
        if (tot <= r_lo) {
            lo = pivot;
            c_lo = tot;
        } else if (tot > r_hi) {
            hi = pivot;
            c_hi = tot;
        } else {  // even n, the pivot separates the two middle values
            double below = -CUDART_INF, above = CUDART_INF;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                if (val[i] < pivot) below = fmax(below, val[i]);
                else above = fmin(above, val[i]);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                below = fmax(below, __shfl_xor_sync(0xffffffffu, below, o));
                above = fmin(above, __shfl_xor_sync(0xffffffffu, above, o));
            }
            if (lane == 0) {
                shd[2 + warp] = below;
                shd[2 + 16 + warp] = above;
            }
            __syncthreads();
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                below = fmax(below, shd[2 + w]);
                above = fmin(above, shd[2 + 16 + w]);
            }
            __syncthreads();
            return div_rn(add_rn(below, above), 2.0);
        }
    }
    if (c_hi - c_lo > 32) {  // the loop ends like this only with ONE distinct value in the bracket: both middle order statistics
        __syncthreads();
        return lo;
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {  // the counter was zeroed in the prologue; the list area was last read before round 0's barrier
        if (val[i] >= lo && val[i] < hi) {
            const unsigned pos = atomicAdd(gather_cnt, 1u);
            if (pos < 32u) shd[2 + pos] = val[i];
        }
    }
    __syncthreads();
    const int cnt = (int)*gather_cnt;
    if (warp == 0) {
        double med;
        if (cnt > 32) {
            med = lo;  // a single distinct value fills the bracket
        } else {
            const double mine = lane < cnt ? shd[2 + lane] : CUDART_INF;
            int rank = 0;
            for (int j = 0; j < cnt; ++j) {
                const double other = shd[2 + j];
                rank += (other < mine) || (other == mine && j < lane);
            }
            const unsigned m_lo = __ballot_sync(0xffffffffu, lane < cnt && rank == r_lo - c_lo);
            const unsigned m_hi = __ballot_sync(0xffffffffu, lane < cnt && rank == r_hi - c_lo);
            const double a = __shfl_sync(0xffffffffu, mine, __ffs(m_lo) - 1);
            const double b = __shfl_sync(0xffffffffu, mine, __ffs(m_hi) - 1);
            med = div_rn(add_rn(a, b), 2.0);
        }
        if (lane == 0) shd[40] = med;
    }
    __syncthreads();
    return shd[40];  // nothing writes this word again before the window is done
}

// Twiddles of stages 1..4 (entries [0, 15) of the recurrence table: stage with half-span h at [h-1, 2h-1)).  They depend
// on the stage only, not on N, and every thread of pass 0 needs the same 15: in the constant bank they are operands of
// the multiplications themselves - no loads, no scoreboard waits, no registers (ncu: pass 0 spent half of its stall
// samples on the L1 round trips of these 15 loads).
__constant__ double2 c_tw_head[15];

#ifndef APDA_K1F64_PF
#define APDA_K1F64_PF 1
#endif

__device__ __forceinline__ void stages_head(double2 *v) {
#pragma unroll
    for (int t = 1; t <= 4; ++t) {
        const int h = 1 << (t - 1);
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            if ((r & h) == 0) bfly(v[r], v[r + h], c_tw_head[(h - 1) + (r & (h - 1))]);
        }
    }
}

// q radix-2 stages on 2^q register values of one work item (stride 2^s0, low index bits `lo`)
template <int Q>
__device__ __forceinline__ void stages(double2 *v, const double2 *__restrict__ tw, int s0, int lo) {
#pragma unroll
    for (int t = 1; t <= Q; ++t) {
        const int h = 1 << (t - 1);
        const double2 *tab = tw + (((1 << (s0 + t - 1)) - 1) + lo);
        double2 w[8];
#pragma unroll
        for (int j = 0; j < h; ++j) w[j] = __ldg(tab + (j << s0));
#pragma unroll
        for (int r = 0; r < (1 << Q); ++r) {
            if ((r & h) == 0) bfly(v[r], v[r + h], w[r & (h - 1)]);
        }
    }
}

template <int LOGN, int P, int Q>
__device__ __forceinline__ void middle_or_last_pass(double2 *s, const double2 *__restrict__ tw, double2 *out, int t) {
    using PL = Plan64<LOGN>;
    constexpr int S0 = PL::s0(P), G = 16 >> Q, T = PL::T;
    constexpr bool LAST = (P == PL::NP - 1);
#pragma unroll
    for (int g = 0; g < G; ++g) {
        const int w = t + T * g;                 // work item: lo = low S0 bits, hi = the rest
        const int lo = w & ((1 << S0) - 1), hi = w >> S0;
        const int base = (hi << (S0 + Q)) + lo;
        // pad16(base + (r << S0)) = pad16(base) + r * (2^S0 + 2^(S0-4)) because S0 >= 4: one padded base per work item and
        // compile-time offsets, instead of a shift and two adds per access
        static_assert(S0 >= 4, "the padded offsets are linear in r only from the second pass on");
        constexpr int RSTEP = (1 << S0) + (1 << (S0 - 4));
        double2 *sp = s + pad16(base);
        double2 v[1 << Q];
#pragma unroll
        for (int r = 0; r < (1 << Q); ++r) v[r] = sp[r * RSTEP];
        stages<Q>(v, tw, S0, lo);
        if (LAST) {
#pragma unroll
            for (int r = 0; r < (1 << Q); ++r) {
                double2 val = v[r];
                if (base + (r << S0) == 0) val = make_double2(0.0, 0.0);  // reference: res[0] = 0
                out[base + (r << S0)] = val;
            }
        } else {
#pragma unroll
            for (int r = 0; r < (1 << Q); ++r) sp[r * RSTEP] = v[r];
        }
    }
}

template <int LOGN>
__global__ void __launch_bounds__(Plan64<LOGN>::T, (LOGN <= 12 ? 768 / Plan64<LOGN>::T : 1))
fft_f64_fast_kernel(const double *__restrict__ samples, int n_samples, int64_t ld, const double2 *__restrict__ tw,
                    double2 *__restrict__ spec, int center, const int *__restrict__ nv) {
    using PL = Plan64<LOGN>;
    constexpr int N = PL::N, T = PL::T;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double shd[72];
    double2 *s = reinterpret_cast<double2 *>(smem_raw);
    double *raw = reinterpret_cast<double *>(smem_raw);  // staging of the real samples (padded by i>>3), aliases s
    const int t = threadIdx.x;
    const int64_t win = blockIdx.x;
    if (nv && nv[win] != n_samples) return;  // ragged batch: windows of another length go to the general kernel
    const double *x = samples + win * ld;

    {   // streaming loads (L1 is kept for the twiddle table); all 16 loads of a thread are issued before the first store,
        // so one HBM latency is exposed instead of sixteen
        double ld_v[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            const int i = t + T * u;
            ld_v[u] = i < n_samples ? __ldcs(x + i) : CUDART_INF;
        }
        double *rp = raw + (t + (t >> 3));  // (t + T*u) + ((t + T*u) >> 3) = t + (t >> 3) + u * (T + T/8): T is a multiple of 8
#pragma unroll
        for (int u = 0; u < 16; ++u) rp[u * (T + T / 8)] = ld_v[u];
    }
    if (APDA_L2_PREFETCH && APDA_K1F64_PF) {  // samples of the window the next CTA of this slot will load (see l2_prefetch_span)
        const int64_t wn = win + (int64_t)sm_count_reg() * (LOGN <= 12 ? 768 / T : 1);
        if (wn < (int64_t)gridDim.x) l2_prefetch_span(samples + wn * ld, n_samples * (int)sizeof(double), t, T);
    }
    __syncthreads();
    // work item hi = t of pass 0 owns outputs idx = 16*t + r, i.e. inputs bitrev(idx) = bitrev4(r) * N/16 + bitrev(t)
    const int tb = (int)(__brev((unsigned)t) >> (32 - (LOGN - 4)));
    double val[16];
    {
        const double *gp = raw + (tb + (tb >> 3));  // src + (src >> 3) with src = brev4(r) * N/16 + tb: linear in brev4(r)
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            constexpr int RS = (1 << (LOGN - 4)) + (1 << (LOGN - 7));
            val[r] = gp[(int)(__brev((unsigned)r) >> 28) * RS];
        }
    }
    __syncthreads();  // raw is dead from here on (s aliases it)
    double med = 0.0;
    if (center == APDA_CENTER_MEDIAN) med = select_median_f64<T>(val, n_samples, shd, t);
    double2 v[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) v[r] = make_double2(val[r] < CUDART_INF ? sub_rn(val[r], med) : 0.0, 0.0);
    stages_head(v);
#pragma unroll
    for (int r = 0; r < 16; ++r) s[17 * t + r] = v[r];  // pad16(16 t + r) = 16 t + r + t
    __syncthreads();

    double2 *out = spec + win * (int64_t)N;
    middle_or_last_pass<LOGN, 1, PL::q(1)>(s, tw, out, t);
    if (PL::NP > 2) {
        __syncthreads();
        middle_or_last_pass<LOGN, 2 < PL::NP ? 2 : 1, PL::q(2 < PL::NP ? 2 : 1)>(s, tw, out, t);
    }
    if (PL::NP > 3) {
        __syncthreads();
        middle_or_last_pass<LOGN, 3 < PL::NP ? 3 : 1, PL::q(3 < PL::NP ? 3 : 1)>(s, tw, out, t);
    }
}

template <int LOGN>
int launch_n(apda_ctx *ctx, cudaStream_t st, const double *d_samples, int64_t n_samples, int64_t ld, int64_t batch,
             int flags, const double2 *tw, double *d_spec, const int *d_nv) {
    using PL = Plan64<LOGN>;
    const size_t smem = (size_t)PL::SM_ELEMS * sizeof(double2);
    APDA_FUNC_SMEM(ctx, fft_f64_fast_kernel<LOGN>, smem);
    fft_f64_fast_kernel<LOGN><<<(unsigned)batch, PL::T, smem, st>>>(d_samples, (int)n_samples, ld, tw,
                                                                   reinterpret_cast<double2 *>(d_spec), flags, d_nv);
    ctx->launches++;
    APDA_CUDA(cudaGetLastError());
    return APDA_OK;
}

}  // namespace

bool fft_f64_fast_supports(int64_t N) { return N == 1024 || N == 2048 || N == 4096 || N == 8192; }

int launch_fft_f64_fast(apda_ctx *ctx, cudaStream_t st, const double *d_samples, int64_t n_samples, int64_t ld,
                        int64_t batch, int64_t N, int flags, double *d_spec, const int *d_nv) {
    TwiddleTables tw;
    APDA_TRY(apda_get_twiddles(ctx, N, &tw));
    {   // the head twiddles are the same for every N: one upload per device (the symbol is per device, not per context)
        static std::mutex mu;
        static std::set<int> uploaded;
        std::lock_guard<std::mutex> lock(mu);
        if (!uploaded.count(ctx->device)) {
            APDA_CUDA(cudaMemcpyToSymbol(c_tw_head, tw.d64, sizeof(double2) * 15, 0, cudaMemcpyDeviceToDevice));
            uploaded.insert(ctx->device);
        }
    }
    switch (N) {
        case 1024: return launch_n<10>(ctx, st, d_samples, n_samples, ld, batch, flags, tw.d64, d_spec, d_nv);
        case 2048: return launch_n<11>(ctx, st, d_samples, n_samples, ld, batch, flags, tw.d64, d_spec, d_nv);
        case 4096: return launch_n<12>(ctx, st, d_samples, n_samples, ld, batch, flags, tw.d64, d_spec, d_nv);
        case 8192: return launch_n<13>(ctx, st, d_samples, n_samples, ld, batch, flags, tw.d64, d_spec, d_nv);
    }
    apda_set_error("fft_f64_fast: unsupported N=%lld", (long long)N);
    return APDA_ERR_UNSUPPORTED;
}
