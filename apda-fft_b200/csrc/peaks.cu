// K3 (general form): one CTA per window.  magnitude of the half spectrum -> mean / sample sigma -> threshold ->
// strict local maxima -> flexible (prominence) or rigid (resolution) picker -> fixed-size peak record.
// The spectrum is read once from HBM; magnitudes live in shared memory (or in a global workspace for N > 2^14/2^15).
//
// Reference behaviour reproduced (paths relative to the reference checkout):
//   utils/get_peak_prominence.py:149-226 get_top_peaks_prominence (+ :32-54 calculate_prominence,
//       :89-112 calculate_half_power_width_prominenceBased)
//   utils/get_peak_resolution.py:80-128 get_top_peaks_resolution (+ :30-44 width_half_magnitude, :48-62 resolution)
// Every comparison the reference makes in fp64 is made here with individually rounded fp64 operations in the same
// order; round(x, 4) is emulated exactly (round_dec4).  fp32 instantiations keep magnitudes/prominences in fp32 and
// evaluate the statistics and the damping / exclusion / zeroing arithmetic in fp64.
#include <math_constants.h>

#include <algorithm>

#include "common.cuh"
#include "peaks_common.cuh"

namespace {

template <typename T>
__device__ void load_magnitudes(const typename vec2<T>::type *spec, T *mags, int half) {
    // four independent loads in flight per thread before the (long) magnitude arithmetic consumes them
    const int step = blockDim.x;
    int i = threadIdx.x;
    for (; i + 3 * step < half; i += 4 * step) {
        typename vec2<T>::type v0 = spec[i], v1 = spec[i + step], v2 = spec[i + 2 * step], v3 = spec[i + 3 * step];
        mags[i] = magnitude(v0.x, v0.y);
        mags[i + step] = magnitude(v1.x, v1.y);
        mags[i + 2 * step] = magnitude(v2.x, v2.y);
        mags[i + 3 * step] = magnitude(v3.x, v3.y);
    }
    for (; i < half; i += step) {
        typename vec2<T>::type v = spec[i];
        mags[i] = magnitude(v.x, v.y);
    }
    __syncthreads();
}

struct Layout {  // per-window scratch: magnitudes, candidate indices, surviving candidates
    size_t mags_off, cand_off, found_off, acc_off, bytes;
    int cap;
};
template <typename T>
__host__ __device__ inline Layout make_layout(int half) {
    Layout l;
    l.cap = half / 4 + 8;  // bins above mean+2*sigma are < 20 % of all bins (Cantelli), local maxima at most half of those
    l.mags_off = 0;
    size_t o = ((size_t)half * sizeof(T) + 15) & ~(size_t)15;
    l.found_off = o;
    o += (size_t)l.cap * sizeof(Found);
    l.cand_off = o;
    o += (size_t)l.cap * sizeof(int);
    l.acc_off = o;  // accepted peaks (slot / bin index): any k up to cap, not a fixed-size array
    o += (size_t)l.cap * sizeof(int);
    l.bytes = (o + 15) & ~(size_t)15;
    return l;
}

// ---- flexible-structure picker ------------------------------------------------------------------------------------
template <typename T, bool SMEM>
__device__ void peaks_prominence_window(const int64_t win, const typename vec2<T>::type *__restrict__ spec, int64_t n,
                                        int half, double fs_all, const double *__restrict__ d_fs, int k, int rec_cap,
                                        unsigned char *__restrict__ recs, unsigned char *__restrict__ ws) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ dd red[64];
    __shared__ Stats stats_s;
    __shared__ int ncand_s, nfound_s;
    const Layout lay = make_layout<T>(half);
    unsigned char *base = SMEM ? smem_raw : ws + (size_t)blockIdx.x * lay.bytes;
    T *mags = reinterpret_cast<T *>(base + lay.mags_off);
    Found *found = reinterpret_cast<Found *>(base + lay.found_off);
    int *cand = reinterpret_cast<int *>(base + lay.cand_off);
    int *acc_slot = reinterpret_cast<int *>(base + lay.acc_off);
    const int tid = threadIdx.x;
    unsigned char *rec = recs + win * APDA_REC_BYTES(rec_cap);
    const double fs = d_fs ? d_fs[win] : fs_all;
    const double df = div_rn(fs, (double)n);

    __shared__ int tie_s;
    if (tid == 0) ncand_s = nfound_s = tie_s = 0;
    load_magnitudes<T>(spec + win * n, mags, half);
    const Stats st = block_stats<T>(mags, half, red, &stats_s);

    for (int j = 1 + tid; j < half - 1; j += blockDim.x) {
        T m = mags[j];
        if (m > mags[j - 1] && m > mags[j + 1] && (double)m > st.thr) {
            int pos = atomicAdd(&ncand_s, 1);
            if (pos < lay.cap) cand[pos] = j;
        }
        if (sizeof(T) == 4 && fp32_tie_top(mags, half, j, m, st.thr)) tie_s = 1;
    }
    __syncthreads();
    const int ncand = min(ncand_s, lay.cap);
    int status = (ncand_s > lay.cap ? APDA_STATUS_TRUNCATED : 0) | (tie_s ? APDA_STATUS_FP32_TIE : 0);

    const int warp = tid >> 5, nwarp = blockDim.x >> 5, lane = tid & 31;
    const double half_sd = mul_rn(0.5, st.sd);
    for (int c = warp; c < ncand; c += nwarp) {
        const int j = cand[c];
        const T prom = warp_prominence<T>(mags, half, j);
        if (!((double)prom > half_sd)) continue;
        const int bins = half_power_bins<T>(mags, half, prom, j);
        const double width_hz = mul_rn((double)bins, df);
        if (!(width_hz > 0.0)) continue;
        const double fn = mul_rn((double)j, df);
        const double q = div_rn(fn, width_hz);
        const double damping = div_rn(1.0, mul_rn(2.0, q));
        if (0.001 <= damping && damping <= 0.07 && lane == 0) {
            int pos = atomicAdd(&nfound_s, 1);
            Found f;
            f.rmag = round_dec4((double)mags[j]);
            f.prom = (double)prom;
            f.idx = j;
            f.width = bins;
            found[pos] = f;  // pos < cap because nfound <= ncand
        }
    }
    __syncthreads();

    // sorted(candidates, key=rounded mag, reverse=True) is stable, so ties keep ascending idx; the greedy "hump"
    // exclusion then walks that order.  Warp 0 extracts the order one element at a time.
    if (warp == 0) {
        const int nfound = nfound_s;
        int na = 0;
        double prev_mag = CUDART_INF;
        int prev_idx = -1;
        while (na < k) {
            double best = -1.0;
            int best_idx = 0x7fffffff, best_e = -1;
            for (int e = lane; e < nfound; e += 32) {
                double r = found[e].rmag;
                int ix = found[e].idx;
                bool after_prev = r < prev_mag || (r == prev_mag && ix > prev_idx);
                if (after_prev && (r > best || (r == best && ix < best_idx))) {
                    best = r;
                    best_idx = ix;
                    best_e = e;
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                double r = __shfl_xor_sync(0xffffffffu, best, o);
                int ix = __shfl_xor_sync(0xffffffffu, best_idx, o);
                int e = __shfl_xor_sync(0xffffffffu, best_e, o);
                if (e >= 0 && (best_e < 0 || r > best || (r == best && ix < best_idx))) {
                    best = r;
                    best_idx = ix;
                    best_e = e;
                }
            }
            if (best_e < 0) break;
            prev_mag = best;
            prev_idx = best_idx;
            const double cf = round_dec4(mul_rn((double)best_idx, df));
            bool hump = false;
            for (int a = 0; a < na && !hump; ++a) {
                const double af = round_dec4(mul_rn((double)found[acc_slot[a]].idx, df));
                const double rel = div_rn(fabs(sub_rn(cf, af)), af);
                if (rel < 0.05) {
                    const double ratio = div_rn(found[best_e].prom, best);
                    if (ratio < 0.10) hump = true;
                }
            }
            if (!hump) {
                if (lane == 0) acc_slot[na] = best_e;
                ++na;
            }
            __syncwarp();
        }
        if (lane == 0) {
            write_rec_header(rec, na, status);
            for (int a = 0; a < rec_cap; ++a) {
                if (a < na) {
                    const Found f = found[acc_slot[a]];
                    write_rec_peak(rec, a, f.idx, f.width, (double)mags[f.idx], f.prom);
                } else {
                    write_rec_peak(rec, a, -1, 0, 0.0, 0.0);
                }
            }
        }
    }
}

// ---- rigid-structure picker -----------------------------------------------------------------------------------------
template <typename T, bool SMEM>
__device__ void peaks_resolution_window(const int64_t win, const typename vec2<T>::type *__restrict__ spec, int64_t n,
                                        int half, double fs_all, const double *__restrict__ d_fs, int k, int rec_cap,
                                        unsigned char *__restrict__ recs, unsigned char *__restrict__ ws) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ dd red[64];
    __shared__ Stats stats_s;
    __shared__ double best_m[32];
    __shared__ int best_j[32];
    __shared__ int ctl[4];  // chosen idx, zero start, zero end, accepted count
    const Layout lay = make_layout<T>(half);
    unsigned char *base = SMEM ? smem_raw : ws + (size_t)blockIdx.x * lay.bytes;
    T *mags = reinterpret_cast<T *>(base + lay.mags_off);
    int *acc_idx = reinterpret_cast<int *>(base + lay.acc_off);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
    unsigned char *rec = recs + win * APDA_REC_BYTES(rec_cap);
    const double fs = d_fs ? d_fs[win] : fs_all;
    const double df = div_rn(fs, (double)n);
    const double distance = sub_rn(mul_rn(2.0, df), mul_rn(1.0, df));  // frequencies[2] - frequencies[1]

    load_magnitudes<T>(spec + win * n, mags, half);
    const Stats st = block_stats<T>(mags, half, red, &stats_s);
    __shared__ int tie_s;
    if (tid == 0) ctl[3] = tie_s = 0;
    __syncthreads();
    if (sizeof(T) == 4) {  // on the original magnitudes (the zeroing below makes equal bins of its own)
        for (int j = 1 + tid; j < half - 1; j += blockDim.x)
            if (fp32_tie_top(mags, half, j, mags[j], st.thr)) tie_s = 1;
        __syncthreads();
    }

    while (true) {
        // arg-max over strict local maxima above the threshold; ties resolve to the lowest index (first wins)
        double bm = -1.0;
        int bj = -1;
        for (int j = 1 + tid; j < half - 1; j += blockDim.x) {
            T m = mags[j];
            if (m > mags[j - 1] && m > mags[j + 1] && (double)m > bm && (double)m > st.thr) {
                bm = (double)m;
                bj = j;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            double m2 = __shfl_xor_sync(0xffffffffu, bm, o);
            int j2 = __shfl_xor_sync(0xffffffffu, bj, o);
            if (j2 >= 0 && (bj < 0 || m2 > bm || (m2 == bm && j2 < bj))) {
                bm = m2;
                bj = j2;
            }
        }
        if (lane == 0) {
            best_m[warp] = bm;
            best_j[warp] = bj;
        }
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < nwarp; ++w) {
                if (best_j[w] >= 0 && (bj < 0 || best_m[w] > bm || (best_m[w] == bm && best_j[w] < bj))) {
                    bm = best_m[w];
                    bj = best_j[w];
                }
            }
            ctl[0] = bj;
            if (bj >= 0) {
                int na = ctl[3];
                const double f = mul_rn((double)bj, df);
                const int w2 = half_height_bins<T>(mags, half, bj);
                bool separated = true;
                for (int a = 0; a < na && separated; ++a) {
                    const int w1 = half_height_bins<T>(mags, half, acc_idx[a]);
                    double rs = 0.0;
                    if (w1 + w2 != 0) {
                        int dist = bj - acc_idx[a];
                        if (dist < 0) dist = -dist;
                        rs = div_rn(mul_rn(1.18, (double)dist), (double)(w1 + w2));
                    }
                    if (!(rs >= 1.5)) separated = false;
                }
                if (separated) {
                    acc_idx[na] = bj;
                    write_rec_peak(rec, na, bj, w2, bm, 0.0);
                    ctl[3] = na + 1;
                }
                double reach_d = rint(div_rn(mul_rn(f, 0.02), distance));  // Python round(): half to even
                if (!(reach_d >= 0.0)) reach_d = 0.0;
                if (reach_d > (double)half) reach_d = (double)half;
                const int reach = (int)reach_d;
                ctl[1] = max(0, bj - reach);
                ctl[2] = min(half, bj + reach + 1);
            }
        }
        __syncthreads();
        if (ctl[0] < 0) break;
        for (int j = ctl[1] + tid; j < ctl[2]; j += blockDim.x) mags[j] = T(0);
        const bool full = ctl[3] >= k;
        __syncthreads();
        if (full) break;
    }
    if (tid == 0) {
        const int na = ctl[3];
        write_rec_header(rec, na, tie_s ? APDA_STATUS_FP32_TIE : 0);
        for (int a = na; a < rec_cap; ++a) write_rec_peak(rec, a, -1, 0, 0.0, 0.0);
    }
}

// One CTA per window; with a repair list (list[0] = count, list[1..] = window ids, written by the fp32 fast kernel) the
// CTAs stride over the listed windows instead.
template <typename T, bool SMEM, bool FLEX>
__global__ void __launch_bounds__(SMEM ? 256 : 1024) peaks_kernel(const typename vec2<T>::type *__restrict__ spec, int64_t n, int half, int64_t batch,
                             double fs_all, const double *__restrict__ d_fs, int k, int rec_cap,
                             unsigned char *__restrict__ recs, unsigned char *__restrict__ ws,
                             const int *__restrict__ list) {
    for (int64_t it = blockIdx.x;; it += gridDim.x) {
        int64_t win = it;
        if (list) {
            if (it >= list[0]) break;
            win = list[1 + it];
        } else if (it >= batch) {
            break;
        }
        if (FLEX) peaks_prominence_window<T, SMEM>(win, spec, n, half, fs_all, d_fs, k, rec_cap, recs, ws);
        else peaks_resolution_window<T, SMEM>(win, spec, n, half, fs_all, d_fs, k, rec_cap, recs, ws);
        __syncthreads();
    }
}

// ---- module-public helper functions of the reference, on a caller-supplied magnitude list ---------------------------
// out[0] = calculate_prominence(mags, idx); out[1] = half-power bin count for prominence prom_in; out[2] = width_half_magnitude
__global__ void mag_helpers_kernel(const double *__restrict__ mags, int n, int idx, double prom_in, double *out) {
    double prom = warp_prominence<double>(mags, n, idx);
    if (threadIdx.x == 0) {
        out[0] = prom;
        out[1] = (double)half_power_bins<double>(mags, n, prom_in, idx);
        out[2] = (double)half_height_bins<double>(mags, n, idx);
    }
}

}  // namespace

template <typename T>
size_t peaks_mag_workspace_bytes(apda_ctx *ctx, int64_t n, int64_t batch) {
    if (peaks_large_supports(n) && !ctx->generic_only) return peaks_large_workspace_bytes<T>(n, batch);
    Layout lay = make_layout<T>((int)(n / 2));
    if (lay.bytes + 4096 <= (size_t)ctx->smem_optin) return 0;
    return lay.bytes * (size_t)batch;
}
template size_t peaks_mag_workspace_bytes<double>(apda_ctx *, int64_t, int64_t);
template size_t peaks_mag_workspace_bytes<float>(apda_ctx *, int64_t, int64_t);

template <typename T>
static int launch_peaks_general(apda_ctx *ctx, cudaStream_t st, const T *d_spec, int64_t n, int64_t batch, double fs,
                                const double *d_fs, int k, int rec_cap, int flexible, void *d_rec, void *d_mag_ws,
                                const int *list) {
    using V2 = typename vec2<T>::type;
    const int half = (int)(n / 2);
    const Layout lay = make_layout<T>(half);
    const bool in_smem = lay.bytes + 4096 <= (size_t)ctx->smem_optin;
    const V2 *spec = reinterpret_cast<const V2 *>(d_spec);
    unsigned char *recs = reinterpret_cast<unsigned char *>(d_rec);
    unsigned char *ws = reinterpret_cast<unsigned char *>(d_mag_ws);
    const unsigned grid = list ? (unsigned)std::min<int64_t>(batch, 2 * (int64_t)ctx->sm_count) : (unsigned)batch;
    if (in_smem) {
        auto kern = flexible ? peaks_kernel<T, true, true> : peaks_kernel<T, true, false>;
        APDA_FUNC_SMEM(ctx, kern, lay.bytes);
        kern<<<grid, 256, lay.bytes, st>>>(spec, n, half, batch, fs, d_fs, k, rec_cap, recs, nullptr, list);
    } else {
        if (!ws) {
            apda_set_error("launch_peaks: magnitude workspace missing for n=%lld", (long long)n);
            return APDA_ERR_INVALID;
        }
        auto kern = flexible ? peaks_kernel<T, false, true> : peaks_kernel<T, false, false>;
        kern<<<grid, 1024, 0, st>>>(spec, n, half, batch, fs, d_fs, k, rec_cap, recs, ws, list);
    }
    ctx->launches++;
    APDA_CUDA(cudaGetLastError());
    return APDA_OK;
}

template <typename T>
int launch_peaks_general_listed(apda_ctx *ctx, cudaStream_t st, const T *d_spec, int64_t n, int64_t batch, double fs,
                                const double *d_fs, int k, int rec_cap, int flexible, void *d_rec, const int *list) {
    return launch_peaks_general<T>(ctx, st, d_spec, n, batch, fs, d_fs, k, rec_cap, flexible, d_rec, nullptr, list);
}
template int launch_peaks_general_listed<float>(apda_ctx *, cudaStream_t, const float *, int64_t, int64_t, double,
                                                const double *, int, int, int, void *, const int *);
template int launch_peaks_general_listed<double>(apda_ctx *, cudaStream_t, const double *, int64_t, int64_t, double,
                                                 const double *, int, int, int, void *, const int *);

template <typename T>
int launch_peaks(apda_ctx *ctx, cudaStream_t st, const T *d_spec, int64_t n, int64_t batch, double fs,
                 const double *d_fs, int k, int rec_cap, int flexible, void *d_rec, void *d_mag_ws) {
    using V2 = typename vec2<T>::type;
    if (sizeof(T) == 4 && !ctx->generic_only && peaks_f32_fast_supports(n, k, rec_cap))
        return launch_peaks_f32_fast(ctx, st, reinterpret_cast<const float *>(d_spec), n, batch, fs, d_fs, k, flexible,
                                     d_rec);
    if (sizeof(T) == 8 && !ctx->generic_only && peaks_f64_fast_supports(n, k, rec_cap))
        return launch_peaks_f64_fast(ctx, st, reinterpret_cast<const double *>(d_spec), n, batch, fs, d_fs, k, flexible,
                                     d_rec);
    if (peaks_large_supports(n) && !ctx->generic_only && d_mag_ws)
        return launch_peaks_large<T>(ctx, st, d_spec, n, batch, fs, d_fs, k, rec_cap, flexible, d_rec, d_mag_ws);
    return launch_peaks_general<T>(ctx, st, d_spec, n, batch, fs, d_fs, k, rec_cap, flexible, d_rec, d_mag_ws, nullptr);
}
template int launch_peaks<double>(apda_ctx *, cudaStream_t, const double *, int64_t, int64_t, double, const double *, int,
                                  int, int, void *, void *);
template int launch_peaks<float>(apda_ctx *, cudaStream_t, const float *, int64_t, int64_t, double, const double *, int,
                                 int, int, void *, void *);

int launch_mag_helpers_f64(apda_ctx *ctx, cudaStream_t st, const double *d_mags, int64_t n, int64_t idx, double prom_in,
                           double *d_out3) {
    mag_helpers_kernel<<<1, 32, 0, st>>>(d_mags, (int)n, (int)idx, prom_in, d_out3);
    ctx->launches++;
    APDA_CUDA(cudaGetLastError());
    return APDA_OK;
}
