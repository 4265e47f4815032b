// K3 (large form) for power-of-two spectra with n >= 2^16 bins: the picker of one long window spread over the chip.
//
//   A  mags_kernel      one warp per block of 1024 bins: magnitudes -> HBM, block max / min, double-double partial sums
//   B  stats_kernel     one CTA: mean, sample sigma, threshold (same correctly rounded results as peaks.cu), and the
//                       second summary level (max / min of every 32 blocks)
//   C  hot_kernel       one CTA per block whose max exceeds the threshold: strict local maxima (flexible) or every
//                       bin above the threshold (rigid) -> candidate list
//   D  eval_kernel      (flexible) one warp per candidate, grid-wide: prominence, width, damping gate -> found list.
//                       A prominence walk moves through three levels (bins, 1024-bin blocks, 32-block groups), 128
//                       entries per step with four independent loads per lane, so a walk over a 2^23-bin half
//                       spectrum is a handful of dependent memory round trips instead of hundreds
//   E  pick_*_kernel    one CTA of 32 warps: the reference's ordering / exclusion logic on the found (flexible) or
//                       candidate (rigid) list
//
// Decision semantics are those of peaks.cu (utils/get_peak_prominence.py:149-226, utils/get_peak_resolution.py:80-128).
#include <algorithm>
#include <type_traits>

#include "common.cuh"
#include "peaks_common.cuh"

namespace {

constexpr int SB = 1024;  // bins per summary block

// scratch pointer of window blockIdx.y of a launch set: every window has its own slab, `stride` bytes apart
template <typename P>
__device__ __forceinline__ P *win_slab(P *p, size_t stride) {
    return reinterpret_cast<P *>(reinterpret_cast<char *>(const_cast<typename std::remove_const<P>::type *>(p)) + blockIdx.y * stride);
}

struct LargeState {
    double mean, sd, thr;
    int ncand, nfound, overflow, tie;
};

// WPB warps per 1024-bin block: 1 for long spectra (block max / min need no shared memory, eight loads in flight per
// lane), 8 for short ones (one block per CTA, so that a 2^16-point spectrum still spreads over 32 CTAs).  A CTA's eight
// warps fold their double-double sums into one partial.
constexpr int MAGS_WARPS = 8;
template <typename T, int WPB>
__global__ void __launch_bounds__(32 * MAGS_WARPS)
mags_kernel(const typename vec2<T>::type *__restrict__ spec, T *__restrict__ mags, T *__restrict__ bmax,
            T *__restrict__ bmin, dd *__restrict__ part, int nblk, int64_t n, size_t stride) {
    constexpr int SEG = SB / WPB, PER_LANE = SEG / 32, UNR = PER_LANE < 8 ? PER_LANE : 8, BPC = MAGS_WARPS / WPB;
    __shared__ dd red[2 * MAGS_WARPS];
    __shared__ T rmx[MAGS_WARPS], rmn[MAGS_WARPS];
    pdl_trigger();
    pdl_wait();  // the spectrum usually comes from the kernel in front (K2's last pass)
    spec += (int64_t)blockIdx.y * n;
    mags = win_slab(mags, stride);
    bmax = win_slab(bmax, stride);
    bmin = win_slab(bmin, stride);
    part = win_slab(part, stride);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    dd sx[2] = {{0.0, 0.0}, {0.0, 0.0}}, sxx[2] = {{0.0, 0.0}, {0.0, 0.0}};  // two independent chains per lane
    for (int blk = blockIdx.x * BPC + warp / WPB; blk < nblk; blk += gridDim.x * BPC) {
        const int64_t b0 = (int64_t)blk * SB + (warp % WPB) * SEG;
        T mx = T(0), mn = T(0);
#pragma unroll 1
        for (int r = 0; r < PER_LANE / UNR; ++r) {
            typename vec2<T>::type v[UNR];
#pragma unroll
            for (int u = 0; u < UNR; ++u) v[u] = spec[b0 + 32 * UNR * r + 32 * u + lane];
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
                const T m = magnitude_fast(v[u].x, v[u].y);
                mags[b0 + 32 * UNR * r + 32 * u + lane] = m;
                const bool first = r == 0 && u == 0;
                mx = (first || m > mx) ? m : mx;
                mn = (first || m < mn) ? m : mn;
                const double d = (double)m;
                if (sizeof(T) == 8) {
                    // non-negative summands: TwoSum on the high words, plain adds on the error words (exact to ~2^-94)
                    const dd s = two_sum(sx[u & 1].hi, d);
                    sx[u & 1].hi = s.hi;
                    sx[u & 1].lo = add_rn(sx[u & 1].lo, s.lo);
                    const dd pr = two_prod(d, d);
                    const dd s2 = two_sum(sxx[u & 1].hi, pr.hi);
                    sxx[u & 1].hi = s2.hi;
                    sxx[u & 1].lo = add_rn(add_rn(sxx[u & 1].lo, pr.lo), s2.lo);
                } else {
                    sx[u & 1].hi += d;
                    sxx[u & 1].hi = __fma_rn(d, d, sxx[u & 1].hi);
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const T a = __shfl_xor_sync(0xffffffffu, mx, o), b = __shfl_xor_sync(0xffffffffu, mn, o);
            mx = a > mx ? a : mx;
            mn = b < mn ? b : mn;
        }
        if (WPB == 1) {
            if (lane == 0) {
                bmax[blk] = mx;
                bmin[blk] = mn;
            }
        } else {  // BPC == 1: every warp of the CTA is in this iteration
            if (lane == 0) {
                rmx[warp] = mx;
                rmn[warp] = mn;
            }
            __syncthreads();
            if (tid == 0) {
                for (int w = 1; w < MAGS_WARPS; ++w) {
                    mx = rmx[w] > mx ? rmx[w] : mx;
                    mn = rmn[w] < mn ? rmn[w] : mn;
                }
                bmax[blk] = mx;
                bmin[blk] = mn;
            }
            __syncthreads();
        }
    }
    const dd wx = warp_sum_dd(dd_add(two_sum(sx[0].hi, sx[0].lo), two_sum(sx[1].hi, sx[1].lo)));
    const dd wxx = warp_sum_dd(dd_add(two_sum(sxx[0].hi, sxx[0].lo), two_sum(sxx[1].hi, sxx[1].lo)));
    if (lane == 0) {
        red[warp] = wx;
        red[MAGS_WARPS + warp] = wxx;
    }
    __syncthreads();
    if (tid == 0) {
        dd a = red[0], b = red[MAGS_WARPS];
        for (int w = 1; w < MAGS_WARPS; ++w) {
            a = dd_add(a, red[w]);
            b = dd_add(b, red[MAGS_WARPS + w]);
        }
        part[2 * blockIdx.x] = a;
        part[2 * blockIdx.x + 1] = b;
    }
}

// npart <= 1024 partial sums and nblk block summaries -> mean / sigma / threshold and the 32-block group summaries
template <typename T>
__global__ void __launch_bounds__(1024)
stats_kernel(const dd *__restrict__ part, int npart, const T *__restrict__ bmax, const T *__restrict__ bmin,
             T *__restrict__ smax, T *__restrict__ smin, int nblk, int half, LargeState *st, size_t stride) {
    __shared__ dd red[64];
    pdl_trigger();
    pdl_wait();
    part = win_slab(part, stride);
    bmax = win_slab(bmax, stride);
    bmin = win_slab(bmin, stride);
    smax = win_slab(smax, stride);
    smin = win_slab(smin, stride);
    st = win_slab(st, stride);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < nblk; i += 1024) {  // nblk is a multiple of 32: a warp's lanes hold one group of 32 blocks
        T mx = bmax[i], mn = bmin[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const T x = __shfl_xor_sync(0xffffffffu, mx, o), y = __shfl_xor_sync(0xffffffffu, mn, o);
            mx = x > mx ? x : mx;
            mn = y < mn ? y : mn;
        }
        if (lane == 0) {
            smax[i >> 5] = mx;
            smin[i >> 5] = mn;
        }
    }
    dd a = {0.0, 0.0}, b = {0.0, 0.0};
    for (int i = tid; i < npart; i += 1024) {
        a = dd_add(a, part[2 * i]);
        b = dd_add(b, part[2 * i + 1]);
    }
    a = warp_sum_dd(a);
    b = warp_sum_dd(b);
    if (lane == 0) {
        red[warp] = a;
        red[32 + warp] = b;
    }
    __syncthreads();
    if (warp == 0) {
        a = warp_sum_dd(red[lane]);
        b = warp_sum_dd(red[32 + lane]);
    }
    if (tid == 0) {
        const double n = (double)half;
        const dd mean = dd_div_d(a, n);
        const dd ss = dd_add(b, dd_neg(dd_div_d(dd_mul(a, a), n)));
        const dd var = dd_div_d(ss, n - 1.0);
        st->mean = add_rn(mean.hi, mean.lo);
        st->sd = dd_sqrt_to_double(var);
        st->thr = add_rn(st->mean, mul_rn(2.0, st->sd));
        st->ncand = 0;
        st->nfound = 0;
        st->overflow = 0;
        st->tie = 0;
    }
}

template <typename T, bool FLEX>
__global__ void __launch_bounds__(256)
hot_kernel(const T *__restrict__ mags, const T *__restrict__ bmax, int half, LargeState *st, int *__restrict__ cand, int cap,
           size_t stride) {
    pdl_trigger();
    pdl_wait();
    mags = win_slab(mags, stride);
    bmax = win_slab(bmax, stride);
    st = win_slab(st, stride);
    cand = win_slab(cand, stride);
    const double thr = st->thr;
    if (!((double)bmax[blockIdx.x] > thr)) return;
    const int b0 = blockIdx.x * SB, lane = threadIdx.x & 31;
    for (int u = 0; u < SB / 256; ++u) {
        const int j = b0 + threadIdx.x + 256 * u;
        const T m = mags[j];
        bool take = (double)m > thr;
        if (take) {
            if (sizeof(T) == 4 && j >= 1 && j <= half - 2 && fp32_tie_top(mags, half, j, m, thr)) st->tie = 1;
            if (FLEX) take = j >= 1 && j <= half - 2 && m > mags[j - 1] && m > mags[j + 1];
        }
        const unsigned takers = __ballot_sync(0xffffffffu, take);  // one counter update per warp
        if (takers) {
            int pos = 0;
            if (lane == __ffs(takers) - 1) pos = atomicAdd(&st->ncand, __popc(takers));
            pos = __shfl_sync(0xffffffffu, pos, __ffs(takers) - 1) + __popc(takers & ((1u << lane) - 1u));
            if (take) {
                if (pos < cap) cand[pos] = j;
                else st->overflow = 1;
            }
        }
    }
}

// One direction of a prominence walk over the entries [lo, hi] of one summary level: an entry with mx[i] > p stops the
// walk, mn[i] of the entries in front of it lowers the floor (level 0: mx == mn == magnitudes).  128 entries per step,
// four independent loads per lane.  Returns the stopping entry, or -1 when the range ends first (warp-uniform).
template <typename T, int DIR, bool SAME>
__device__ __forceinline__ int walk_level(const T *__restrict__ mx, const T *__restrict__ mn, int lo, int hi, T p, T &floor_lane,
                                          int lane) {
    constexpr int U = 4;
    for (int base = DIR < 0 ? hi : lo; DIR < 0 ? base >= lo : base <= hi; base += DIR * 32 * U) {
        T vx[U], vn[U];
        bool valid[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i = base + DIR * (lane + 32 * u);
            valid[u] = DIR < 0 ? i >= lo : i <= hi;
            vx[u] = valid[u] ? mx[i] : p;
            vn[u] = SAME ? vx[u] : (valid[u] ? mn[i] : p);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const unsigned higher = __ballot_sync(0xffffffffu, valid[u] && vx[u] > p);
            const int stop = higher ? (__ffs(higher) - 1) : 32;
            if (lane < stop && vn[u] < floor_lane) floor_lane = vn[u];
            if (higher) return base + DIR * (stop + 32 * u);
        }
    }
    return -1;
}

// utils/get_peak_prominence.py:32-54 over three levels: bins of the peak's own block, blocks of its own group, groups;
// then back down into the group and the block where the walk stops
template <typename T>
__device__ T summary_prominence(const T *mags, const T *bmax, const T *bmin, const T *smax, const T *smin, int half, int j,
                                int lane) {
    const T p = mags[j];
    const int nblk = half / SB, bj = j / SB, sj = bj >> 5, nsup = nblk >> 5;
    T fl = p, fr = p;
    if (walk_level<T, -1, true>(mags, mags, bj * SB, j - 1, p, fl, lane) < 0) {
        int b = walk_level<T, -1, false>(bmax, bmin, sj * 32, bj - 1, p, fl, lane);
        if (b < 0) {
            const int s = walk_level<T, -1, false>(smax, smin, 0, sj - 1, p, fl, lane);
            if (s >= 0) b = walk_level<T, -1, false>(bmax, bmin, s * 32, s * 32 + 31, p, fl, lane);
        }
        if (b >= 0) walk_level<T, -1, true>(mags, mags, b * SB, b * SB + SB - 1, p, fl, lane);
    }
    if (walk_level<T, +1, true>(mags, mags, j + 1, bj * SB + SB - 1, p, fr, lane) < 0) {
        int b = walk_level<T, +1, false>(bmax, bmin, bj + 1, sj * 32 + 31, p, fr, lane);
        if (b < 0) {
            const int s = walk_level<T, +1, false>(smax, smin, sj + 1, nsup - 1, p, fr, lane);
            if (s >= 0) b = walk_level<T, +1, false>(bmax, bmin, s * 32, s * 32 + 31, p, fr, lane);
        }
        if (b >= 0) walk_level<T, +1, true>(mags, mags, b * SB, b * SB + SB - 1, p, fr, lane);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const T a = __shfl_xor_sync(0xffffffffu, fl, o), b = __shfl_xor_sync(0xffffffffu, fr, o);
        fl = a < fl ? a : fl;
        fr = b < fr ? b : fr;
    }
    return sub_rn(p, fl > fr ? fl : fr);
}

// flexible picker, per-candidate half: prominence gate, half-power width, damping gate (utils/get_peak_prominence.py:171-204)
constexpr int EVAL_THREADS = 256;
template <typename T>
__global__ void __launch_bounds__(EVAL_THREADS)
eval_kernel(const T *__restrict__ mags, const T *__restrict__ bmax, const T *__restrict__ bmin, const T *__restrict__ smax,
            const T *__restrict__ smin, int64_t n, int half, double fs_all, const double *__restrict__ fs_ptr, LargeState *st,
            const int *__restrict__ cand, Found *__restrict__ found, int cap, size_t stride) {
    pdl_trigger();
    pdl_wait();
    mags = win_slab(mags, stride);
    bmax = win_slab(bmax, stride);
    bmin = win_slab(bmin, stride);
    smax = win_slab(smax, stride);
    smin = win_slab(smin, stride);
    st = win_slab(st, stride);
    cand = win_slab(cand, stride);
    found = win_slab(found, stride);
    if (fs_ptr) fs_ptr += blockIdx.y;
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * EVAL_THREADS + threadIdx.x) >> 5, nwarp = (gridDim.x * EVAL_THREADS) >> 5;
    const int ncand = min(st->ncand, cap);
    const double fs = fs_ptr ? *fs_ptr : fs_all;
    const double df = div_rn(fs, (double)n);
    const double half_sd = mul_rn(0.5, st->sd);
    for (int c = warp; c < ncand; c += nwarp) {
        const int j = cand[c];
        const T prom = summary_prominence<T>(mags, bmax, bmin, smax, smin, half, j, lane);
        if (!((double)prom > half_sd)) continue;
        const int bins = half_power_bins<T>(mags, half, prom, j);
        const double width_hz = mul_rn((double)bins, df);
        if (!(width_hz > 0.0)) continue;
        const double fn = mul_rn((double)j, df);
        const double q = div_rn(fn, width_hz);
        const double damping = div_rn(1.0, mul_rn(2.0, q));
        if (0.001 <= damping && damping <= 0.07 && lane == 0) {
            const int pos = atomicAdd(&st->nfound, 1);
            Found f;
            f.rmag = round_dec4((double)mags[j]);
            f.prom = (double)prom;
            f.idx = j;
            f.width = bins;
            found[pos] = f;
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(1024)
pick_flexible_kernel(const T *__restrict__ mags, int64_t n, double fs_all, const double *__restrict__ fs_ptr, int k, int rec_cap,
                     LargeState *st, const Found *__restrict__ found, int *__restrict__ acc_slot, unsigned char *__restrict__ rec,
                     size_t stride) {
    pdl_trigger();
    pdl_wait();
    mags = win_slab(mags, stride);
    st = win_slab(st, stride);
    found = win_slab(found, stride);
    acc_slot = win_slab(acc_slot, stride);
    rec += (size_t)blockIdx.y * APDA_REC_BYTES(rec_cap);
    if (fs_ptr) fs_ptr += blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double fs = fs_ptr ? *fs_ptr : fs_all;
    const double df = div_rn(fs, (double)n);
    // order by (rounded magnitude desc, idx asc) one element at a time; block-wide arg-max per step
    __shared__ double bk[32];
    __shared__ int bi[32], be[32];
    __shared__ int sel_e, na_s;
    __shared__ double prev_mag_s;
    __shared__ int prev_idx_s;
    if (tid == 0) {
        na_s = 0;
        prev_mag_s = CUDART_INF;
        prev_idx_s = -1;
    }
    __syncthreads();
    const int nfound = st->nfound;
    while (true) {
        const double prev_mag = prev_mag_s;
        const int prev_idx = prev_idx_s;
        double best = -1.0;
        int best_idx = 0x7fffffff, best_e = -1;
        for (int e = tid; e < nfound; e += 1024) {
            const double r = found[e].rmag;
            const int ix = found[e].idx;
            const bool after_prev = r < prev_mag || (r == prev_mag && ix > prev_idx);
            if (after_prev && (r > best || (r == best && ix < best_idx))) {
                best = r;
                best_idx = ix;
                best_e = e;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double r = __shfl_xor_sync(0xffffffffu, best, o);
            const int ix = __shfl_xor_sync(0xffffffffu, best_idx, o);
            const int e = __shfl_xor_sync(0xffffffffu, best_e, o);
            if (e >= 0 && (best_e < 0 || r > best || (r == best && ix < best_idx))) {
                best = r;
                best_idx = ix;
                best_e = e;
            }
        }
        if (lane == 0) {
            bk[warp] = best;
            bi[warp] = best_idx;
            be[warp] = best_e;
        }
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < 32; ++w) {
                if (be[w] >= 0 && (best_e < 0 || bk[w] > best || (bk[w] == best && bi[w] < best_idx))) {
                    best = bk[w];
                    best_idx = bi[w];
                    best_e = be[w];
                }
            }
            sel_e = best_e;
            if (best_e >= 0) {
                prev_mag_s = best;
                prev_idx_s = best_idx;
                const int na = na_s;
                const double cf = round_dec4(mul_rn((double)best_idx, df));
                bool hump = false;
                for (int a = 0; a < na && !hump; ++a) {
                    const double af = round_dec4(mul_rn((double)found[acc_slot[a]].idx, df));
                    const double rel = div_rn(fabs(sub_rn(cf, af)), af);
                    if (rel < 0.05 && div_rn(found[best_e].prom, best) < 0.10) hump = true;
                }
                if (!hump) {
                    acc_slot[na] = best_e;
                    na_s = na + 1;
                }
            }
        }
        __syncthreads();
        if (sel_e < 0 || na_s >= k) break;
    }
    if (tid == 0) {
        const int na = na_s;
        write_rec_header(rec, na, (st->overflow ? APDA_STATUS_TRUNCATED : 0) | (st->tie ? APDA_STATUS_FP32_TIE : 0));
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

template <typename T>
__global__ void __launch_bounds__(1024)
pick_rigid_kernel(T *__restrict__ mags, int64_t n, int half, double fs_all, const double *__restrict__ fs_ptr, int k, int rec_cap, LargeState *st,
                  const int *__restrict__ cand, int cap, int *__restrict__ acc_idx, unsigned char *__restrict__ rec, size_t stride) {
    __shared__ double best_m[32];
    __shared__ int best_j[32];
    __shared__ int ctl[4];
    pdl_trigger();
    pdl_wait();
    mags = win_slab(mags, stride);
    st = win_slab(st, stride);
    cand = win_slab(cand, stride);
    acc_idx = win_slab(acc_idx, stride);
    rec += (size_t)blockIdx.y * APDA_REC_BYTES(rec_cap);
    if (fs_ptr) fs_ptr += blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nhot = min(st->ncand, cap);
    const double thr = st->thr;
    const double fs = fs_ptr ? *fs_ptr : fs_all;
    const double df = div_rn(fs, (double)n);
    const double distance = sub_rn(mul_rn(2.0, df), mul_rn(1.0, df));
    if (tid == 0) ctl[3] = 0;
    __syncthreads();
    while (true) {
        double bm = -1.0;
        int bj = -1;
        // four candidates in flight per thread; the neighbours are only read for a bin that would beat the thread's best
        for (int e0 = tid; e0 < nhot; e0 += 4 * 1024) {
            int js[4];
            T ms[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) js[u] = e0 + 1024 * u < nhot ? cand[e0 + 1024 * u] : -1;
#pragma unroll
            for (int u = 0; u < 4; ++u) ms[u] = js[u] >= 0 ? mags[js[u]] : T(0);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int j = js[u];
                const T m = ms[u];
                if (j >= 1 && j <= half - 2 && (double)m > thr && ((double)m > bm || ((double)m == bm && j < bj)) &&
                    m > mags[j - 1] && m > mags[j + 1]) {
                    bm = (double)m;
                    bj = j;
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double m2 = __shfl_xor_sync(0xffffffffu, bm, o);
            const int j2 = __shfl_xor_sync(0xffffffffu, bj, o);
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
            for (int w = 1; w < 32; ++w) {
                if (best_j[w] >= 0 && (bj < 0 || best_m[w] > bm || (best_m[w] == bm && best_j[w] < bj))) {
                    bm = best_m[w];
                    bj = best_j[w];
                }
            }
            ctl[0] = bj;
            if (bj >= 0) {
                const int na = ctl[3];
                const double f = mul_rn((double)bj, df);
                const int w2 = half_height_bins<T>(mags, half, bj);
                bool separated = true;
                for (int a = 0; a < na && separated; ++a) {
                    const int w1 = half_height_bins<T>(mags, half, acc_idx[a]);
                    double rs = 0.0;
                    if (w1 + w2 != 0) rs = div_rn(mul_rn(1.18, (double)abs(bj - acc_idx[a])), (double)(w1 + w2));
                    if (!(rs >= 1.5)) separated = false;
                }
                if (separated) {
                    acc_idx[na] = bj;
                    write_rec_peak(rec, na, bj, w2, bm, 0.0);
                    ctl[3] = na + 1;
                }
                double reach_d = rint(div_rn(mul_rn(f, 0.02), distance));
                if (!(reach_d >= 0.0)) reach_d = 0.0;
                if (reach_d > (double)half) reach_d = (double)half;
                const int reach = (int)reach_d;
                ctl[1] = max(0, bj - reach);
                ctl[2] = min(half, bj + reach + 1);
            }
        }
        __syncthreads();
        if (ctl[0] < 0) break;
        for (int j = ctl[1] + tid; j < ctl[2]; j += 1024) mags[j] = T(0);
        const bool full = ctl[3] >= k;
        __threadfence_block();
        __syncthreads();
        if (full) break;
    }
    if (tid == 0) {
        const int na = ctl[3];
        write_rec_header(rec, na, (st->overflow ? APDA_STATUS_TRUNCATED : 0) | (st->tie ? APDA_STATUS_FP32_TIE : 0));
        for (int a = na; a < rec_cap; ++a) write_rec_peak(rec, a, -1, 0, 0.0, 0.0);
    }
}

struct LargeLayout {
    size_t mags, bmax, bmin, smax, smin, part, state, cand, found, acc, bytes;
    int cap, nblk;
};
template <typename T>
LargeLayout large_layout(int64_t half) {
    LargeLayout l;
    l.nblk = (int)(half / SB);
    l.cap = (int)std::min<int64_t>(half / 4 + 8, 1 << 20);
    size_t o = 0;
    auto take = [&](size_t bytes) {
        size_t at = o;
        o += (bytes + 255) & ~(size_t)255;
        return at;
    };
    l.mags = take((size_t)half * sizeof(T));
    l.bmax = take((size_t)l.nblk * sizeof(T));
    l.bmin = take((size_t)l.nblk * sizeof(T));
    l.smax = take((size_t)(l.nblk / 32) * sizeof(T));
    l.smin = take((size_t)(l.nblk / 32) * sizeof(T));
    l.part = take((size_t)1024 * 2 * sizeof(dd));
    l.state = take(sizeof(LargeState));
    l.cand = take((size_t)l.cap * sizeof(int));
    l.found = take((size_t)l.cap * sizeof(Found));
    l.acc = take((size_t)l.cap * sizeof(int));  // accepted peaks (any k up to cap)
    l.bytes = o;
    return l;
}

}  // namespace

bool peaks_large_supports(int64_t n) { return is_pow2_i64(n) && n >= (int64_t(1) << 16) && n <= (int64_t(1) << 31); }

// windows of one call that share a launch set (blockIdx.y), each with its own scratch slab: a batch of 2^16 ... 2^18-point
// spectra is launch bound otherwise (five dependent launches per window)
template <typename T>
static int64_t large_windows(int64_t n, int64_t batch) {
    const int64_t slab = (int64_t)large_layout<T>(n / 2).bytes;
    const int64_t fit = std::max<int64_t>(1, ((int64_t)192 << 20) / slab);
    return std::max<int64_t>(1, std::min<int64_t>(batch, std::min<int64_t>(fit, 64)));
}

template <typename T>
size_t peaks_large_workspace_bytes(int64_t n, int64_t batch) {
    return large_layout<T>(n / 2).bytes * (size_t)large_windows<T>(n, batch);
}
template size_t peaks_large_workspace_bytes<double>(int64_t, int64_t);
template size_t peaks_large_workspace_bytes<float>(int64_t, int64_t);

// ws: peaks_large_workspace_bytes(n, batch') bytes for some batch' >= batch
template <typename T>
int launch_peaks_large(apda_ctx *ctx, cudaStream_t st, const T *d_spec, int64_t n, int64_t batch, double fs,
                       const double *d_fs, int k, int rec_cap, int flexible, void *d_rec, void *ws) {
    using V2 = typename vec2<T>::type;
    const int64_t half = n / 2;
    const LargeLayout l = large_layout<T>(half);
    const size_t stride = l.bytes;
    const int64_t wins = large_windows<T>(n, batch);
    char *base = reinterpret_cast<char *>(ws);
    T *mags = reinterpret_cast<T *>(base + l.mags);
    T *bmax = reinterpret_cast<T *>(base + l.bmax), *bmin = reinterpret_cast<T *>(base + l.bmin);
    T *smax = reinterpret_cast<T *>(base + l.smax), *smin = reinterpret_cast<T *>(base + l.smin);
    dd *part = reinterpret_cast<dd *>(base + l.part);
    LargeState *state = reinterpret_cast<LargeState *>(base + l.state);
    int *cand = reinterpret_cast<int *>(base + l.cand);
    Found *found = reinterpret_cast<Found *>(base + l.found);
    int *acc = reinterpret_cast<int *>(base + l.acc);
    // one warp per candidate; tone spectra have a handful, noise-like ones ~2 % of the bins.  The chip's resident CTAs are
    // shared among the windows of a launch set (tens of thousands of CTAs that find no candidate would cost more than the walks)
    const int64_t eval_max = std::max<int64_t>(1, l.cap / (EVAL_THREADS / 32));
    const int eval_ctas = (int)std::min<int64_t>(eval_max, std::max<int64_t>(4, (int64_t)ctx->sm_count * 8 / wins));
    const int mags_ctas = l.nblk <= 1024 ? l.nblk : std::min(l.nblk / MAGS_WARPS, 1024);  // <= 1024 partial sums
    for (int64_t w0 = 0; w0 < batch; w0 += wins) {
        const unsigned wy = (unsigned)std::min<int64_t>(wins, batch - w0);
        const V2 *spec = reinterpret_cast<const V2 *>(d_spec) + w0 * n;
        unsigned char *rec = reinterpret_cast<unsigned char *>(d_rec) + w0 * APDA_REC_BYTES(rec_cap);
        const double *fs_ptr = d_fs ? d_fs + w0 : nullptr;
        const dim3 mg(mags_ctas, wy), mb(32 * MAGS_WARPS), hg(l.nblk, wy), one(1, wy), eg(eval_ctas, wy);
        if (l.nblk <= 1024)
            APDA_CUDA(apda_launch_pdl(APDA_PDL_K3, mags_kernel<T, MAGS_WARPS>, mg, mb, 0, st, spec, mags, bmax, bmin, part, l.nblk, n, stride));
        else
            APDA_CUDA(apda_launch_pdl(APDA_PDL_K3, mags_kernel<T, 1>, mg, mb, 0, st, spec, mags, bmax, bmin, part, l.nblk, n, stride));
        APDA_CUDA(apda_launch_pdl(APDA_PDL_K3, stats_kernel<T>, one, dim3(1024), 0, st, (const dd *)part, mags_ctas, (const T *)bmax,
                                  (const T *)bmin, smax, smin, l.nblk, (int)half, state, stride));
        if (flexible) {
            APDA_CUDA(apda_launch_pdl(APDA_PDL_K3, hot_kernel<T, true>, hg, dim3(256), 0, st, (const T *)mags, (const T *)bmax, (int)half, state, cand,
                                      l.cap, stride));
            APDA_CUDA(apda_launch_pdl(APDA_PDL_K3, eval_kernel<T>, eg, dim3(EVAL_THREADS), 0, st, (const T *)mags, (const T *)bmax,
                                      (const T *)bmin, (const T *)smax, (const T *)smin, n, (int)half, fs, fs_ptr, state,
                                      (const int *)cand, found, l.cap, stride));
            APDA_CUDA(apda_launch_pdl(APDA_PDL_K3, pick_flexible_kernel<T>, one, dim3(1024), 0, st, (const T *)mags, n, fs, fs_ptr, k, rec_cap,
                                      state, (const Found *)found, acc, rec, stride));
            ctx->launches += 1;
        } else {
            APDA_CUDA(apda_launch_pdl(APDA_PDL_K3, hot_kernel<T, false>, hg, dim3(256), 0, st, (const T *)mags, (const T *)bmax, (int)half, state, cand,
                                      l.cap, stride));
            APDA_CUDA(apda_launch_pdl(APDA_PDL_K3, pick_rigid_kernel<T>, one, dim3(1024), 0, st, mags, n, (int)half, fs, fs_ptr, k, rec_cap, state,
                                      (const int *)cand, l.cap, acc, rec, stride));
        }
        ctx->launches += 4;
    }
    APDA_CUDA(cudaGetLastError());
    return APDA_OK;
}
template int launch_peaks_large<double>(apda_ctx *, cudaStream_t, const double *, int64_t, int64_t, double, const double *,
                                        int, int, int, void *, void *);
template int launch_peaks_large<float>(apda_ctx *, cudaStream_t, const float *, int64_t, int64_t, double, const double *, int,
                                       int, int, void *, void *);
