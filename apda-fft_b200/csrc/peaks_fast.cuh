// K3 (fast form) for n = 1024 / 2048 / 4096 / 8192 bins per window, k <= 5, 128-byte records: the headline picker.
// Templated on the magnitude type T: float (peaks_f32_fast.cu, fused_f32.cu) and double (peaks_f64_fast.cu).
//
// One WARP per window, four windows per CTA, no block-level barrier.
//   phase 1  32 lanes stream the half spectrum with coalesced 128-bit loads (8 in flight per lane), take magnitudes,
//            accumulate sum / sum of squares in registers and stage the magnitudes in shared memory;
//   phase 2  every lane re-reads one CONTIGUOUS chunk of HALF/32 bins (conflict-free 128-bit LDS thanks to a 4-word
//            pad per chunk), keeps the chunk's max and min in registers and appends the few bins above
//            mean + 2 sigma ("hot" bins, < 20 % of the bins by Cantelli's inequality, typically ~10) to a slot list;
//   flexible picker: for every hot strict local maximum the prominence walk runs on the chunk summaries - whole
//            chunks are skipped with one ballot over the lanes' chunk maxima, only the two boundary chunks are
//            scanned (warp-cooperatively, 32 bins per step); the scalar epilogue (half-power width, damping gate,
//            decimal rounding) runs lane-parallel, one candidate per lane; a warp arg-max extracts the order
//            "descending round(mag, 4), ascending idx" for the greedy hump exclusion;
//   rigid picker: the iterative arg-max / resolution test / +-2 % zeroing loop works on the hot list only (zeroing
//            can only create new maxima among bins that were already above the threshold).
// The record (128 B) is assembled in shared memory and written with one coalesced 128-byte store.
//
// Decision semantics are those of peaks.cu (the general kernel), which documents the reference line by line:
//   utils/get_peak_prominence.py:149-226, utils/get_peak_resolution.py:80-128.
#pragma once
#include "common.cuh"

namespace {

#ifndef APDA_K3_WPC
#define APDA_K3_WPC 2
#endif
#ifndef APDA_K3_LAZY_MIN
#define APDA_K3_LAZY_MIN 8  // flexible picker: more candidates than this -> evaluate in output order until k are accepted
#endif
constexpr int kWPC = APDA_K3_WPC;  // windows (warps) per CTA

template <typename T>
struct SlotT {
    uint16_t idx;
    uint16_t width;  // flexible: half-power bins if the candidate passed every gate, else 0
    T prom;
};
using Slot = SlotT<float>;

template <typename T, int HALF>
struct K3 {
    static constexpr int C = HALF / 32;                // bins per lane chunk
    static constexpr int PAD = 16 / (int)sizeof(T);    // 16-byte pad per chunk: the lanes' 128-bit LDS are conflict free
    static constexpr int MAGW = HALF + PAD * 32;       // magnitude elements incl. the pads
    // candidates (flexible) kept on chip; more -> repair list (general kernel).  160 for the longest windows (a pure-noise
    // spectrum of 4096 bins has ~95 +- 10 hot local maxima) costs no residency: 6 CTAs of 2 windows per SM either way.
    // The rigid picker keeps bare 16-bit bin indices in the same bytes: 4 (fp32) or 8 (fp64) times as many hot bins.
    static constexpr int SLOTS = HALF >= 4096 ? 160 : 96;
    static constexpr int HOT_CAP = SLOTS * (int)sizeof(SlotT<T>) / 2;
    static constexpr int REC_OFF = MAGW * (int)sizeof(T) + SLOTS * (int)sizeof(SlotT<T>);
    static constexpr int BYTES = (REC_OFF + 128 + 15) & ~15;
    __device__ static __forceinline__ int addr(int b) { return b + PAD * (b / C); }
    // element offset of the 64-bin row R (R = r0 + u with r0 a multiple of 8: row_off is additive in that split)
    __device__ static __forceinline__ int row_off(int R) { return 64 * R + (C <= 64 ? PAD * (64 / C) * R : PAD * (R / (C / 64))); }
};

// type-generic spellings of the few operations that differ between the two magnitude types
__device__ __forceinline__ float vmin(float a, float b) { return fminf(a, b); }
__device__ __forceinline__ float vmax(float a, float b) { return fmaxf(a, b); }
__device__ __forceinline__ double vmin(double a, double b) { return fmin(a, b); }
__device__ __forceinline__ double vmax(double a, double b) { return fmax(a, b); }
// warp minimum of non-negative magnitudes; fp32: the bit patterns order like the values, one REDUX instruction
__device__ __forceinline__ float warp_min_nonneg(float v) {
    return __uint_as_float(__reduce_min_sync(0xffffffffu, __float_as_uint(v)));
}
// fp64: non-negative doubles order like their 64-bit patterns - REDUX on the high words, then on the low words of the
// lanes that hold the winning high word
__device__ __forceinline__ double warp_min_nonneg(double v) {
    const unsigned hi = (unsigned)__double2hiint(v), lo = (unsigned)__double2loint(v);
    const unsigned bhi = __reduce_min_sync(0xffffffffu, hi);
    const unsigned blo = __reduce_min_sync(0xffffffffu, hi == bhi ? lo : 0xffffffffu);
    return __hiloint2double((int)bhi, (int)blo);
}
// warp arg-max of (m, j) over the lanes with j >= 0 (their m is > 0): largest m, ties -> lowest j; j = -1 if no lane has one
__device__ __forceinline__ void warp_argmax(float &m, int &j) {
    const unsigned key = j >= 0 ? __float_as_uint(m) : 0u;
    const unsigned best = __reduce_max_sync(0xffffffffu, key);
    const int jb = __reduce_min_sync(0xffffffffu, (j >= 0 && key == best) ? j : 0x7fffffff);
    m = __uint_as_float(best);
    j = best ? jb : -1;
}
__device__ __forceinline__ void warp_argmax(double &m, int &j) {
    const unsigned hi = j >= 0 ? (unsigned)__double2hiint(m) : 0u, lo = j >= 0 ? (unsigned)__double2loint(m) : 0u;
    const unsigned bhi = __reduce_max_sync(0xffffffffu, hi);
    const unsigned blo = __reduce_max_sync(0xffffffffu, hi == bhi ? lo : 0u);
    const int jb = __reduce_min_sync(0xffffffffu, (j >= 0 && hi == bhi && lo == blo) ? j : 0x7fffffff);
    m = __hiloint2double((int)bhi, (int)blo);
    j = jb == 0x7fffffff ? -1 : jb;
}
template <typename T>
struct Quad {
    T x, y, z, w;
};
__device__ __forceinline__ Quad<float> lds_quad(const float *p) {  // 4 consecutive magnitudes, 16-byte aligned
    const float4 v = *reinterpret_cast<const float4 *>(p);
    return {v.x, v.y, v.z, v.w};
}
__device__ __forceinline__ Quad<double> lds_quad(const double *p) {
    const double2 a = reinterpret_cast<const double2 *>(p)[0], b = reinterpret_cast<const double2 *>(p)[1];
    return {a.x, a.y, b.x, b.y};
}

__device__ __forceinline__ double round_dec4_d(double x) {  // exact emulation of Python round(x, 4); see peaks.cu
    const double p = 1e4;
    double hi = mul_rn(x, p);
    double lo = __fma_rn(x, p, -hi);
    double n = rint(hi);
    double d = sub_rn(hi, n);
    if (d == 0.5 && lo > 0.0) n += 1.0;
    if (d == -0.5 && lo < 0.0) n -= 1.0;
    return div_rn(n, p);
}

// MUFU.SQRT: <= 1 ulp, far inside the fp32 path's 1e-5 contract.  .ftz: one instruction instead of the five of the
// denormal-preserving form (a squared magnitude below 1.2e-38 - a bin below 1e-19 - counts as zero)
__device__ __forceinline__ float sqrt_fast(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// div_rn(x, y) < c, decided by one multiplication unless the quotient is within a few ulps of c (then by the division)
__device__ __forceinline__ bool ratio_lt(double x, double y, double c) {
    const double t = c * y;
    if (y > 0.0 && x < t * (1.0 - 0x1p-48)) return true;
    if (y > 0.0 && x > t * (1.0 + 0x1p-48)) return false;
    return div_rn(x, y) < c;
}

// integer n with round(x, 4) == n / 1e4 (see round_dec4_d); ordering by n == ordering by the rounded value
__device__ __forceinline__ double round_dec4_units(double x) {
    const double p = 1e4;
    double hi = mul_rn(x, p);
    double lo = __fma_rn(x, p, -hi);
    double n = rint(hi);
    double d = sub_rn(hi, n);
    if (d == 0.5 && lo > 0.0) n += 1.0;
    if (d == -0.5 && lo < 0.0) n -= 1.0;
    return n;
}

// Exact half of the flexible picker's "hump" test (utils/get_peak_prominence.py:199-207) for ONE accepted peak; out of
// line on purpose: the callers first run a product-only pre-filter that proves most pairs more than 5 % apart, and the
// roundings / divisions below must not be hoisted in front of that filter (they were: ~140 warp instructions per window).
static __device__ __noinline__ bool hump_exact(int c_idx, int ja, double df, double cprom, double key_units) {
    const double cf = round_dec4_d(mul_rn((double)c_idx, df)), af = round_dec4_d(mul_rn((double)ja, df));
    return div_rn(fabs(sub_rn(cf, af)), af) < 0.05 && div_rn(cprom, div_rn(key_units, 1e4)) < 0.10;
}

__device__ __forceinline__ float4 ldg_stream(const float4 *p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}

// one direction of the prominence walk inside [lo_b, hi_b] (a piece of one chunk), 32 bins per step.
// DIR = -1: from hi_b downwards, +1: from lo_b upwards.  Returns true when a bin strictly higher than p stopped it.
template <typename T, int HALF, int DIR>
__device__ __forceinline__ bool scan_piece(const T *mags, int lo_b, int hi_b, T p, T &floor_lane, int lane) {
    if (DIR < 0) {
        for (int base = hi_b; base >= lo_b; base -= 32) {
            const int i = base - lane;
            const bool valid = i >= lo_b;
            const T v = valid ? mags[K3<T, HALF>::addr(i)] : p;
            const unsigned higher = __ballot_sync(0xffffffffu, valid && v > p);
            const int stop = higher ? (__ffs(higher) - 1) : 32;
            if (lane < stop && v < floor_lane) floor_lane = v;
            if (higher) return true;
        }
    } else {
        for (int base = lo_b; base <= hi_b; base += 32) {
            const int i = base + lane;
            const bool valid = i <= hi_b;
            const T v = valid ? mags[K3<T, HALF>::addr(i)] : p;
            const unsigned higher = __ballot_sync(0xffffffffu, valid && v > p);
            const int stop = higher ? (__ffs(higher) - 1) : 32;
            if (lane < stop && v < floor_lane) floor_lane = v;
            if (higher) return true;
        }
    }
    return false;
}

// utils/get_peak_prominence.py:32-54 on the chunk summaries (cmax/cmin: this lane's chunk maximum / minimum)
template <typename T, int HALF>
__device__ T coop_prominence(const T *mags, int j, T cmax, T cmin, int lane) {
    constexpr int C = K3<T, HALF>::C;
    const T p = mags[K3<T, HALF>::addr(j)];
    const int cj = j / C;
    const unsigned above = __ballot_sync(0xffffffffu, cmax > p);
    T fl = p, fr = p;
    if (!scan_piece<T, HALF, -1>(mags, C * cj, j - 1, p, fl, lane)) {
        const unsigned hl = above & ((1u << cj) - 1u);
        const int L = hl ? 31 - __clz(hl) : -1;
        if (lane > L && lane < cj && cmin < fl) fl = cmin;
        if (L >= 0) scan_piece<T, HALF, -1>(mags, C * L, C * (L + 1) - 1, p, fl, lane);
    }
    if (!scan_piece<T, HALF, +1>(mags, j + 1, C * (cj + 1) - 1, p, fr, lane)) {
        const unsigned hr = above & ~((2u << cj) - 1u);
        const int R = hr ? __ffs(hr) - 1 : 32;
        if (lane > cj && lane < R && cmin < fr) fr = cmin;
        if (R < 32) scan_piece<T, HALF, +1>(mags, C * R, C * (R + 1) - 1, p, fr, lane);
    }
    fl = warp_min_nonneg(fl);
    fr = warp_min_nonneg(fr);
    return sub_rn(p, vmax(fl, fr));
}

template <typename T, int HALF, typename P = K3<T, HALF>>
__device__ __forceinline__ int half_power_bins_f(const T *mags, T prom, int j) {
    const T top = mags[P::addr(j)];
    const T level = add_rn(sub_rn(top, prom), mul_rn(prom, (T)0.707));
    int lo = j;
    while (lo > 0 && mags[P::addr(lo)] > level) {
        if (mags[P::addr(lo)] > top) break;
        --lo;
    }
    int hi = j;
    while (hi < HALF - 1 && mags[P::addr(hi)] > level) {
        if (mags[P::addr(hi)] > top) break;
        ++hi;
    }
    const int w = hi - lo;
    return w > 1 ? w : 1;
}

template <typename T, int HALF, typename P = K3<T, HALF>>
__device__ __forceinline__ int half_height_bins_f(const T *mags, int j) {
    const T level = mul_rn((T)0.707, mags[P::addr(j)]);
    int lo = j;
    while (lo > 0 && mags[P::addr(lo)] > level) --lo;
    int hi = j;
    while (hi < HALF && mags[P::addr(hi)] > level) ++hi;
    return hi - lo;
}

// width_half_magnitude with the whole warp (all lanes call it with the same j): 32 bins per side and step instead of a
// serial walk - same result as half_height_bins_f (utils/get_peak_resolution.py:30-44)
template <typename T, int HALF, typename P = K3<T, HALF>>
__device__ __forceinline__ int half_height_bins_warp(const T *mags, int j, int lane) {
    const T level = mul_rn((T)0.707, mags[P::addr(j)]);
    int lo = 0, hi = HALF;
    for (int base = j; base > 0; base -= 32) {  // left = first i <= j with mags[i] <= level, else 0 (bin 0 is never tested)
        const int i = base - lane;
        const unsigned stop = __ballot_sync(0xffffffffu, i <= 0 || !(mags[P::addr(i > 0 ? i : 0)] > level));
        if (stop) {
            lo = max(base - (__ffs(stop) - 1), 0);
            break;
        }
    }
    for (int base = j; base < HALF; base += 32) {  // right = first i >= j with i == HALF or mags[i] <= level
        const int i = base + lane;
        const unsigned stop = __ballot_sync(0xffffffffu, i >= HALF || !(mags[P::addr(i < HALF ? i : HALF - 1)] > level));
        if (stop) {
            hi = base + (__ffs(stop) - 1);
            break;
        }
    }
    return hi - lo;
}

// The reference sorts the gated candidates by round(mag, 4) descending (stable: ties keep ascending idx) and walks that
// order with the greedy "hump" exclusion.  Each lane owns PER slots; a slot's place in the order is its rank (number of
// passing slots that precede it), computed once with shuffles.  Accepted peaks go straight into the record.
template <typename T, int HALF, int PER, typename P = K3<T, HALF>>
__device__ __forceinline__ int order_and_exclude(const SlotT<T> *slots, int nslot, const T *mags, unsigned char *rec_s,
                                                 double df, int k, int lane) {
    // each lane owns PER slots; the order is extracted one element at a time with a warp arg-max (REDUX on the key's
    // words: largest round(mag, 4), ties -> lowest idx), at most k + rejected times
    double key[PER];
    int sidx[PER];
    int left = 0;
#pragma unroll
    for (int r = 0; r < PER; ++r) {
        const int e = lane + 32 * r;
        const bool ok = e < nslot && slots[e].width != 0;
        sidx[r] = ok ? (int)slots[e].idx : -1;
        key[r] = ok ? round_dec4_units((double)mags[P::addr(sidx[r])]) : -1.0;
        left += ok ? 1 : 0;
    }
    left = __reduce_add_sync(0xffffffffu, left);  // gated candidates not yet extracted: no arg-max round for an empty list
    int na = 0;
    while (na < k && left > 0) {
        --left;
        double bk = key[0];
        int bi = sidx[0], br = 0;
#pragma unroll
        for (int r = 1; r < PER; ++r) {
            if (sidx[r] >= 0 && (bi < 0 || key[r] > bk || (key[r] == bk && sidx[r] < bi))) {
                bk = key[r];
                bi = sidx[r];
                br = r;
            }
        }
        const int mine = bi;
        warp_argmax(bk, bi);
        if (bi < 0) break;
        const bool owner = mine == bi;  // bin indices are unique
        const int e_sel = __reduce_max_sync(0xffffffffu, owner ? lane + 32 * br : -1);
        if (owner) {
#pragma unroll
            for (int r = 0; r < PER; ++r)
                if (r == br) sidx[r] = -1;
        }
        const int c_idx = bi;
        const T cprom = slots[e_sel].prom;
        const T cmag = mags[P::addr(c_idx)];
        bool hump = false;
        for (int a = 0; a < na && !hump; ++a) {
            const int ja = reinterpret_cast<const int *>(rec_s + 8 + 24 * a)[0];
            const double fc = mul_rn((double)c_idx, df), fa = mul_rn((double)ja, df);
            // |round4(fc) - round4(fa)| >= |fc - fa| - 1e-4 and round4(fa) <= fa + 5e-5: most pairs are provably > 5 % apart
            if (fabs(fc - fa) - 1.0e-4 > 0.05 * (fa + 5.0e-5) * (1.0 + 1e-9)) continue;
            if (hump_exact(c_idx, ja, df, (double)cprom, bk)) hump = true;
        }
        if (!hump) {
            if (lane == 0) {
                unsigned char *pk = rec_s + 8 + 24 * na;
                reinterpret_cast<int *>(pk)[0] = c_idx;
                reinterpret_cast<int *>(pk)[1] = slots[e_sel].width;
                reinterpret_cast<double *>(pk + 8)[0] = (double)cmag;
                reinterpret_cast<double *>(pk + 8)[1] = (double)cprom;
            }
            ++na;
            __syncwarp();
        }
    }
    return na;
}

// Same order / exclusion for any number of slots (only reached with > 96 gated candidates, i.e. noise-like windows in
// the fused kernel, whose slot list lives in the free FFT buffer): extract the order one element at a time.
template <typename T, int HALF, typename P = K3<T, HALF>>
__device__ int order_and_exclude_any(const SlotT<T> *slots, int nslot, const T *mags, unsigned char *rec_s, double df,
                                     int k, int lane) {
    double prev_key = CUDART_INF;
    int prev_idx = -1, na = 0;
    while (na < k) {
        double best = -1.0;
        int best_idx = 0x7fffffff, best_e = -1;
        for (int e = lane; e < nslot; e += 32) {
            if (slots[e].width == 0) continue;
            const int ix = slots[e].idx;
            const double r = round_dec4_units((double)mags[P::addr(ix)]);
            const bool after_prev = r < prev_key || (r == prev_key && ix > prev_idx);
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
        if (best_e < 0) break;
        prev_key = best;
        prev_idx = best_idx;
        const T cprom = slots[best_e].prom;
        const T cmag = mags[P::addr(best_idx)];
        bool hump = false;
        for (int a = 0; a < na && !hump; ++a) {
            const int ja = reinterpret_cast<const int *>(rec_s + 8 + 24 * a)[0];
            const double cf = round_dec4_d(mul_rn((double)best_idx, df)), af = round_dec4_d(mul_rn((double)ja, df));
            if (div_rn(fabs(sub_rn(cf, af)), af) < 0.05 && div_rn((double)cprom, div_rn(best, 1e4)) < 0.10) hump = true;
        }
        if (!hump) {
            if (lane == 0) {
                unsigned char *pk = rec_s + 8 + 24 * na;
                reinterpret_cast<int *>(pk)[0] = best_idx;
                reinterpret_cast<int *>(pk)[1] = slots[best_e].width;
                reinterpret_cast<double *>(pk + 8)[0] = (double)cmag;
                reinterpret_cast<double *>(pk + 8)[1] = (double)cprom;
            }
            ++na;
            __syncwarp();
        }
    }
    return na;
}

// Damping / width gates of one flexible-picker candidate (utils/get_peak_prominence.py:177-186): returns the half-power
// width in bins if the candidate passes every gate, else 0.
template <typename T, int HALF, typename P>
__device__ __forceinline__ int k3_gate(const T *mags, int j, T prom, double half_sd, double df) {
    if (!((double)prom > half_sd)) return 0;
    const int bins = half_power_bins_f<T, HALF, P>(mags, prom, j);
    const double width_hz = mul_rn((double)bins, df);
    if (!(width_hz > 0.0)) return 0;
    const double fn = mul_rn((double)j, df);
    // 0.001 <= 1/(2*(fn/width_hz)) <= 0.07, decided by products unless within 1e-12 of a bound
    const double lo_b = 0.002 * fn, hi_b = 0.14 * fn;
    if (width_hz >= lo_b * (1.0 + 1e-12) && width_hz <= hi_b * (1.0 - 1e-12)) return bins;
    if (!(width_hz < lo_b * (1.0 - 1e-12) || width_hz > hi_b * (1.0 + 1e-12))) {
        const double q = div_rn(fn, width_hz);
        const double damping = div_rn(1.0, mul_rn(2.0, q));
        if (0.001 <= damping && damping <= 0.07) return bins;
    }
    return 0;
}

// Flexible picker for windows with many candidates (noise-like spectra: dozens of hot local maxima of which only the k
// largest can be reported).  A candidate's gates do not depend on the other candidates, and the reference walks the
// gated candidates in the order "descending round(mag, 4), ascending idx" until k are accepted: so the candidates are
// extracted in that order FIRST (warp arg-max, as in order_and_exclude) and prominence, width and damping are computed
// only for the ones reached.  Same records as evaluating every candidate (the general kernel does; tests compare).
template <typename T, int HALF, int PER, typename P = K3<T, HALF>>
__device__ __forceinline__ int order_eval_exclude_lazy(const SlotT<T> *slots, int nslot, const T *mags, unsigned char *rec_s,
                                                       T cmax, T cmin, double sd, double df, int k, int lane) {
    double key[PER];
    int sidx[PER];
#pragma unroll
    for (int r = 0; r < PER; ++r) {
        const int e = lane + 32 * r;
        const bool ok = e < nslot;
        sidx[r] = ok ? (int)slots[e].idx : -1;
        key[r] = ok ? round_dec4_units((double)mags[P::addr(sidx[r])]) : -1.0;
    }
    const double half_sd = mul_rn(0.5, sd);
    int na = 0;
    while (na < k) {
        double bk = key[0];
        int bi = sidx[0], br = 0;
#pragma unroll
        for (int r = 1; r < PER; ++r) {
            if (sidx[r] >= 0 && (bi < 0 || key[r] > bk || (key[r] == bk && sidx[r] < bi))) {
                bk = key[r];
                bi = sidx[r];
                br = r;
            }
        }
        const int mine = bi;
        warp_argmax(bk, bi);
        if (bi < 0) break;
        if (mine == bi) {  // bin indices are unique: one lane owns the extracted candidate
#pragma unroll
            for (int r = 0; r < PER; ++r)
                if (r == br) sidx[r] = -1;
        }
        const int c_idx = bi;
        const T cprom = coop_prominence<T, HALF>(mags, c_idx, cmax, cmin, lane);
        const int width = k3_gate<T, HALF, P>(mags, c_idx, cprom, half_sd, df);  // same value in every lane
        if (width == 0) continue;
        const T cmag = mags[P::addr(c_idx)];
        bool hump = false;
        for (int a = 0; a < na && !hump; ++a) {
            const int ja = reinterpret_cast<const int *>(rec_s + 8 + 24 * a)[0];
            const double fc = mul_rn((double)c_idx, df), fa = mul_rn((double)ja, df);
            if (fabs(fc - fa) - 1.0e-4 > 0.05 * (fa + 5.0e-5) * (1.0 + 1e-9)) continue;
            if (hump_exact(c_idx, ja, df, (double)cprom, bk)) hump = true;
        }
        if (!hump) {
            if (lane == 0) {
                unsigned char *pk = rec_s + 8 + 24 * na;
                reinterpret_cast<int *>(pk)[0] = c_idx;
                reinterpret_cast<int *>(pk)[1] = width;
                reinterpret_cast<double *>(pk + 8)[0] = (double)cmag;
                reinterpret_cast<double *>(pk + 8)[1] = (double)cprom;
            }
            ++na;
            __syncwarp();
        }
    }
    return na;
}

// Rigid picker on the hot list (utils/get_peak_resolution.py:94-126), one warp; returns the number of accepted peaks and
// writes them into the record under construction.  Zeroes magnitudes in shared memory as the reference does.
__device__ __forceinline__ int hot_bin(const uint16_t *list, int e) { return list[e]; }
template <typename T>
__device__ __forceinline__ int hot_bin(const SlotT<T> *list, int e) { return list[e].idx; }
// `hot`: bare 16-bit bin indices (pipeline kernels) or slot structs (fused kernel)
template <typename T, int HALF, typename P, typename List>
__device__ __forceinline__ int k3_rigid(T *mags, const List *hot, int nslot, T thr_f, double df, int k, int lane,
                                        unsigned char *rec_s) {
    const double distance = sub_rn(mul_rn(2.0, df), mul_rn(1.0, df));
    const bool df_plain = df > 1e-300 && df < 1e300;
    int na = 0;
    __syncwarp();
    while (na < k) {
        T bm = (T)-1;
        int bj = -1;
        for (int e = lane; e < nslot; e += 32) {
            const int j = hot_bin(hot, e);
            const T m = mags[P::addr(j)];
            if (j >= 1 && j <= HALF - 2 && m > thr_f && (m > bm || (m == bm && j < bj)) && m > mags[P::addr(j - 1)] &&
                m > mags[P::addr(j + 1)]) {
                bm = m;
                bj = j;
            }
        }
        warp_argmax(bm, bj);
        if (bj < 0) break;
        const int w2 = half_height_bins_warp<T, HALF, P>(mags, bj, lane);
        // resolution against every accepted peak, one accepted peak per lane (their bins are in the record under
        // construction; the barrier that ends every round orders lane 0's stores before these reads)
        bool clash = false;
        if (lane < na) {
            const int ja = reinterpret_cast<const int *>(rec_s + 8 + 24 * lane)[0];
            // an accepted peak's own bin was zeroed when it was found, so its half-height width is 0 (the
            // reference's resolution() degenerates to 1.18*dist/w_candidate); walk only if that ever fails
            const int w1 = mags[P::addr(ja)] == (T)0 ? 0 : half_height_bins_f<T, HALF, P>(mags, ja);
            bool ok = false;
            if (w1 + w2 != 0) {
                const double num = mul_rn(1.18, (double)abs(bj - ja)), den = 1.5 * (double)(w1 + w2);
                if (num >= den * (1.0 + 1e-12)) ok = true;                       // rs >= 1.5 by a clear margin
                else if (!(num < den * (1.0 - 1e-12))) ok = div_rn(num, (double)(w1 + w2)) >= 1.5;
            }
            clash = !ok;
        }
        const bool separated = !__any_sync(0xffffffffu, clash);
        if (separated) {
            if (lane == 0) {
                unsigned char *pk = rec_s + 8 + 24 * na;
                reinterpret_cast<int *>(pk)[0] = bj;
                reinterpret_cast<int *>(pk)[1] = w2;
                reinterpret_cast<double *>(pk + 8)[0] = (double)bm;
            }
            ++na;
        }
        // zeroing radius round((freq * 0.02) / (frequencies[2] - frequencies[1])): equals round-half-even(idx / 50),
        // which is floor(idx / 50) + (idx mod 50 > 25) on integers, unless idx / 50 is an exact tie; only then (or for a
        // degenerate df) the exact expression runs
        int reach;
        {
            const int q50 = bj / 50, r50 = bj - 50 * q50;
            if (df_plain && r50 != 25) {
                reach = q50 + (r50 > 25 ? 1 : 0);
            } else {
                const double f = mul_rn((double)bj, df);
                double reach_d = rint(div_rn(mul_rn(f, 0.02), distance));
                if (!(reach_d >= 0.0)) reach_d = 0.0;
                if (reach_d > (double)HALF) reach_d = (double)HALF;
                reach = (int)reach_d;
            }
        }
        const int z0 = max(0, bj - reach), z1 = min(HALF, bj + reach + 1);
        __syncwarp();
        for (int b = z0 + lane; b < z1; b += 32) mags[P::addr(b)] = (T)0;
        __syncwarp();
    }
    return na;
}

// fp32 tie test of the four bins jb..jb+3 (see APDA_STATUS_FP32_TIE): a hot bin equal to its right neighbour and higher
// than both outer neighbours.  Out of line: reached only when the quad holds two equal adjacent magnitudes.
template <typename T, int HALF, typename P = K3<T, HALF>>
static __device__ __noinline__ bool plateau_top_in_quad(const T *mags, int jb, T thr_f) {
    bool tie = false;
    for (int j = jb; j < jb + 4; ++j) {
        const T e = mags[P::addr(j)];
        if (e > thr_f && j >= 1 && j + 1 <= HALF - 1 && e == mags[P::addr(j + 1)] && e > mags[P::addr(j - 1)] &&
            (j + 2 > HALF - 1 || e > mags[P::addr(j + 2)]))
            tie = true;
    }
    return tie;
}

// Everything after the magnitudes are in shared memory: hot-bin list, picker, record.  Shared by the pipeline kernels
// (peaks_f32_fast.cu, peaks_f64_fast.cu).  Runs on ONE warp.
// thr_f: largest T with  m > thr  <=>  m > thr_f  for every magnitude m (the threshold itself when T is double).
template <typename T, int HALF, bool FLEX>
__device__ __forceinline__ void k3_tail(T *mags, SlotT<T> *slots, const int slot_cap, unsigned char *rec_s,
                                        int *nslot_ptr, const double sd,
                                        const T thr_f, const double df, const int k, const int lane,
                                        const int64_t win, unsigned char *__restrict__ recs, int *__restrict__ repair) {
    using P = K3<T, HALF>;
    constexpr int C = P::C;
    // ---- phase 2: contiguous chunk per lane: chunk max/min, hot bins -> slot list -------------------------------------
    T cmax = -(T)CUDART_INF_F, cmin = (T)CUDART_INF_F;
    bool tie_lane = false;
    {
        const T *ch = mags + P::addr(C * lane);
        unsigned hotq = 0;  // bit q: the q-th group of 4 bins of this chunk holds a bin above the threshold (C/4 <= 32 groups)
        bool &tie = tie_lane;  // fp32 only: a hot two-bin plateau top (see APDA_STATUS_FP32_TIE)
#pragma unroll
        for (int q = 0; q < C / 4; ++q) {
            const Quad<T> v = lds_quad(ch + 4 * q);
            const T m4 = vmax(vmax(v.x, v.y), vmax(v.z, v.w));
            cmax = vmax(cmax, m4);
            cmin = vmin(cmin, vmin(vmin(v.x, v.y), vmin(v.z, v.w)));
            if (m4 > thr_f) hotq |= 1u << q;
        }
        while (hotq) {  // rare: a handful of quads per window; every test of the quad on registers, one counter update
            const int q = __ffs(hotq) - 1;
            hotq &= hotq - 1;
            const int jb = C * lane + 4 * q;  // first bin of the quad
            const Quad<T> v = lds_quad(ch + 4 * q);
            const T left = mags[P::addr(jb > 0 ? jb - 1 : 0)];                   // jb == 0: v.x itself, v.x > left fails
            const T right = mags[P::addr(jb + 4 < HALF ? jb + 4 : HALF - 1)];    // last quad: v.w itself
            unsigned take;
            if (FLEX)  // strict local maxima above the threshold, candidates j in [1, HALF-2]
                take = (v.x > thr_f && jb >= 1 && v.x > left && v.x > v.y ? 1u : 0u) |
                       (v.y > thr_f && v.y > v.x && v.y > v.z ? 2u : 0u) |
                       (v.z > thr_f && v.z > v.y && v.z > v.w ? 4u : 0u) |
                       (v.w > thr_f && jb + 3 <= HALF - 2 && v.w > v.z && v.w > right ? 8u : 0u);
            else
                take = (v.x > thr_f ? 1u : 0u) | (v.y > thr_f ? 2u : 0u) | (v.z > thr_f ? 4u : 0u) | (v.w > thr_f ? 8u : 0u);
            if (take) {
                int pos = atomicAdd(&(*nslot_ptr), __popc(take));
                do {
                    const int j = jb + __ffs(take) - 1;
                    take &= take - 1;
                    if (FLEX) {
                        if (pos < slot_cap) slots[pos].idx = (uint16_t)j;
                    } else {
                        if (pos < P::HOT_CAP) reinterpret_cast<uint16_t *>(slots)[pos] = (uint16_t)j;
                    }
                    ++pos;
                } while (take);
            }
            // Two adjacent bins that are EQUAL in fp32 and higher than both outer neighbours: neither is a strict
            // local maximum, so no peak is reported there, while the fp64 reference (whose magnitudes differ in
            // the bits fp32 drops) reports one of them.  The window is flagged so the caller can re-run it in fp64.
            if (sizeof(T) == 4 && (v.x == v.y || v.y == v.z || v.z == v.w || v.w == right))
                tie = tie || plateau_top_in_quad<T, HALF>(mags, jb, thr_f);
        }
    }
    __syncwarp();
    const int nslot_raw = (*nslot_ptr);
    if (nslot_raw > (FLEX ? slot_cap : P::HOT_CAP)) {  // more than the on-chip list holds: hand the window to the general kernel
        if (lane == 0) repair[1 + atomicAdd(&repair[0], 1)] = (int)win;
        return;
    }
    const int nslot = nslot_raw;
    const int status = (sizeof(T) == 4 && __any_sync(0xffffffffu, tie_lane)) ? APDA_STATUS_FP32_TIE : 0;

    int na = 0;
    if (FLEX && nslot > APDA_K3_LAZY_MIN) {
        // noise-like window (dozens of hot local maxima): the candidates are visited in the reference's output order and
        // evaluated only until k of them are accepted
        na = P::SLOTS > 96 ? order_eval_exclude_lazy<T, HALF, 5>(slots, nslot, mags, rec_s, cmax, cmin, sd, df, k, lane)
                           : order_eval_exclude_lazy<T, HALF, 3>(slots, nslot, mags, rec_s, cmax, cmin, sd, df, k, lane);
    } else if (FLEX) {
        // ---- A: cooperative prominence per candidate -----------------------------------------------------------------
        for (int c = 0; c < nslot; ++c) {
            const int j = slots[c].idx;
            const T prom = coop_prominence<T, HALF>(mags, j, cmax, cmin, lane);
            if (lane == 0) slots[c].prom = prom;
        }
        __syncwarp();
        // ---- B: lane-parallel gates (one candidate per lane) ---------------------------------------------------------
        const double half_sd = mul_rn(0.5, sd);
        for (int c = lane; c < nslot; c += 32)
            slots[c].width = (uint16_t)k3_gate<T, HALF, P>(mags, slots[c].idx, slots[c].prom, half_sd, df);
        __syncwarp();
        // ---- C: order "descending round(mag,4), ascending idx" (stable sort of the reference), greedy hump exclusion ------
        na = order_and_exclude<T, HALF, 1>(slots, nslot, mags, rec_s, df, k, lane);
    } else {
        na = k3_rigid<T, HALF, P>(mags, reinterpret_cast<const uint16_t *>(slots), nslot, thr_f, df, k, lane, rec_s);
    }
    if (lane == 0) {
        reinterpret_cast<int *>(rec_s)[0] = na;
        reinterpret_cast<int *>(rec_s)[1] = status;
    }
    __syncwarp();
    if (lane < 16)  // one coalesced 128-byte store (local HBM, or the fleet table in a peer's HBM over NVLink)
        reinterpret_cast<uint64_t *>(recs + win * 128)[lane] = reinterpret_cast<const uint64_t *>(rec_s)[lane];
}

// ---- the same picker with the WHOLE CTA of the fused kernel (T threads = T/32 warps per window) ---------------------------
// In the fused window->record kernel the CTA's FFT warps would idle while one warp runs the tail (and the CTA keeps its
// registers and shared memory meanwhile); here every thread owns a chunk of 16 bins in phase 2, the candidates'
// prominence walks are dealt out to the warps (chunk summaries in shared memory: whole chunks are skipped with ballots
// over the 32 * W chunk maxima, only boundary chunks are scanned), the gates run thread-parallel, and only the short
// ordering / record epilogue is left to warp 0.  Same decisions as k3_tail (tests compare the records).  (Measured and
// rejected for the PIPELINE picker, where one warp per window hides the tail's latency behind other windows: a CTA per
// window with this tail ran 4.5 instead of 3.3 ns per window.)
template <int NT>
__device__ __forceinline__ void window_sync() {  // all NT threads that share the window (the whole CTA of the fused kernel)
    if (NT == 32) __syncwarp();
    else asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
}

template <int HALF>
struct K3M {  // magnitude layout of the multi-warp tail: chunks of 16 bins, 4-word pad per chunk (conflict-free LDS.128)
    static constexpr int C = 16;
    static constexpr int NCH = HALF / C;
    static constexpr int MAGW = HALF + 4 * NCH;
    __device__ static __forceinline__ int addr(int b) { return b + ((b >> 4) << 2); }
};

// prominence of bin j (utils/get_peak_prominence.py:32-54) on chunk summaries held in shared memory; one warp
template <int HALF>
__device__ float prominence_mw(const float *mags, const float *cmaxs, const float *cmins, int j, int lane) {
    using P = K3M<HALF>;
    constexpr int NW = P::NCH / 32;  // 32-chunk words
    const float p = mags[P::addr(j)];
    const int cj = j >> 4, jo = j & 15;
    // own chunk: lanes 0..15 hold its bins
    const float own = lane < 16 ? mags[P::addr(16 * cj + lane)] : p;
    const unsigned hi_own = __ballot_sync(0xffffffffu, lane < 16 && own > p);
    const unsigned left_hi = hi_own & ((1u << jo) - 1u);          // higher bins left of j in its chunk
    const unsigned right_hi = hi_own & ~((2u << jo) - 1u) & 0xffffu;  // ... right of j
    const int lstop = left_hi ? 31 - __clz(left_hi) : -1;         // walk covers (lstop, jo)
    const int rstop = right_hi ? __ffs(right_hi) - 1 : 16;        // walk covers (jo, rstop)
    float fl = p, fr = p;
    if (lane < 16 && lane > lstop && lane < jo) fl = own;
    if (lane < 16 && lane > jo && lane < rstop) fr = own;
    // chunk maxima above p, NW words of 32 chunks
    unsigned above[NW];
#pragma unroll
    for (int w = 0; w < NW; ++w) above[w] = __ballot_sync(0xffffffffu, cmaxs[32 * w + lane] > p);
    if (!left_hi) {  // the walk leaves the chunk to the left: nearest chunk L < cj with a higher bin, else down to bin 0
        int L = -1;
#pragma unroll
        for (int w = NW - 1; w >= 0; --w) {
            if (L < 0 && 32 * w < cj) {
                const unsigned m = 32 * w + 32 <= cj ? above[w] : (above[w] & ((1u << (cj - 32 * w)) - 1u));
                if (m) L = 32 * w + 31 - __clz(m);
            }
        }
#pragma unroll
        for (int w = 0; w < NW; ++w) {  // chunks strictly between L and cj are walked completely
            const int id = 32 * w + lane;
            if (id > L && id < cj) fl = fminf(fl, cmins[id]);
        }
        if (L >= 0) {  // inside L: bins right of its last higher bin
            const float v = lane < 16 ? mags[P::addr(16 * L + lane)] : p;
            const unsigned h = __ballot_sync(0xffffffffu, lane < 16 && v > p);
            const int stop = 31 - __clz(h);  // h != 0: the chunk maximum is higher than p
            if (lane < 16 && lane > stop) fl = fminf(fl, v);
        }
    }
    if (!right_hi) {
        int R = P::NCH;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            if (R == P::NCH && 32 * w + 31 > cj) {
                const unsigned m = 32 * w > cj ? above[w] : (above[w] & ~((2u << (cj - 32 * w)) - 1u));
                if (m) R = 32 * w + __ffs(m) - 1;
            }
        }
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            const int id = 32 * w + lane;
            if (id > cj && id < R) fr = fminf(fr, cmins[id]);
        }
        if (R < P::NCH) {
            const float v = lane < 16 ? mags[P::addr(16 * R + lane)] : p;
            const unsigned h = __ballot_sync(0xffffffffu, lane < 16 && v > p);
            const int stop = __ffs(h) - 1;
            if (lane < 16 && lane < stop) fr = fminf(fr, v);
        }
    }
    fl = warp_min_nonneg(fl);
    fr = warp_min_nonneg(fr);
    return sub_rn(p, fmaxf(fl, fr));
}

// scratch: cmaxs[NCH], cmins[NCH] floats, then the slot list (slot_cap entries), then two ints (slot count, tie flag);
// all threads of the window's CTA call this after the magnitudes are in shared memory (and a barrier)
template <int HALF, bool FLEX, int NT>
__device__ __forceinline__ void k3_tail_mw(float *mags, float *cmaxs, float *cmins, Slot *slots, const int slot_cap,
                                           int *ctl /* [0] slots, [1] tie */, unsigned char *rec_s, const double sd,
                                           const float thr_f, const double df, const int k, const int t, const int64_t win,
                                           unsigned char *__restrict__ recs) {
    using P = K3M<HALF>;
    static_assert(P::NCH == NT, "one 16-bin chunk per thread");
    const int lane = t & 31, warp = t >> 5;
    constexpr int W = NT / 32;
    // ---- phase 2: one 16-bin chunk per thread --------------------------------------------------------------------------
    {
        const float *ch = mags + P::addr(16 * t);
        float cmax = -CUDART_INF_F, cmin = CUDART_INF_F;
        unsigned hotq = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 v = *reinterpret_cast<const float4 *>(ch + 4 * q);
            const float m4 = fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w));
            cmax = fmaxf(cmax, m4);
            cmin = fminf(cmin, fminf(fminf(v.x, v.y), fminf(v.z, v.w)));
            if (m4 > thr_f) hotq |= 1u << q;
        }
        cmaxs[t] = cmax;
        cmins[t] = cmin;
        while (hotq) {  // every test of the quad on registers, one counter update per quad (see k3_tail)
            const int q = __ffs(hotq) - 1;
            hotq &= hotq - 1;
            const int jb = 16 * t + 4 * q;
            const float4 v = *reinterpret_cast<const float4 *>(ch + 4 * q);
            const float left = mags[P::addr(jb > 0 ? jb - 1 : 0)];
            const float right = mags[P::addr(jb + 4 < HALF ? jb + 4 : HALF - 1)];
            unsigned take;
            if (FLEX)
                take = (v.x > thr_f && jb >= 1 && v.x > left && v.x > v.y ? 1u : 0u) |
                       (v.y > thr_f && v.y > v.x && v.y > v.z ? 2u : 0u) |
                       (v.z > thr_f && v.z > v.y && v.z > v.w ? 4u : 0u) |
                       (v.w > thr_f && jb + 3 <= HALF - 2 && v.w > v.z && v.w > right ? 8u : 0u);
            else
                take = (v.x > thr_f ? 1u : 0u) | (v.y > thr_f ? 2u : 0u) | (v.z > thr_f ? 4u : 0u) | (v.w > thr_f ? 8u : 0u);
            if (take) {
                int pos = atomicAdd(&ctl[0], __popc(take));
                do {
                    const int j = jb + __ffs(take) - 1;
                    take &= take - 1;
                    if (pos < slot_cap) slots[pos].idx = (uint16_t)j;
                    ++pos;
                } while (take);
            }
            if ((v.x == v.y || v.y == v.z || v.z == v.w || v.w == right) &&
                plateau_top_in_quad<float, HALF, P>(mags, jb, thr_f))
                ctl[1] = 1;  // APDA_STATUS_FP32_TIE, see k3_tail
        }
    }
    window_sync<NT>();
    const int nslot = min(ctl[0], slot_cap);
    const int status = (ctl[0] > slot_cap ? APDA_STATUS_TRUNCATED : 0) | (ctl[1] ? APDA_STATUS_FP32_TIE : 0);
    int na = 0;
    if (FLEX) {
        for (int c = warp; c < nslot; c += W) {  // A: the candidates' prominence walks, dealt out to the warps
            const float prom = prominence_mw<HALF>(mags, cmaxs, cmins, slots[c].idx, lane);
            if (lane == 0) slots[c].prom = prom;
        }
        window_sync<NT>();
        const double half_sd = mul_rn(0.5, sd);
        for (int c = t; c < nslot; c += NT)  // B: gates, one candidate per thread
            slots[c].width = (uint16_t)k3_gate<float, HALF, P>(mags, slots[c].idx, slots[c].prom, half_sd, df);
        window_sync<NT>();
        if (warp != 0) return;
        na = nslot <= 32   ? order_and_exclude<float, HALF, 1, P>(slots, nslot, mags, rec_s, df, k, lane)
             : nslot <= 96 ? order_and_exclude<float, HALF, 3, P>(slots, nslot, mags, rec_s, df, k, lane)
                           : order_and_exclude_any<float, HALF, P>(slots, nslot, mags, rec_s, df, k, lane);
    } else {
        if (warp != 0) return;
        na = k3_rigid<float, HALF, P>(mags, slots, nslot, thr_f, df, k, lane, rec_s);
    }
    if (lane == 0) {
        reinterpret_cast<int *>(rec_s)[0] = na;
        reinterpret_cast<int *>(rec_s)[1] = status;
    }
    __syncwarp();
    if (lane < 16) reinterpret_cast<uint64_t *>(recs + win * 128)[lane] = reinterpret_cast<const uint64_t *>(rec_s)[lane];
}

}  // namespace
