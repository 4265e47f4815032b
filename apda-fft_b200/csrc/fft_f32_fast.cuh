// K1 (fp32 fast form) for N = 1024 / 2048 / 4096 / 8192: the headline kernel.
//
// A real window of N samples is transformed as an M = N/2 point complex FFT of z[n] = x[2n] + i*x[2n+1] followed by
// the split ("untangle") step  X[k] = E[k] + W_N^k O[k],  X[k+M] = E[k] - W_N^k O[k].  M = R1*R2*R3 (16*16*8 for
// N = 4096) is evaluated in three register-resident radix passes, 16 complex values per thread:
//
//   pass 1  radix R1 over the slowest input digit (coalesced 64-bit HBM loads, centring fused into the load),
//           inter-pass twiddles from an L1-resident table laid out [k1][column] (coalesced),
//   pass 2  radix R2, in place in shared memory, twiddles W_{R2*R3}^{n3*k2} from __constant__ memory,
//   pass 3  radix R3, results rewritten to shared memory in natural bin order,
//   split   two bins per thread-step: 128-bit conflict-free LDS of Z[k], Z[k+1], partner bins Z[M-k], twiddle
//           -i/2 * W_N^k from a table, two 128-bit coalesced HBM stores (bins k..k+1 and M+k..M+k+1); bin 0 := 0.
//
// Only 2 shared-memory exchanges + the natural-order staging separate the HBM load from the HBM store; the
// radix-16 butterflies keep their W_16 constants in registers.  Bit reversal never materialises: the digit
// permutation is folded into the register/shared-memory index maps.
//
// Semantics reproduced (reference metrics/fft_iterativa.py:74-87): centre (exact median by default, see
// select_median below), zero pad n_samples -> N AFTER centring, forward unscaled DFT, bin 0 forced to zero.
// fp32 results agree with the reference to ~1e-6 of the window maximum (tests: 1e-5 on magnitudes/prominences,
// exact peak indices).
#pragma once
#include "common.cuh"

namespace {

// Blackwell packed fp32 arithmetic (FADD2 / FMUL2 / FFMA2): one instruction per (re, im) pair.  The scalar 3-operand
// FADD/FFMA issue at half rate on sm_100, so the butterflies' complex adds are written in packed form.
__device__ __forceinline__ unsigned long long &u64(float2 &v) { return reinterpret_cast<unsigned long long &>(v); }
__device__ __forceinline__ const unsigned long long &u64(const float2 &v) {
    return reinterpret_cast<const unsigned long long &>(v);
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) {
    float2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(u64(r)) : "l"(u64(a)), "l"(u64(b)));
    return r;
}
__device__ __forceinline__ float2 csub(float2 a, float2 b) {
    float2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(u64(r)) : "l"(u64(a)), "l"(u64(b)));
    return r;
}
__device__ __forceinline__ float2 pmul(float2 a, float2 b) {
    float2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(u64(r)) : "l"(u64(a)), "l"(u64(b)));
    return r;
}
__device__ __forceinline__ float2 pfma(float2 a, float2 b, float2 c) {
    float2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(u64(r)) : "l"(u64(a)), "l"(u64(b)), "l"(u64(c)));
    return r;
}
__device__ __forceinline__ float2 cmul(float2 a, float2 w) {
    return make_float2(fmaf(a.x, w.x, -a.y * w.y), fmaf(a.x, w.y, a.y * w.x));
}

#define SQRT1_2f 0.70710678118654752440f
#define COS_PI_8f 0.92387953251128675613f
#define SIN_PI_8f 0.38268343236508977173f

// u + (-i)*d and u - (-i)*d for d = (d.x, d.y):  (-i)*d = (d.y, -d.x)
__device__ __forceinline__ void add_sub_mi(float2 u, float2 d, float2 &plus, float2 &minus) {
    plus = make_float2(u.x + d.y, u.y - d.x);
    minus = make_float2(u.x - d.y, u.y + d.x);
}

// forward DFTs on register arrays, natural order in and out
__device__ __forceinline__ void fft4(float2 &a0, float2 &a1, float2 &a2, float2 &a3) {
    const float2 t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), d = csub(a1, a3);
    a0 = cadd(t0, t2);
    a2 = csub(t0, t2);
    add_sub_mi(t1, d, a1, a3);
}
__device__ __forceinline__ void fft8(float2 *v) {  // v[0..7]
    float2 e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6];
    float2 o0 = v[1], o1 = v[3], o2 = v[5], o3 = v[7];
    fft4(e0, e1, e2, e3);
    fft4(o0, o1, o2, o3);
    // W8^1 = (1 - i)/sqrt2, W8^2 = -i, W8^3 = (-1 - i)/sqrt2
    o1 = pmul(make_float2(o1.x + o1.y, o1.y - o1.x), make_float2(SQRT1_2f, SQRT1_2f));
    o3 = pmul(make_float2(o3.y - o3.x, o3.x + o3.y), make_float2(SQRT1_2f, -SQRT1_2f));
    v[0] = cadd(e0, o0);
    v[4] = csub(e0, o0);
    v[1] = cadd(e1, o1);
    v[5] = csub(e1, o1);
    add_sub_mi(e2, o2, v[2], v[6]);
    v[3] = cadd(e3, o3);
    v[7] = csub(e3, o3);
}
__device__ __forceinline__ void fft16(float2 *v) {  // v[0..15]
    float2 e[8], o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        e[i] = v[2 * i];
        o[i] = v[2 * i + 1];
    }
    fft8(e);
    fft8(o);
    // W16^k, k = 1..7
    const float2 w1 = make_float2(COS_PI_8f, -SIN_PI_8f), w3 = make_float2(SIN_PI_8f, -COS_PI_8f);
    const float2 w5 = make_float2(-SIN_PI_8f, -COS_PI_8f), w7 = make_float2(-COS_PI_8f, -SIN_PI_8f);
    o[1] = cmul(o[1], w1);
    o[2] = pmul(make_float2(o[2].x + o[2].y, o[2].y - o[2].x), make_float2(SQRT1_2f, SQRT1_2f));
    o[3] = cmul(o[3], w3);
    o[5] = cmul(o[5], w5);
    o[6] = pmul(make_float2(o[6].y - o[6].x, o[6].x + o[6].y), make_float2(SQRT1_2f, -SQRT1_2f));
    o[7] = cmul(o[7], w7);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        if (i == 4) {
            add_sub_mi(e[4], o[4], v[4], v[12]);  // W16^4 = -i
        } else {
            v[i] = cadd(e[i], o[i]);
            v[i + 8] = csub(e[i], o[i]);
        }
    }
}
template <int R>
__device__ __forceinline__ void fft_r(float2 *v) {
    if (R == 16) fft16(v);
    else if (R == 8) fft8(v);
    else fft4(v[0], v[1], v[2], v[3]);
}

// pass-2 twiddles W_{R2*R3}^{n3*k2}, laid out [k2][n3]; one array per supported N (constant memory is per module)
__constant__ float2 c_tw2_1024[64];
__constant__ float2 c_tw2_2048[64];
__constant__ float2 c_tw2_4096[128];
__constant__ float2 c_tw2_8192[256];
template <int N>
__device__ __forceinline__ const float2 *const_tw2() {
    return N == 1024 ? c_tw2_1024 : N == 2048 ? c_tw2_2048 : N == 4096 ? c_tw2_4096 : c_tw2_8192;
}

// Resident CTAs per SM the pipeline K1 is compiled for at N = 4096.  One window per CTA and a serial chain of phases
// (load, median rounds, three passes, split) make the kernel's rate "resident CTAs / per-window latency": 9 CTAs
// (56 registers, no spills) measured 7-9 % faster than the 8 that 60 registers allow, 10 (48 registers, spills) slower;
// a shared-memory carve-out that leaves less L1 for the twiddle tables costs 15 %.
// Same reasoning for the other sizes (same-box A/B, exact-median K1): N = 8192 with 4 CTAs (64 registers) instead of the 3
// that 84 registers allow: 5.32 -> 3.91 ms per 200k windows (mean centring: 6.0-6.2 TB/s, 92-94 % of the copy
// bandwidth); N = 2048 / 1024 with 8 CTAs of 128 threads (64 registers instead of 88 / 84): -12 % / -8 %.
#ifndef APDA_K1_MINB_2048
#define APDA_K1_MINB_2048 8
#endif
#ifndef APDA_K1_MINB_1024
#define APDA_K1_MINB_1024 8
#endif
#ifndef APDA_K1_MINB_8192
#define APDA_K1_MINB_8192 4
#endif
#ifndef APDA_K1_MINB
#define APDA_K1_MINB 9
#endif
template <int N>
struct Plan;
template <>
struct Plan<8192> { static constexpr int R1 = 16, R2 = 16, R3 = 16, WPB = 1, MINB = APDA_K1_MINB_8192; };
template <>
struct Plan<4096> { static constexpr int R1 = 16, R2 = 16, R3 = 8, WPB = 1, MINB = APDA_K1_MINB; };
template <>
struct Plan<2048> { static constexpr int R1 = 16, R2 = 8, R3 = 8, WPB = 2, MINB = APDA_K1_MINB_2048; };
template <>
struct Plan<1024> { static constexpr int R1 = 8, R2 = 8, R3 = 8, WPB = 4, MINB = APDA_K1_MINB_1024; };

// ---- per-window barrier: windows that share a block do not run in lockstep ------------------------------------------
template <int T>
__device__ __forceinline__ void group_sync(int slot) {
    // literal barrier ids (at most two windows per block use barriers): a register id would reserve all 16 per CTA
    if (T == 32) __syncwarp();
    else if (slot == 0) asm volatile("bar.sync 1, %0;" ::"n"(T) : "memory");
    else asm volatile("bar.sync 2, %0;" ::"n"(T) : "memory");
}

__device__ __forceinline__ float next_above(float v) {  // smallest float strictly greater than v (finite v)
    float n = key_value(ordered_key(v) + 1u, 0.f);
    if (!(n > v)) n = key_value(ordered_key(v) + 2u, 0.f);  // -0.0 -> +0.0 compares equal
    return n;
}

// ---- exact median of the window, values held in registers ---------------------------------------------------------
// Counting selection: every round counts the values below a pivot (2 instructions per value) and moves one end of the
// bracket [lo, hi) that holds the two middle order statistics.  The first pivot is the window mean, the second a
// Newton step sized by the standard deviation, later ones interpolate inside the bracket (every third round bisects
// the key space, which bounds the worst case).  When at most 32 values remain in the bracket one warp ranks them.
// Exact for any input; the estimates only steer the pivots.  Padding slots (index >= n_valid) must hold +inf.
// Returns statistics.median: the middle value (odd n) or the mean of the two middle values (even n).
// smallest / largest of the warp's values whose bit is set in `inb` (bit 31 - q: value q); out of line on purpose: it
// serves the rare windows of quantised samples (see select_median) and must not lengthen everybody's instruction stream
__device__ __noinline__ float2 bracket_minmax(float2 a0, float2 a1, float2 a2, float2 a3, float2 a4, float2 a5, float2 a6,
                                              float2 a7, float2 a8, float2 a9, float2 a10, float2 a11, float2 a12, float2 a13,
                                              float2 a14, float2 a15, unsigned inb) {
    const float2 a[16] = {a0, a1, a2, a3, a4, a5, a6, a7, a8, a9, a10, a11, a12, a13, a14, a15};
    float vmin = CUDART_INF_F, vmax = -CUDART_INF_F;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        if (inb & (0x80000000u >> (2 * i))) {
            vmin = fminf(vmin, a[i].x);
            vmax = fmaxf(vmax, a[i].x);
        }
        if (inb & (0x80000000u >> (2 * i + 1))) {
            vmin = fminf(vmin, a[i].y);
            vmax = fmaxf(vmax, a[i].y);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
        vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
    }
    return make_float2(vmin, vmax);
}

constexpr int kTightenFrom = 4;  // counting rounds before the bracket is cut to the values inside it
template <int T, bool FULL, typename Reload>
__device__ float select_median(const float2 (&v2)[16], int n_valid, uint32_t *sh /* 64 words */, int t, int slot,
                               Reload reload /* reload(q): value q (0..31) of this thread, re-read from memory */) {
#define VAL(i) (((i) & 1) ? v2[(i) >> 1].y : v2[(i) >> 1].x)
    constexpr int NW = (T + 31) / 32;
    const int lane = t & 31, warp = t >> 5;
    const int r_lo = (n_valid - 1) >> 1, r_hi = n_valid >> 1;  // 0-based ranks of the two middle order statistics
    float *shf = reinterpret_cast<float *>(sh);

    // Shared-memory words: [0, 32) the survivors of the final gather (first the per-warp partial sums of the prologue,
    // or the even-n temporaries: never live at the same time), [32, 48) two banks of per-warp round counts, 48 the gather
    // counter, 49 the result.
    //
    // Mean and standard deviation only steer the first two pivots: every thread contributes 8 of its 32 values (slots
    // 0, 4, 8, 12: samples spread over the whole window), every warp reduces its own share, and all threads combine the
    // NW partial sums after ONE barrier - no warp idles while another one prepares the estimate.
    {
        float s1 = 0.f, s2 = 0.f;
        int cntv = 0;
#pragma unroll
        for (int c = 0; c < 16; c += 4) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const float raw = h ? v2[c].y : v2[c].x;
                const bool ok = FULL || raw < CUDART_INF_F;
                const float x = ok ? raw : 0.f;
                s1 += x;
                s2 = fmaf(x, x, s2);
                cntv += ok;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        }
        if (!FULL) cntv = __reduce_add_sync(0xffffffffu, cntv);
        if (lane == 0) {
            shf[warp] = s1;
            shf[8 + warp] = s2;
            if (!FULL) sh[16 + warp] = (uint32_t)cntv;
        }
        if (t == 0) sh[48] = 0;  // gather counter, used after the last round
    }
    group_sync<T>(slot);
    float mean, sd;
    {
        float s1 = 0.f, s2 = 0.f;
        int cntv = FULL ? 8 * T : 0;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            s1 += shf[w];
            s2 += shf[8 + w];
            if (!FULL) cntv += (int)sh[16 + w];
        }
        // steering values only (the counts decide): approximate reciprocal / root
        const float inv = rcp_approx((float)max(cntv, 1));
        mean = s1 * inv;
        sd = sqrt_approx(fmaxf(fmaf(-mean, mean, s2 * inv), 0.f));
    }
    const float inv_density = fmaxf(2.5f * sd, 1e-30f) * rcp_approx((float)n_valid);  // 1 / (values per unit near the centre)

    float lo = -CUDART_INF_F, hi = CUDART_INF_F;  // bracket [lo, hi): c_lo = #(v < lo) <= r_lo, c_hi = #(v < hi) > r_hi
    int c_lo = 0, c_hi = n_valid;
    // per-thread sign words of the rounds that set the bracket ends (bit 31-q: value q is below that end); their
    // difference is the set of values inside the final bracket, so the gather needs no further comparisons
    unsigned mask_lo = 0u, mask_hi = 0xffffffffu;
    if (!FULL) {
        mask_hi = 0u;
#pragma unroll
        for (int i = 0; i < 32; ++i) mask_hi = (mask_hi << 1) | (VAL(i) < CUDART_INF_F ? 1u : 0u);
    }
    float pivot = mean;
    for (int round = 0;; ++round) {
        if (round > 0) {
            if (c_hi - c_lo <= 32) break;
            if (round >= kTightenFrom) {
                // Still more than 32 values in the bracket after several rounds: quantised samples (an ADC's few hundred
                // levels, a sensor at rest), where the middle value is repeated more often than the final ranking holds and
                // the key-space bisection would need one round per bit.  Take the smallest and largest value INSIDE the
                // bracket: if they are equal, that value is both middle order statistics; otherwise they are the new ends
                // (no value lies between the old and the new ends, so counts and sign words stay valid) and every further
                // pivot in (vmin, vmax] removes at least one level from one side.
                const float2 mm = bracket_minmax(v2[0], v2[1], v2[2], v2[3], v2[4], v2[5], v2[6], v2[7], v2[8], v2[9], v2[10],
                                                 v2[11], v2[12], v2[13], v2[14], v2[15], ~mask_lo & mask_hi);
                if (lane == 0) {
                    shf[warp] = mm.x;
                    shf[8 + warp] = mm.y;
                }
                group_sync<T>(slot);
                float vmin = shf[0], vmax = shf[8];
#pragma unroll
                for (int w = 1; w < NW; ++w) {
                    vmin = fminf(vmin, shf[w]);
                    vmax = fmaxf(vmax, shf[8 + w]);
                }
                if (vmin == vmax) {
                    group_sync<T>(slot);
                    return vmin;
                }
                lo = vmin;
                hi = next_above(vmax);
            }
            const float lo_next = lo > -CUDART_INF_F ? next_above(lo) : -3.4028234e38f;
            if (!(lo_next < hi)) break;  // a single distinct value is left in the bracket
            const float want = (float)r_lo + 0.5f;
            if (lo == -CUDART_INF_F) {
                pivot = hi - 1.5f * fmaxf((float)c_hi - want, 1.f) * inv_density * (float)(1 << min(round - 1, 20));
            } else if (hi == CUDART_INF_F) {
                pivot = lo + 1.5f * fmaxf(want - (float)c_lo, 1.f) * inv_density * (float)(1 << min(round - 1, 20));
            } else if (round % 3 == 2) {  // key-space bisection: guarantees termination in <= 3*32 rounds
                const uint32_t a = ordered_key(lo), b = ordered_key(hi);
                pivot = key_value(a + ((b - a) >> 1), 0.f);
            } else {
                pivot = lo + (hi - lo) * __fdividef(want - (float)c_lo, (float)(c_hi - c_lo));
            }
            if (!(pivot > lo)) pivot = lo_next;       // also catches NaN
            if (!(pivot < hi)) pivot = lo_next;
        }
        // count(v < pivot): the sign bit of v - pivot is exact (x - x = +0, differences of floats never round across 0),
        // so one packed subtraction per value pair and one funnel shift per value collect 32 sign bits into a word
        unsigned signs = 0;
        const float2 pv = make_float2(pivot, pivot);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const float2 d = csub(v2[i], pv);
            signs = __funnelshift_l(__float_as_uint(d.x), signs, 1);
            signs = __funnelshift_l(__float_as_uint(d.y), signs, 1);
        }
        int cnt = __popc(signs);
        cnt = __reduce_add_sync(0xffffffffu, cnt);  // REDUX: one instruction instead of a shuffle tree
        // one barrier per round: the per-warp partial counts alternate between two banks of slots
        uint32_t *cslot = sh + 32 + 8 * (round & 1);
        if (lane == 0) cslot[warp] = (uint32_t)cnt;
        group_sync<T>(slot);
        int tot = 0;
#pragma unroll
        for (int w = 0; w < NW; ++w) tot += (int)cslot[w];
        if (tot <= r_lo) {  // both middle order statistics are >= pivot
            lo = pivot;
            c_lo = tot;
            mask_lo = signs;
        } else if (tot > r_hi) {  // both are < pivot
            hi = pivot;
            c_hi = tot;
            mask_hi = signs;
        } else {
            // even n and the pivot separates the two middles: lower = max{v < pivot}, upper = min{v >= pivot}
            float below = -CUDART_INF_F, above = CUDART_INF_F;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                if (VAL(i) < pivot) below = fmaxf(below, VAL(i));
                else above = fminf(above, VAL(i));
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                below = fmaxf(below, __shfl_xor_sync(0xffffffffu, below, o));
                above = fminf(above, __shfl_xor_sync(0xffffffffu, above, o));
            }
            if (lane == 0) {
                shf[warp] = below;
                shf[8 + warp] = above;
            }
            group_sync<T>(slot);
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                below = fmaxf(below, shf[w]);
                above = fminf(above, shf[8 + w]);
            }
            group_sync<T>(slot);
            return (below + above) * 0.5f;
        }
    }
    if (c_hi - c_lo > 32) {  // the loop ends like this only with ONE distinct value in the bracket: both middle order statistics
        group_sync<T>(slot);
        return lo;
    }
    // <= 32 values in [lo, hi): gather (the counter was zeroed in the prologue; the list area [0, 32) was last read
    // before the first round's barrier) and rank
    for (unsigned inb = ~mask_lo & mask_hi; inb; ) {  // at most 32 bits in the whole window (or one repeated value)
        const int q = __clz(inb);
        inb &= ~(0x80000000u >> q);
        const uint32_t pos = atomicAdd(&sh[48], 1u);
        if (pos < 32u) shf[pos] = reload(q);
    }
    group_sync<T>(slot);
    const int cnt = (int)sh[48];
    if (warp == 0) {
        float med;
        if (cnt > 32) {
            med = lo;  // every value in the bracket equals lo
        } else {
            const float mine = lane < cnt ? shf[lane] : CUDART_INF_F;
            int rank = 0;
            for (int j = 0; j < cnt; ++j) {
                const float other = shf[j];
                rank += (other < mine) || (other == mine && j < lane);
            }
            const uint32_t m_lo = __ballot_sync(0xffffffffu, lane < cnt && rank == r_lo - c_lo);
            const uint32_t m_hi = __ballot_sync(0xffffffffu, lane < cnt && rank == r_hi - c_lo);
            const float a = __shfl_sync(0xffffffffu, mine, __ffs(m_lo) - 1);
            const float b = __shfl_sync(0xffffffffu, mine, __ffs(m_hi) - 1);
            med = (a + b) * 0.5f;
        }
        if (lane == 0) shf[49] = med;
    }
    group_sync<T>(slot);
    return shf[49];  // nothing writes these words again before the window is done
}
#undef VAL

// Passes 1-3 of one window: HBM samples -> centred -> M-point complex FFT Z[k] in natural order in `s`
// (ends with the window barrier, so every thread may read any Z[k]).  Shared by the pipeline kernel
// (fft_f32_fast.cu) and the fused window->record kernel (fused_f32.cu).
template <int N, int CENTER, bool FULL, int PF_AHEAD = 0>
__device__ __forceinline__ void k1_forward(const float *__restrict__ samples, const int n_samples, const int64_t ld,
                                           const int64_t winc, const float2 *__restrict__ tw1, float2 *s, uint32_t *sel_w,
                                           float *red_w, const int wslot, const int t, const int64_t pf_batch = 0) {
    constexpr int center = CENTER;
    using P = Plan<N>;
    constexpr int M = N / 2, R1 = P::R1, R2 = P::R2, R3 = P::R3, T = M / 16;
    constexpr int S1 = R2 * R3;
    constexpr int LD = S1 + 16 / R1;
    constexpr int G1 = 16 / R1, G2 = 16 / R2, G3 = 16 / R3;
    float2 v[16];
    // ---------------- pass 1: load (coalesced 64-bit), centre, radix R1, twiddle, store [k1][c] ---------------------
    const float *x = samples + winc * ld;
    const float2 *z = reinterpret_cast<const float2 *>(x);
    const bool full = FULL;  // n_samples == N and 8-byte aligned rows: no predicates on the hot path
#pragma unroll
    for (int g = 0; g < G1; ++g) {
        const int c = t + T * g;
#pragma unroll
        for (int n1 = 0; n1 < R1; ++n1) {
            const int zi = n1 * S1 + c;
            float2 val;
            if (FULL) {
                val = __ldg(z + zi);
            } else {
                val.x = (2 * zi < n_samples) ? __ldg(x + 2 * zi) : 0.f;
                val.y = (2 * zi + 1 < n_samples) ? __ldg(x + 2 * zi + 1) : 0.f;
            }
            v[g * R1 + n1] = val;
        }
    }
    if (APDA_L2_PREFETCH && PF_AHEAD > 0) {  // samples of the window a later CTA of this slot will load
        const int64_t wn = winc + (int64_t)sm_count_reg() * PF_AHEAD;
        if (wn < pf_batch) l2_prefetch_span(samples + wn * ld, n_samples * (int)sizeof(float), t, T);
    }
    float shift = 0.f;
    if (center == APDA_CENTER_MEDIAN) {
        if (!full) {  // padding slots take no part in the median: park them at +inf for the selection
#pragma unroll
            for (int g = 0; g < G1; ++g)
#pragma unroll
                for (int n1 = 0; n1 < R1; ++n1) {
                    const int zi = n1 * S1 + t + T * g;
                    if (2 * zi >= n_samples) v[g * R1 + n1].x = CUDART_INF_F;
                    if (2 * zi + 1 >= n_samples) v[g * R1 + n1].y = CUDART_INF_F;
                }
        }
        // value q of this thread = component q&1 of complex slot q>>1 = (g, n1) -> sample 2*(n1*S1 + t + T*g) + (q&1)
        auto reload = [&](int q) {
            const int j = q >> 1;
            return __ldg(x + 2 * ((j % R1) * S1 + t + T * (j / R1)) + (q & 1));
        };
        shift = select_median<T, FULL>(v, n_samples, sel_w, t, wslot, reload);
        if (!full) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                if (v[i].x == CUDART_INF_F) v[i].x = 0.f;
                if (v[i].y == CUDART_INF_F) v[i].y = 0.f;
            }
        }
    } else if (center == APDA_CENTER_MEAN) {
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) acc += v[i].x + v[i].y;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if ((t & 31) == 0) red_w[t >> 5] = acc;
        group_sync<T>(wslot);
        float tot = 0.f;
#pragma unroll
        for (int w = 0; w < (T + 31) / 32; ++w) tot += red_w[w];
        shift = tot / (float)n_samples;
    }
#pragma unroll
    for (int g = 0; g < G1; ++g) {
        const int c = t + T * g;
        if (center != APDA_CENTER_NONE) {
#pragma unroll
            for (int n1 = 0; n1 < R1; ++n1) {
                const int zi = n1 * S1 + c;
                if (full) {
                    v[g * R1 + n1] = csub(v[g * R1 + n1], make_float2(shift, shift));
                } else {
                    if (2 * zi < n_samples) v[g * R1 + n1].x -= shift;
                    if (2 * zi + 1 < n_samples) v[g * R1 + n1].y -= shift;
                }
            }
        }
        fft_r<R1>(v + g * R1);
        s[c] = v[g * R1];
#pragma unroll
        for (int k1 = 1; k1 < R1; ++k1) s[k1 * LD + c] = cmul(v[g * R1 + k1], __ldg(tw1 + k1 * S1 + c));
    }
    group_sync<T>(wslot);

    // ---------------- pass 2: radix R2 over n2, in place; lanes run over k1 (conflict-free with the LD padding) ------
    const float2 *tw2 = const_tw2<N>();
#pragma unroll
    for (int g = 0; g < G2; ++g) {
        const int w = t + T * g;
        const int k1 = w % R1, n3 = w / R1;
        float2 *col = s + k1 * LD + n3;
#pragma unroll
        for (int n2 = 0; n2 < R2; ++n2) v[g * R2 + n2] = col[n2 * R3];
        fft_r<R2>(v + g * R2);
        col[0] = v[g * R2];
#pragma unroll
        for (int k2 = 1; k2 < R2; ++k2) col[k2 * R3] = cmul(v[g * R2 + k2], tw2[k2 * R3 + n3]);
    }
    group_sync<T>(wslot);

    // ---------------- pass 3: radix R3 over n3; results go back in natural bin order ---------------------------------
#pragma unroll
    for (int g = 0; g < G3; ++g) {
        const int w = t + T * g;
        const int k1 = w % R1, k2 = w / R1;
        const float2 *row = s + k1 * LD + k2 * R3;
#pragma unroll
        for (int n3 = 0; n3 < R3; ++n3) v[g * R3 + n3] = row[n3];
        fft_r<R3>(v + g * R3);
    }
    group_sync<T>(wslot);
#pragma unroll
    for (int g = 0; g < G3; ++g) {
        const int w = t + T * g;  // == k1 + R1*k2
#pragma unroll
        for (int k3 = 0; k3 < R3; ++k3) s[w + R1 * R2 * k3] = v[g * R3 + k3];
    }
    group_sync<T>(wslot);

}

}  // namespace
