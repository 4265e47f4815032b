// K2: multi-pass FFT for transforms that exceed shared memory (N > 2^13 fp64 / 2^14 fp32, up to 2^30).
//
// The reference's radix-2 DIT dataflow graph (metrics/fft_iterativa.py:38-70) is cut into P passes of q_p stages
// (sum q_p = log2 N, q_p <= 10 fp64 / 11 fp32).  Every pass moves the data through HBM exactly once:
//
//   head pass  (stages 1..q_1)   In bit-reversed order the first q_1 stages act on contiguous blocks of 2^q_1
//              outputs whose inputs are the samples j = j_hi * 2^(n-q_1) + j_lo with j_lo fixed: a COLUMN of the
//              input viewed as a [2^q_1][2^(n-q_1)] matrix.  A CTA takes C adjacent columns (C*sizeof(T)-byte
//              coalesced row segments), centres / pads / bit-reverses on the way into shared memory, runs the
//              stages on C contiguous column arrays and writes each column out as one contiguous block.
//   tail passes (stages s0+1..s0+q) act on index bits [s0, s0+q): a tile is [2^q rows (stride 2^s0)][C adjacent
//              columns], loaded and stored in place with coalesced C-element row segments; the stage twiddles
//              T_s[(r mod 2^(t-1)) * 2^s0 + column] are read coalesced from the recurrence table.
//
// The butterflies use the same host-built recurrence twiddle table and the same individually rounded operations as
// K1, so the fp64 result is bit-identical to the reference for every N (the table is what makes N >= 2^16 match to
// 1e-12: the reference's own twiddle drift reaches 3e-10 at N = 2^24).  fp32 runs the same passes on the rounded table.
// The exact median of up to 2^30 samples is a multi-CTA radix select (histogram passes over HBM).
#include <cuda.h>

#include <algorithm>

#include "common.cuh"

namespace {

constexpr int kThreads = 512;

template <typename T>
__device__ __forceinline__ void butterfly(typename vec2<T>::type &u, typename vec2<T>::type &v,
                                          const typename vec2<T>::type w) {
    // v*w with Python's complex product (ac-bd, ad+bc), then u+v, u-v; every operation rounds on its own
    const T vr = sub_rn(mul_rn(v.x, w.x), mul_rn(v.y, w.y));
    const T vi = add_rn(mul_rn(v.x, w.y), mul_rn(v.y, w.x));
    typename vec2<T>::type a, b;
    a.x = add_rn(u.x, vr);
    a.y = add_rn(u.y, vi);
    b.x = sub_rn(u.x, vr);
    b.y = sub_rn(u.y, vi);
    u = a;
    v = b;
}

// QR radix-2 stages (t0+1 .. t0+QR of this pass) on 2^QR register-resident values per work item: one shared-memory round
// trip and one barrier per QR stages instead of per stage, 2^QR - 1 twiddle loads per QR * 2^(QR-1) butterflies.
// Element (row r, column c) of the pass's working set lives at tile[r * rs + c * cs]; rows r = (hi << (t0+QR)) + (k << t0) + jr.
// Twiddle of stage t for row r: T_{s0+t}[((r mod 2^(t-1)) << s0) + lo0 + c]  (head pass: s0 = 0 and no column term - its
// columns are independent sub-transforms).  Same dataflow graph and individually rounded operations as before.
template <typename T, int QR, bool HEAD>
__device__ __forceinline__ void stage_round(typename vec2<T>::type *tile, int rs, int cs, int logC, int q, int t0,
                                            const typename vec2<T>::type *__restrict__ tw, int s0, int64_t lo0, int tid) {
    using V2 = typename vec2<T>::type;
    constexpr int R = 1 << QR;
    const int items = 1 << (logC + q - QR);
    for (int item = tid; item < items; item += kThreads) {
        const int c = item & ((1 << logC) - 1), rest = item >> logC;
        const int jr = rest & ((1 << t0) - 1), hi = rest >> t0;
        V2 *p = tile + (size_t)(((hi << (t0 + QR)) + jr)) * rs + (size_t)c * cs;
        const int kstride = rs << t0;
        V2 v[R];
#pragma unroll
        for (int k = 0; k < R; ++k) v[k] = p[k * kstride];
#pragma unroll
        for (int u = 1; u <= QR; ++u) {
            const int t = t0 + u;  // stage of this pass (1-based): pairs rows that differ in bit t-1
            const V2 *tab = tw + (((int64_t)1 << (s0 + t - 1)) - 1) + (HEAD ? 0 : lo0 + c);
            V2 w[R / 2];
#pragma unroll
            for (int m = 0; m < (1 << (u - 1)); ++m) w[m] = __ldg(tab + ((int64_t)(jr + (m << t0)) << s0));
#pragma unroll
            for (int k = 0; k < R; ++k)
                if ((k & (1 << (u - 1))) == 0) butterfly<T>(v[k], v[k + (1 << (u - 1))], w[k & ((1 << (u - 1)) - 1)]);
        }
#pragma unroll
        for (int k = 0; k < R; ++k) p[k * kstride] = v[k];
    }
}

// all q stages of a pass, three at a time (the remainder as 2+2, 2 or 1), one barrier per round
template <typename T, bool HEAD>
__device__ __forceinline__ void run_stages(typename vec2<T>::type *tile, int rs, int cs, int logC, int q,
                                           const typename vec2<T>::type *__restrict__ tw, int s0, int64_t lo0, int tid) {
    int t0 = 0;
    while (t0 < q) {
        const int left = q - t0;
        const int qr = left > 4 ? 3 : (left == 4 ? 2 : left);
        if (qr == 3) stage_round<T, 3, HEAD>(tile, rs, cs, logC, q, t0, tw, s0, lo0, tid);
        else if (qr == 2) stage_round<T, 2, HEAD>(tile, rs, cs, logC, q, t0, tw, s0, lo0, tid);
        else stage_round<T, 1, HEAD>(tile, rs, cs, logC, q, t0, tw, s0, lo0, tid);
        __syncthreads();
        t0 += qr;
    }
}

// ---- head pass ------------------------------------------------------------------------------------------------------
template <typename T, bool COMPLEX_IN>
__global__ void __launch_bounds__(kThreads)
large_head_kernel(const T *__restrict__ samples, int64_t n_samples, int64_t ld, int n, int q, int C,
                  const typename vec2<T>::type *__restrict__ tw, typename vec2<T>::type *__restrict__ spec,
                  const T *__restrict__ med_ptr) {
    using V2 = typename vec2<T>::type;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    V2 *work = reinterpret_cast<V2 *>(smem_raw);
    const int LDW = (1 << q) + 1;
    const int tid = threadIdx.x;
    const int64_t win = blockIdx.y;
    const int64_t N = (int64_t)1 << n;
    const int64_t lo0 = (int64_t)blockIdx.x * C;
    const T med = med_ptr ? med_ptr[win] : T(0);
    const int rows = 1 << q;

    const int logC = 31 - __clz(C);  // C is a power of two
    for (int e = tid; e < (C << q); e += kThreads) {
        const int c = e & (C - 1), jh = e >> logC;
        const int64_t j = ((int64_t)jh << (n - q)) + lo0 + c;
        V2 val;
        if (COMPLEX_IN) {
            val = reinterpret_cast<const V2 *>(samples)[win * N + j];
        } else {
            val.x = j < n_samples ? sub_rn(samples[win * ld + j], med) : T(0);
            val.y = T(0);
        }
        const int il = (int)(__brev((unsigned)jh) >> (32 - q));
        work[c * LDW + il] = val;
    }
    __syncthreads();

    run_stages<T, true>(work, 1, LDW, logC, q, tw, 0, 0, tid);

    V2 *out = spec + win * N;
    for (int e = tid; e < (C << q); e += kThreads) {
        const int c = e >> q, il = e & (rows - 1);
        const int64_t ih = (int64_t)(__brev((unsigned)(lo0 + c)) >> (32 - (n - q)));
        out[(ih << q) + il] = work[c * LDW + il];
    }
}

// ---- tail passes ----------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads)
large_tail_kernel(typename vec2<T>::type *__restrict__ spec, int n, int s0, int q, int C,
                  const typename vec2<T>::type *__restrict__ tw, int zero_dc) {
    using V2 = typename vec2<T>::type;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    V2 *tile = reinterpret_cast<V2 *>(smem_raw);
    const int LD = C + 1;
    const int tid = threadIdx.x;
    const int64_t N = (int64_t)1 << n;
    const int64_t tiles_lo = ((int64_t)1 << s0) / C;
    const int64_t hi = blockIdx.x / tiles_lo;
    const int64_t lo0 = (blockIdx.x % tiles_lo) * C;
    V2 *base = spec + (int64_t)blockIdx.y * N + (hi << (s0 + q)) + lo0;
    const int rows = 1 << q;

    const int logC = 31 - __clz(C);  // C is a power of two
    for (int e = tid; e < (C << q); e += kThreads) {
        const int c = e & (C - 1), r = e >> logC;
        tile[r * LD + c] = base[((int64_t)r << s0) + c];
    }
    __syncthreads();

    run_stages<T, false>(tile, LD, 1, logC, q, tw, s0, lo0, tid);

    for (int e = tid; e < (C << q); e += kThreads) {
        const int c = e & (C - 1), r = e >> logC;
        V2 val = tile[r * LD + c];
        if (zero_dc && hi == 0 && lo0 == 0 && r == 0 && c == 0) val.x = val.y = T(0);  // reference: res[0] = 0
        base[((int64_t)r << s0) + c] = val;
    }
    (void)rows;
}

// ---- tail pass, TMA staged -----------------------------------------------------------------------------------------
// Same arithmetic as large_tail_kernel; the [2^q][C] tile is brought in by cp.async.bulk.tensor (one elected thread,
// completion on an mbarrier) and written back in place by a bulk tensor store, so no thread spends registers or LSU
// issue slots on the HBM traffic.  The tensor map views the spectra as [batch * n_hi][2^q][2^s0] complex values
// (innermost dimension counted in scalars: 2 per complex value).
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <typename T>
__global__ void __launch_bounds__(kThreads, sizeof(T) == 4 ? 3 : 1)  // fp32: 3 tiles per SM in flight (load / compute / store)
large_tail_tma_kernel(const __grid_constant__ CUtensorMap tmap, int n, int s0, int q, int C,
                      const typename vec2<T>::type *__restrict__ tw, int zero_dc) {
    using V2 = typename vec2<T>::type;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long bar;
    // TMA wants a 128-byte aligned shared-memory destination; the dynamic segment only promises 16
    V2 *tile = reinterpret_cast<V2 *>(smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u));
    const int tid = threadIdx.x;
    const int64_t tiles_lo = ((int64_t)1 << s0) / C;
    const int64_t n_hi = (int64_t)1 << (n - s0 - q);
    const int64_t hi = blockIdx.x / tiles_lo;
    const int64_t lo0 = (blockIdx.x % tiles_lo) * C;
    const int rows = 1 << q;
    const int box_rows = rows < 256 ? rows : 256;
    const int c0 = (int)(2 * lo0), c2 = (int)((int64_t)blockIdx.y * n_hi + hi);
    const uint32_t bar_a = smem_u32(&bar);

    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        const uint32_t bytes = (uint32_t)((size_t)rows * C * sizeof(V2));
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(bytes) : "memory");
        for (int r0 = 0; r0 < rows; r0 += box_rows) {
            const uint32_t dst = smem_u32(tile + (size_t)r0 * C);
            asm volatile(
                "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                ::"r"(dst), "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(c0), "r"(r0), "r"(c2), "r"(bar_a)
                : "memory");
        }
    }
    {  // every thread waits for the tile (phase 0 of the barrier)
        uint32_t done = 0;
        while (!done) {
            asm volatile(
                "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(done)
                : "r"(bar_a), "r"(0u)
                : "memory");
        }
    }

    run_stages<T, false>(tile, C, 1, 31 - __clz(C), q, tw, s0, lo0, tid);
    if (zero_dc && hi == 0 && lo0 == 0 && tid == 0) tile[0].x = tile[0].y = T(0);  // reference: res[0] = 0
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the bulk store
    __syncthreads();
    if (tid == 0) {
        for (int r0 = 0; r0 < rows; r0 += box_rows) {
            const uint32_t src = smem_u32(tile + (size_t)r0 * C);
            asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(
                             reinterpret_cast<uint64_t>(&tmap)),
                         "r"(c0), "r"(r0), "r"(c2), "r"(src)
                         : "memory");
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn lookup_encode_fn() {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    EncodeTiledFn fn = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
        fn = reinterpret_cast<EncodeTiledFn>(p);
    (void)cudaGetLastError();
    return fn;
}
static EncodeTiledFn get_encode_fn() {
    static const EncodeTiledFn fn = lookup_encode_fn();  // initialised once, thread-safe
    return fn;
}

// tensor map of one tail pass; false if the driver entry point is missing or rejects the shape (caller falls back)
template <typename T>
static bool make_tail_tmap(CUtensorMap *tm, T *d_spec, int n, int s0, int q, int C, int64_t batch) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)2 << s0, (cuuint64_t)1 << q, (cuuint64_t)batch << (n - s0 - q)};
    const cuuint64_t strides[2] = {((cuuint64_t)1 << s0) * 2 * sizeof(T), ((cuuint64_t)1 << (s0 + q)) * 2 * sizeof(T)};
    const cuuint32_t box[3] = {(cuuint32_t)(2 * C), (cuuint32_t)std::min(1 << q, 256), 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUtensorMapDataType dt = sizeof(T) == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    if (box[0] > 256 || dims[2] >= ((cuuint64_t)1 << 32)) return false;
    return enc(tm, dt, 3, d_spec, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// ---- exact median of one long window: MSB-first 8-bit radix select with HBM histogram passes -----------------------
template <typename T>
struct KeyT;
template <>
struct KeyT<double> {
    using type = uint64_t;
};
template <>
struct KeyT<float> {
    using type = uint32_t;
};

struct SelectState {            // device resident
    unsigned long long prefix;  // key bits decided so far
    unsigned long long mask;
    long long rank;             // rank still to resolve inside the prefix bucket
    unsigned hist[256];
    unsigned long long found[2];  // lower / upper middle keys
    // compaction after the first two digit passes: the bucket that holds the median, copied out
    unsigned long long above_min;  // smallest key in the buckets above it
    unsigned long long m;          // elements in the bucket
};

template <typename T>
__global__ void __launch_bounds__(256) select_hist_kernel(const T *__restrict__ x, int64_t n, int shift, SelectState *st) {
    using K = typename KeyT<T>::type;
    __shared__ unsigned h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const K prefix = (K)st->prefix, mask = (K)st->mask;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const K k = ordered_key(x[i]);
        if ((k & mask) == prefix) atomicAdd(&h[(unsigned)(k >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (h[threadIdx.x]) atomicAdd(&st->hist[threadIdx.x], h[threadIdx.x]);
}

__global__ void __launch_bounds__(256) select_pick_kernel(SelectState *st, int shift, int which, int last) {
    // 256 threads: inclusive scan of the digit histogram, the digit whose cumulative count first exceeds the rank wins
    __shared__ unsigned long long cum[256];
    const int d = threadIdx.x;
    const unsigned mine = st->hist[d];
    cum[d] = mine;
    __syncthreads();
    for (int o = 1; o < 256; o <<= 1) {
        const unsigned long long add = d >= o ? cum[d - o] : 0ull;
        __syncthreads();
        cum[d] += add;
        __syncthreads();
    }
    const long long rank = st->rank;
    const unsigned long long before = cum[d] - mine;
    const bool winner = rank >= (long long)before && rank < (long long)cum[d];
    const bool none = d == 255 && rank >= (long long)cum[255];  // cannot happen for a consistent state: keep digit 255
    __syncthreads();
    st->hist[d] = 0;
    if (winner || none) {
        st->rank = rank - (long long)(none ? cum[254] : before);
        st->prefix |= (unsigned long long)d << shift;
        st->mask |= 255ull << shift;
        if (last) st->found[which] = st->prefix;
    }
}

__global__ void select_reset_kernel(SelectState *st, long long rank) {
    if (threadIdx.x == 0) {
        st->prefix = 0;
        st->mask = 0;
        st->rank = rank;
        st->above_min = ~0ull;
        st->m = 0;
    }
    st->hist[threadIdx.x] = 0;
}

// upper middle order statistic from the lower one in a single pass: it equals the lower key when that key is
// duplicated across the midpoint, else it is the smallest key above it
template <typename T>
__global__ void __launch_bounds__(256) select_upper_kernel(const T *__restrict__ x, int64_t n, SelectState *st) {
    using K = typename KeyT<T>::type;
    const K lo = (K)st->found[0];
    unsigned long long le = 0, above = ~0ull;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const K k = ordered_key(x[i]);
        le += (k <= lo);
        if (k > lo && (unsigned long long)k < above) above = (unsigned long long)k;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        le += __shfl_xor_sync(0xffffffffu, le, o);
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, above, o);
        above = other < above ? other : above;
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&st->prefix, le);       // prefix / mask are free after the digit passes: reuse as count / min
        atomicMin(&st->mask, above);
    }
}

template <typename T>
__global__ void select_finish_kernel(SelectState *st, int64_t n, T *med_out) {
    using K = typename KeyT<T>::type;
    const K lo = (K)st->found[0];
    const K hi = ((long long)st->prefix >= (long long)(n / 2 + 1)) ? lo : (K)st->mask;
    const T a = key_value(lo, T(0)), b = key_value(hi, T(0));
    *med_out = div_rn(add_rn(a, b), T(2));  // statistics.median: middle value, or (a + b) / 2 for even n
}

__global__ void select_prepare_upper_kernel(SelectState *st) {
    st->prefix = 0;
    st->mask = ~0ull;
}

// After two digit passes the bucket of the median holds a small fraction of the window (sensor samples: a few percent).
// One more pass over the window copies that bucket out (warp-aggregated append) and records the smallest key of the
// buckets above it; everything else runs on the copy, in one CTA.
template <typename T>
__global__ void __launch_bounds__(256) select_compact_kernel(const T *__restrict__ x, int64_t n, SelectState *st,
                                                             T *__restrict__ bucket) {
    using K = typename KeyT<T>::type;
    __shared__ unsigned long long blk_min[8];
    const K prefix = (K)st->prefix, mask = (K)st->mask;
    const int lane = threadIdx.x & 31;
    unsigned long long above = ~0ull;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t rounds = (n + stride - 1) / stride;
    for (int64_t it = 0; it < rounds; ++it) {
        const int64_t i = it * stride + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        const bool valid = i < n;
        const T v = valid ? x[i] : T(0);
        const K k = ordered_key(v);
        const bool in = valid && (k & mask) == prefix;
        if (valid && (k & mask) > prefix && (unsigned long long)k < above) above = (unsigned long long)k;
        const unsigned hit = __ballot_sync(0xffffffffu, in);
        if (hit) {
            unsigned long long base = 0;
            if (lane == __ffs(hit) - 1) base = atomicAdd(&st->m, (unsigned long long)__popc(hit));
            base = __shfl_sync(0xffffffffu, base, __ffs(hit) - 1);
            if (in) bucket[base + __popc(hit & ((1u << lane) - 1u))] = v;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, above, o);
        above = other < above ? other : above;
    }
    if (lane == 0) blk_min[threadIdx.x >> 5] = above;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) above = blk_min[w] < above ? blk_min[w] : above;
        if (above != ~0ull) atomicMin(&st->above_min, above);
    }
}

// the remaining digits, the upper middle value and statistics.median itself, on the copied bucket, one CTA
template <typename T>
__global__ void __launch_bounds__(1024) select_small_kernel(const T *__restrict__ bucket, SelectState *st, int64_t n,
                                                             int first_shift, T *med_out) {
    using K = typename KeyT<T>::type;
    __shared__ unsigned h[256];
    __shared__ unsigned long long s_prefix, s_mask, s_le, s_above;
    __shared__ long long s_rank;
    const long long m = (long long)st->m;
    if (threadIdx.x == 0) {
        s_prefix = st->prefix;
        s_mask = st->mask;
        s_rank = st->rank;  // rank of the lower middle inside the bucket
        s_le = 0;
        s_above = ~0ull;
    }
    for (int shift = first_shift; shift >= 0; shift -= 8) {
        if (threadIdx.x < 256) h[threadIdx.x] = 0;
        __syncthreads();
        const K prefix = (K)s_prefix, mask = (K)s_mask;
        for (long long i = threadIdx.x; i < m; i += blockDim.x) {
            const K k = ordered_key(bucket[i]);
            if ((k & mask) == prefix) atomicAdd(&h[(unsigned)(k >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            long long rank = s_rank;
            unsigned long long acc = 0;
            int digit = 255;
            for (int d = 0; d < 256; ++d) {
                if (rank < (long long)(acc + h[d])) {
                    digit = d;
                    break;
                }
                acc += h[d];
            }
            s_rank = rank - (long long)acc;
            s_prefix |= (unsigned long long)digit << shift;
            s_mask |= 255ull << shift;
        }
        __syncthreads();
    }
    // lower middle = s_prefix.  Upper middle: the same key if its duplicates reach across the midpoint, else the smallest
    // key above it (inside the bucket, or the smallest key of the buckets above)
    const K lo = (K)s_prefix;
    unsigned long long le = 0, above = ~0ull;
    for (long long i = threadIdx.x; i < m; i += blockDim.x) {
        const K k = ordered_key(bucket[i]);
        le += (k <= lo);
        if (k > lo && (unsigned long long)k < above) above = (unsigned long long)k;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        le += __shfl_xor_sync(0xffffffffu, le, o);
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, above, o);
        above = other < above ? other : above;
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&s_le, le);
        atomicMin(&s_above, above);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        // elements below the bucket: the rank the search started with minus the rank left when the bucket was fixed
        const long long below = (long long)((n - 1) / 2) - st->rank;
        const unsigned long long up = s_above < st->above_min ? s_above : st->above_min;
        const K hi = (below + (long long)s_le >= (long long)(n / 2 + 1)) ? lo : (K)up;
        const T a = key_value(lo, T(0)), b = key_value(hi, T(0));
        *med_out = div_rn(add_rn(a, b), T(2));  // statistics.median: middle value, or (a + b) / 2 for even n
    }
}

template <typename T>
int large_median(apda_ctx *ctx, cudaStream_t st, const T *d_x, int64_t n, SelectState *state, T *d_med, T *d_bucket) {
    const int bits = (int)sizeof(T) * 8;
    const int grid = (int)std::min<int64_t>((n + 255) / 256, (int64_t)ctx->sm_count * 8);
    select_reset_kernel<<<1, 256, 0, st>>>(state, (long long)((n - 1) / 2));
    for (int shift = bits - 8; shift >= bits - 16; shift -= 8) {  // two most significant digits on the whole window
        select_hist_kernel<T><<<grid, 256, 0, st>>>(d_x, n, shift, state);
        select_pick_kernel<<<1, 256, 0, st>>>(state, shift, 0, 0);
    }
    select_compact_kernel<T><<<grid, 256, 0, st>>>(d_x, n, state, d_bucket);
    select_small_kernel<T><<<1, 1024, 0, st>>>(d_bucket, state, n, bits - 24, d_med);
    ctx->launches += 7;
    APDA_CUDA(cudaGetLastError());
    return APDA_OK;
}

// the earlier form (eight full histogram passes + one pass for the upper middle); kept as the reference implementation for
// apda_ctx_set_generic_only
template <typename T>
int large_median_passes(apda_ctx *ctx, cudaStream_t st, const T *d_x, int64_t n, SelectState *state, T *d_med) {
    const int bits = (int)sizeof(T) * 8;
    const int grid = (int)std::min<int64_t>((n + 255) / 256, (int64_t)ctx->sm_count * 8);
    select_reset_kernel<<<1, 256, 0, st>>>(state, (long long)((n - 1) / 2));
    for (int shift = bits - 8; shift >= 0; shift -= 8) {
        select_hist_kernel<T><<<grid, 256, 0, st>>>(d_x, n, shift, state);
        select_pick_kernel<<<1, 256, 0, st>>>(state, shift, 0, shift == 0);
        ctx->launches += 2;
    }
    select_prepare_upper_kernel<<<1, 1, 0, st>>>(state);
    select_upper_kernel<T><<<grid, 256, 0, st>>>(d_x, n, state);
    select_finish_kernel<T><<<1, 1, 0, st>>>(state, n, d_med);
    ctx->launches += 4;
    APDA_CUDA(cudaGetLastError());
    return APDA_OK;
}

struct PassPlan {
    int npass;
    int q[8];
};
static PassPlan make_plan(int n, int qmax) {
    PassPlan p;
    p.npass = (n + qmax - 1) / qmax;
    if (p.npass < 2) p.npass = 2;
    const int base = n / p.npass, rem = n % p.npass;
    for (int i = 0; i < p.npass; ++i) p.q[i] = base + (i < rem ? 1 : 0);
    return p;
}

}  // namespace

template <typename T>
int launch_fft_large(apda_ctx *ctx, cudaStream_t st, const T *d_samples, int64_t n_samples, int64_t ld, int64_t batch,
                     int64_t N, int flags, T *d_spec, bool complex_input) {
    using V2 = typename vec2<T>::type;
    const int n = ilog2_i64(N);
    if (n > 30 || batch > 65535) {
        apda_set_error("fft_large: N=2^%d batch=%lld outside the supported range", n, (long long)batch);
        return APDA_ERR_UNSUPPORTED;
    }
    TwiddleTables tw;
    APDA_TRY(apda_get_twiddles(ctx, N, &tw));
    const V2 *twp = sizeof(T) == 8 ? reinterpret_cast<const V2 *>(tw.d64) : reinterpret_cast<const V2 *>(tw.d32);
    const int max_tile = sizeof(T) == 8 ? 4096 : 8192;  // complex elements per 64 KB tile: 3 CTAs per SM overlap load, compute and store
    const PassPlan plan = make_plan(n, sizeof(T) == 8 ? 10 : 11);

    // centring constant per window
    T *d_med = nullptr;
    if (!complex_input && flags != APDA_CENTER_NONE) {
        // per-stream scratch: the two host-pipeline streams may run long transforms concurrently
        const size_t med_bytes = ((size_t)batch * sizeof(T) + 255) & ~(size_t)255;
        const size_t need = 2048 + med_bytes + (size_t)n_samples * sizeof(T);  // state, medians, bucket copy (worst case: all samples)
        auto &slot = ctx->stream_scratch[st];
        if (need > slot.second) {
            APDA_CUDA(cudaStreamSynchronize(st));
            APDA_TRY(apda_reserve(&slot.first, &slot.second, need));
        }
        SelectState *state = reinterpret_cast<SelectState *>(slot.first);
        d_med = reinterpret_cast<T *>(reinterpret_cast<char *>(slot.first) + 2048);
        T *d_bucket = reinterpret_cast<T *>(reinterpret_cast<char *>(slot.first) + 2048 + med_bytes);
        for (int64_t w = 0; w < batch; ++w) {
            if (ctx->generic_only) APDA_TRY(large_median_passes<T>(ctx, st, d_samples + w * ld, n_samples, state, d_med + w));
            else APDA_TRY(large_median<T>(ctx, st, d_samples + w * ld, n_samples, state, d_med + w, d_bucket));
        }
    }

    {  // head pass
        const int q = plan.q[0];
        const int64_t cols = (int64_t)1 << (n - q);
        const int C = (int)std::min<int64_t>(std::min<int64_t>(max_tile >> q, 64), cols);
        const size_t smem = (size_t)C * ((1u << q) + 1) * sizeof(V2);
        auto kern = complex_input ? large_head_kernel<T, true> : large_head_kernel<T, false>;
        APDA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dim3 grid((unsigned)(cols / C), (unsigned)batch);
        kern<<<grid, kThreads, smem, st>>>(d_samples, n_samples, ld, n, q, C, twp, reinterpret_cast<V2 *>(d_spec), d_med);
        ctx->launches++;
        APDA_CUDA(cudaGetLastError());
    }
    int s0 = plan.q[0];
    for (int p = 1; p < plan.npass; ++p) {
        const int q = plan.q[p];
        const int64_t cols = (int64_t)1 << s0;
        const int C = (int)std::min<int64_t>(std::min<int64_t>(max_tile >> q, 64), cols);
        const int64_t tiles = (cols / C) * ((int64_t)1 << (n - s0 - q));
        dim3 grid((unsigned)tiles, (unsigned)batch);
        const int zero_dc = (!complex_input && p == plan.npass - 1) ? 1 : 0;
        CUtensorMap tm;
        if (!ctx->generic_only && make_tail_tmap<T>(&tm, d_spec, n, s0, q, C, batch)) {
            const size_t smem = (size_t)C * (1u << q) * sizeof(V2) + 128;
            APDA_CUDA(cudaFuncSetAttribute(large_tail_tma_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            large_tail_tma_kernel<T><<<grid, kThreads, smem, st>>>(tm, n, s0, q, C, twp, zero_dc);
        } else {
            const size_t smem = (size_t)(C + 1) * (1u << q) * sizeof(V2);
            APDA_CUDA(cudaFuncSetAttribute(large_tail_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            large_tail_kernel<T><<<grid, kThreads, smem, st>>>(reinterpret_cast<V2 *>(d_spec), n, s0, q, C, twp, zero_dc);
        }
        ctx->launches++;
        APDA_CUDA(cudaGetLastError());
        s0 += q;
    }
    return APDA_OK;
}
template int launch_fft_large<double>(apda_ctx *, cudaStream_t, const double *, int64_t, int64_t, int64_t, int64_t, int,
                                      double *, bool);
template int launch_fft_large<float>(apda_ctx *, cudaStream_t, const float *, int64_t, int64_t, int64_t, int64_t, int,
                                     float *, bool);
