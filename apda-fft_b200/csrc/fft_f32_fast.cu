// K1 (fp32 fast form): the pipeline kernel (samples -> N complex bins in HBM).  The transform lives in fft_f32_fast.cuh.
#include <mutex>

#include "fft_f32_fast.cuh"

namespace {

// HALF_OUT: only bins [0, N/2) are written - all the pickers ever read (utils/get_peak_prominence.py:159,
// utils/get_peak_resolution.py:84); used where the spectrum is an internal workspace of the pipeline (apda_analyze_*),
// never where the caller receives the spectrum (start_fft materialises N bins).  SURVEY 8d "half-spectrum pipeline".
template <int N, int CENTER, bool FULL, bool HALF_OUT>
__global__ void __launch_bounds__(Plan<N>::WPB *(N / 32), Plan<N>::MINB)
fft_f32_fast_kernel(const float *__restrict__ samples, int n_samples, int64_t ld, int64_t batch,
                    const float2 *__restrict__ tw1,  // [R1][S1]: W_M^{c*k1}
                    const float2 *__restrict__ twu,  // [M]: -i/2 * W_N^k
                    float2 *__restrict__ spec, const int *__restrict__ nv) {
    using P = Plan<N>;
    constexpr int M = N / 2, R1 = P::R1, R2 = P::R2, R3 = P::R3, T = M / 16, WPB = P::WPB;
    constexpr int S1 = R2 * R3;            // columns of pass 1
    constexpr int LD = S1 + 16 / R1;       // padded row stride (complex elements): LD*R1 >= M
    static_assert(R1 * R2 * R3 == M, "plan");
    __shared__ __align__(16) float2 sm[WPB][R1 * LD];
    __shared__ uint32_t sel[WPB][64];
    __shared__ float red[WPB][8];

    const int wslot = threadIdx.x / T;  // window within the block
    const int t = threadIdx.x % T;
    const int64_t win = (int64_t)blockIdx.x * WPB + wslot;
    if (win >= batch) return;  // windows synchronise on their own named barrier, so a dead slot can leave
    if (nv && nv[win] != n_samples) return;  // ragged batch: windows of another length go to the general kernel
    const int64_t winc = win;
    float2 *s = sm[wslot];

    // (no L2 prefetch of later windows here: K1 is bound by HBM bandwidth itself, measured 8.52 -> 8.78 ns with it)
    k1_forward<N, CENTER, FULL>(samples, n_samples, ld, winc, tw1, s, sel[wslot], red[wslot], wslot, t);

    // ---------------- split step + 128-bit coalesced stores -----------------------------------------------------------
    float4 *out = reinterpret_cast<float4 *>(spec + win * (int64_t)N);
    const float4 *s4 = reinterpret_cast<const float4 *>(s);
    const float4 *twu4 = reinterpret_cast<const float4 *>(twu);
#pragma unroll
    for (int g = 0; g < (M / 2) / T; ++g) {
        const int p = t + T * g;  // bins k = 2p, 2p+1
        const float4 zk = s4[p];
        const float2 za = s[(M - 2 * p) & (M - 1)];  // partner of bin 2p
        const float2 zb = s[M - 2 * p - 1];          // partner of bin 2p+1
        const float4 w = __ldg(twu4 + p);
        // S = Z[k] + conj(Z[M-k]),  D = Z[k] - conj(Z[M-k]);  X[k] = S/2 + Wt*D,  X[k+M] = S/2 - Wt*D
        const float2 cj = make_float2(1.f, -1.f), ncj = make_float2(-1.f, 1.f), hf = make_float2(0.5f, 0.5f);
        const float2 zk0 = make_float2(zk.x, zk.y), zk1 = make_float2(zk.z, zk.w);
        const float2 s0 = pfma(za, cj, zk0), d0 = pfma(za, ncj, zk0);
        const float2 s1 = pfma(zb, cj, zk1), d1 = pfma(zb, ncj, zk1);
        const float2 t0 = cmul(d0, make_float2(w.x, w.y)), t1 = cmul(d1, make_float2(w.z, w.w));
        const float2 x0 = pfma(s0, hf, t0), x1 = pfma(s1, hf, t1);
        float4 lo4 = make_float4(x0.x, x0.y, x1.x, x1.y);
        if (p == 0) lo4.x = lo4.y = 0.f;  // reference: res[0] = 0
        out[p] = lo4;
        if (!HALF_OUT) {
            const float2 y0 = pfma(s0, hf, make_float2(-t0.x, -t0.y)), y1 = pfma(s1, hf, make_float2(-t1.x, -t1.y));
            out[M / 2 + p] = make_float4(y0.x, y0.y, y1.x, y1.y);
        }
    }
}

struct FastTables {
    float2 *tw1 = nullptr, *twu = nullptr;
};

}  // namespace

// per-context cache of the fast-path tables (keyed by N); constant tables are uploaded once per process/device
struct FastCache {
    std::map<int64_t, FastTables> tabs;
};
// one entry per context; a context belongs to one host thread, but different threads may hold different contexts, so the
// map itself is guarded (the entries are only touched by their owner)
static std::map<apda_ctx *, FastCache> g_fast;
static std::mutex g_fast_mu;
static FastCache &fast_cache_of(apda_ctx *ctx) {
    std::lock_guard<std::mutex> lock(g_fast_mu);
    return g_fast[ctx];  // std::map nodes are stable: the reference stays valid while other contexts come and go
}

template <int N>
static int fast_tables(apda_ctx *ctx, FastTables *out) {
    using P = Plan<N>;
    constexpr int M = N / 2, R1 = P::R1, R2 = P::R2, R3 = P::R3, S1 = R2 * R3;
    auto &cache = fast_cache_of(ctx).tabs;
    auto it = cache.find(N);
    if (it != cache.end()) {
        *out = it->second;
        return APDA_OK;
    }
    const double two_pi = 6.283185307179586476925286766559;
    std::vector<float2> h1((size_t)R1 * S1), hu(M), h2((size_t)R2 * R3);
    for (int k1 = 0; k1 < R1; ++k1)
        for (int c = 0; c < S1; ++c) {
            const double a = -two_pi * (double)((int64_t)k1 * c % M) / (double)M;
            h1[(size_t)k1 * S1 + c] = make_float2((float)cos(a), (float)sin(a));
        }
    for (int k = 0; k < M; ++k) {  // -i/2 * W_N^k = -i/2 (cos a + i sin a) = (sin a)/2 - i (cos a)/2, a = -2 pi k/N
        const double a = -two_pi * (double)k / (double)N;
        hu[k] = make_float2((float)(0.5 * sin(a)), (float)(-0.5 * cos(a)));
    }
    for (int k2 = 0; k2 < R2; ++k2)
        for (int n3 = 0; n3 < R3; ++n3) {
            const double a = -two_pi * (double)(k2 * n3) / (double)(R2 * R3);
            h2[(size_t)k2 * R3 + n3] = make_float2((float)cos(a), (float)sin(a));
        }
    FastTables ft;
    APDA_CUDA(cudaMalloc(&ft.tw1, h1.size() * sizeof(float2)));
    APDA_CUDA(cudaMalloc(&ft.twu, hu.size() * sizeof(float2)));
    APDA_CUDA(cudaMemcpy(ft.tw1, h1.data(), h1.size() * sizeof(float2), cudaMemcpyHostToDevice));
    APDA_CUDA(cudaMemcpy(ft.twu, hu.data(), hu.size() * sizeof(float2), cudaMemcpyHostToDevice));
    const void *sym = N == 1024 ? (const void *)c_tw2_1024 : N == 2048 ? (const void *)c_tw2_2048
                    : N == 4096 ? (const void *)c_tw2_4096 : (const void *)c_tw2_8192;
    APDA_CUDA(cudaMemcpyToSymbol(sym, h2.data(), h2.size() * sizeof(float2)));
    cache[N] = ft;
    *out = ft;
    return APDA_OK;
}

template <int N>
static int launch_fast_n(apda_ctx *ctx, cudaStream_t st, const float *d_samples, int64_t n_samples, int64_t ld,
                         int64_t batch, int flags, float *d_spec, const int *d_nv, bool half_out) {
    using P = Plan<N>;
    FastTables ft;
    APDA_TRY(fast_tables<N>(ctx, &ft));
    const int threads = P::WPB * (N / 32);
    const int64_t blocks = (batch + P::WPB - 1) / P::WPB;
    const bool full = n_samples == N && (reinterpret_cast<uintptr_t>(d_samples) & 7u) == 0 && (ld & 1) == 0;
    void (*kern)(const float *, int, int64_t, int64_t, const float2 *, const float2 *, float2 *, const int *);
#define APDA_K1(C, F) (half_out ? fft_f32_fast_kernel<N, C, F, true> : fft_f32_fast_kernel<N, C, F, false>)
    kern = flags == APDA_CENTER_MEDIAN ? (full ? APDA_K1(APDA_CENTER_MEDIAN, true) : APDA_K1(APDA_CENTER_MEDIAN, false))
           : flags == APDA_CENTER_MEAN ? (full ? APDA_K1(APDA_CENTER_MEAN, true) : APDA_K1(APDA_CENTER_MEAN, false))
                                       : (full ? APDA_K1(APDA_CENTER_NONE, true) : APDA_K1(APDA_CENTER_NONE, false));
#undef APDA_K1
    kern<<<(unsigned)blocks, threads, 0, st>>>(d_samples, (int)n_samples, ld, batch, ft.tw1, ft.twu,
                                               reinterpret_cast<float2 *>(d_spec), d_nv);
    ctx->launches++;
    APDA_CUDA(cudaGetLastError());
    return APDA_OK;
}

// device tables of the fast path for other translation units (the fused kernel); uploads them on first use
int fft_f32_fast_get_tables(apda_ctx *ctx, int64_t N, const float2 **tw1, const float2 **twu) {
    FastTables ft;
    switch (N) {
        case 1024: APDA_TRY(fast_tables<1024>(ctx, &ft)); break;
        case 2048: APDA_TRY(fast_tables<2048>(ctx, &ft)); break;
        case 4096: APDA_TRY(fast_tables<4096>(ctx, &ft)); break;
        case 8192: APDA_TRY(fast_tables<8192>(ctx, &ft)); break;
        default: apda_set_error("fft_f32_fast: unsupported N=%lld", (long long)N); return APDA_ERR_UNSUPPORTED;
    }
    *tw1 = ft.tw1;
    *twu = ft.twu;
    return APDA_OK;
}

bool fft_f32_fast_supports(int64_t N) { return N == 1024 || N == 2048 || N == 4096 || N == 8192; }

int launch_fft_f32_fast(apda_ctx *ctx, cudaStream_t st, const float *d_samples, int64_t n_samples, int64_t ld,
                        int64_t batch, int64_t N, int flags, float *d_spec, const int *d_nv, bool half_out) {
    switch (N) {
        case 1024: return launch_fast_n<1024>(ctx, st, d_samples, n_samples, ld, batch, flags, d_spec, d_nv, half_out);
        case 2048: return launch_fast_n<2048>(ctx, st, d_samples, n_samples, ld, batch, flags, d_spec, d_nv, half_out);
        case 4096: return launch_fast_n<4096>(ctx, st, d_samples, n_samples, ld, batch, flags, d_spec, d_nv, half_out);
        case 8192: return launch_fast_n<8192>(ctx, st, d_samples, n_samples, ld, batch, flags, d_spec, d_nv, half_out);
    }
    apda_set_error("fft_f32_fast: unsupported N=%lld", (long long)N);
    return APDA_ERR_UNSUPPORTED;
}

void fft_f32_fast_release(apda_ctx *ctx) {
    std::lock_guard<std::mutex> lock(g_fast_mu);
    auto it = g_fast.find(ctx);
    if (it == g_fast.end()) return;
    for (auto &kv : it->second.tabs) {
        cudaFree(kv.second.tw1);
        cudaFree(kv.second.twu);
    }
    g_fast.erase(it);
}
