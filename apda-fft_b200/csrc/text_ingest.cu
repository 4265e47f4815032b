// Batched ingest of sensor .log sample lines: many logs' sample regions -> sample arrays on the
// device, without a Python float() per token.
//
// Reference behaviour reproduced (utils/load_data.py:67-80): the region after the four header lines is split into lines
// (universal newlines) and each line at ';'; every non-empty piece goes through float(); pieces that do not parse and
// non-finite values are dropped; the rest are the samples, in file order.
// Device grammar (decided exactly): optional sign, digits with at most one '.', at most 15 significant digits and 22
// fractional digits -> value = M / 10^f with one correctly rounded division (both exact in binary64), which is what
// float() returns for such text (the sensor logs are written with "%8.6f", protocol_decoder.py:174).  Pieces that
// contain a character float() could never accept in a finite number are dropped; pieces float() might accept in a
// form the kernel does not decide (exponents, '_' digit separators, > 15 significant digits, non-ASCII) raise flag
// bit 0 for that log, and the host re-parses that log with float() (apda-fft_b200/utils/load_data.py).
#include "common.cuh"

namespace {

__device__ __forceinline__ bool is_delim(unsigned char c) { return c == ';' || c == '\n' || c == '\r'; }
__device__ __forceinline__ bool is_space(unsigned char c) { return c == ' ' || (c >= 9 && c <= 13) || (c >= 28 && c <= 31); }

// classify / parse the piece text[b, e): 1 = sample (value in *out), 0 = dropped, -1 = undecided (host fallback)
__device__ int parse_piece(const unsigned char *text, int64_t b, int64_t e, double *out) {
    while (b < e && is_space(text[b])) ++b;
    while (e > b && is_space(text[e - 1])) --e;
    if (b == e) return 0;  // float('') / float(' ') -> ValueError
    bool only_numeric_chars = true, non_ascii = false;
    for (int64_t i = b; i < e; ++i) {
        const unsigned char c = text[i];
        if (c >= 0x80) non_ascii = true;
        if (!((c >= '0' && c <= '9') || c == '+' || c == '-' || c == '.' || c == '_' || c == 'e' || c == 'E'))
            only_numeric_chars = false;
    }
    if (non_ascii) return -1;           // e.g. non-ASCII digits are legal for float()
    if (!only_numeric_chars) return 0;  // letters other than e/E: "nan", "inf", markers -> never a finite float
    int64_t i = b;
    bool neg = false;
    if (text[i] == '+' || text[i] == '-') {
        neg = text[i] == '-';
        ++i;
    }
    unsigned long long m = 0;
    int sig = 0, frac = 0, ndig = 0;
    bool seen_dot = false;
    for (; i < e; ++i) {
        const unsigned char c = text[i];
        if (c >= '0' && c <= '9') {
            ++ndig;
            if (m != 0 || c != '0') ++sig;
            if (sig > 15) return -1;
            m = m * 10ull + (unsigned long long)(c - '0');
            if (seen_dot) ++frac;
        } else if (c == '.' && !seen_dot) {
            seen_dot = true;
        } else {
            return -1;  // exponent, '_', second '.', inner sign: float() decides on the host
        }
    }
    if (ndig == 0) return 0;  // "+", ".", "-." -> ValueError
    if (frac > 22) return -1;
    double p = 1.0;
    for (int q = 0; q < frac; ++q) p *= 10.0;  // exact up to 1e22
    const double v = frac ? div_rn((double)m, p) : (double)m;
    *out = neg ? -v : v;
    return 1;
}

template <typename T>
__global__ void __launch_bounds__(256)
parse_samples_kernel(const unsigned char *__restrict__ text, const int64_t *__restrict__ offsets, int64_t ld,
                     T *__restrict__ samples, int *__restrict__ n_valid, int *__restrict__ flags) {
    __shared__ int warp_tot[8];
    __shared__ int flag_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t win = blockIdx.x;
    const int64_t r0 = offsets[win], r1 = offsets[win + 1];
    const int64_t per = (r1 - r0 + 255) / 256;
    const int64_t a = min(r0 + tid * per, r1), z = min(a + per, r1);
    if (tid == 0) flag_s = 0;
    __syncthreads();
    int cnt = 0, flag = 0;
    double v;
    for (int pass = 0; pass < 2; ++pass) {
        int pos = 0;
        if (pass == 1) {
            int incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            if (lane == 31) warp_tot[warp] = incl;
            __syncthreads();
            int base = 0, total = 0;
#pragma unroll
            for (int w = 0; w < 8; ++w) {
                if (w < warp) base += warp_tot[w];
                total += warp_tot[w];
            }
            pos = base + incl - cnt;
            if (tid == 0) {
                n_valid[win] = total < ld ? total : (int)ld;
                if (total > ld) atomicOr(&flag_s, 2);
            }
        }
        for (int64_t p = a; p < z; ++p) {
            // a piece belongs to the thread in whose range it starts
            if (is_delim(text[p]) || (p > r0 && !is_delim(text[p - 1]))) continue;
            int64_t q = p;
            while (q < r1 && !is_delim(text[q])) ++q;
            const int kind = parse_piece(text, p, q, &v);
            if (pass == 0) {
                if (kind == 1 && fabs(v) <= 1.7976931348623157e308) ++cnt;
                if (kind < 0) flag = 1;
            } else if (kind == 1 && fabs(v) <= 1.7976931348623157e308) {
                if (pos < ld) samples[win * ld + pos] = (T)v;
                ++pos;
            }
        }
        if (pass == 0 && flag) atomicOr(&flag_s, 1);
    }
    __syncthreads();
    if (tid == 0) flags[win] = flag_s;
}

}  // namespace

template <typename T>
int launch_parse_samples(apda_ctx *ctx, cudaStream_t st, const char *d_text, const int64_t *d_offsets, int64_t batch,
                         int64_t ld, T *d_samples, int *d_n_valid, int *d_flags) {
    if (batch == 0) return APDA_OK;
    parse_samples_kernel<T><<<(unsigned)batch, 256, 0, st>>>(reinterpret_cast<const unsigned char *>(d_text), d_offsets, ld,
                                                            d_samples, d_n_valid, d_flags);
    ctx->launches++;
    APDA_CUDA(cudaGetLastError());
    return APDA_OK;
}
template int launch_parse_samples<double>(apda_ctx *, cudaStream_t, const char *, const int64_t *, int64_t, int64_t,
                                          double *, int *, int *);
template int launch_parse_samples<float>(apda_ctx *, cudaStream_t, const char *, const int64_t *, int64_t, int64_t, float *,
                                         int *, int *);
