// Ingest of the sensors' 16-bit wire samples: raw packet payload -> the samples the reference's
// FFT would have read back from its .log file, without the text detour.
//
// Reference chain reproduced bit for bit in fp64 (paths relative to the reference checkout):
//   protocol_decoder.py:116-144  decode_float_v2: 1-5-10 bit fields; exponent 31 -> inf/nan, exponent 0 ->
//                                sign * 0.00006103515 * (m/1024) (a non-IEEE subnormal scale), else sign * 2^(e-15) * (1 + m/1024)
//   protocol_decoder.py:146-175  decode_samples: value + first_value (axis baseline), formatted "%8.6f"
//   utils/load_data.py:67-80     the .log is parsed back with float(); non-finite tokens are dropped
// "%8.6f" followed by float() is a correctly rounded 6-decimal quantisation: n = round-half-even(x * 1e6) on the exact
// product (FMA residual), then the correctly rounded n / 1e6.  Dropped samples compact the window, so every window
// carries its own count (ragged batch, see apda_analyze_ragged_*).
#include "common.cuh"

namespace {

__device__ __forceinline__ bool wire16_decode(unsigned h, double first_value, double &out) {
    const unsigned e = (h >> 10) & 31u, m10 = h & 0x3ffu;
    if (e == 31u) return false;  // inf / nan: the log parser drops the token
    const double sign = (h & 0x8000u) ? -1.0 : 1.0;
    const double mant = (double)m10 / 1024.0;  // exact
    double val;
    if (e == 0u) val = m10 ? mul_rn(mul_rn(sign, 0.00006103515), mant) : 0.0;
    else val = sign * ldexp(1.0 + mant, (int)e - 15);  // exact
    const double x = add_rn(val, first_value);
    if (!(fabs(x) <= 1.7976931348623157e308)) return false;
    // "%8.6f" -> float(): correctly rounded 6-decimal quantisation
    const double p = 1e6;
    const double hi = mul_rn(x, p);
    const double lo = __fma_rn(x, p, -hi);
    double n = rint(hi);
    const double d = sub_rn(hi, n);
    if (d == 0.5 && lo > 0.0) n += 1.0;
    if (d == -0.5 && lo < 0.0) n -= 1.0;
    out = div_rn(n, p);
    return true;
}

__device__ __forceinline__ unsigned wire16_load(const unsigned char *p, int64_t i) {
    return ((unsigned)p[2 * i] << 8) | (unsigned)p[2 * i + 1];  // high byte first
}

// one CTA per window: count the kept samples per thread chunk, block scan, decode + compact in order
template <typename T>
__global__ void __launch_bounds__(256)
decode_wire16_kernel(const unsigned char *__restrict__ payload, int n_max, int64_t ld_bytes,
                     const double *__restrict__ first_value, T *__restrict__ samples, int64_t ld_out,
                     int *__restrict__ n_valid) {
    __shared__ int warp_tot[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t win = blockIdx.x;
    const unsigned char *src = payload + win * ld_bytes;
    const double fv = first_value[win];
    const int per = (n_max + 255) / 256;
    const int i0 = min(tid * per, n_max), i1 = min(i0 + per, n_max);
    int cnt = 0;
    double tmp;
    for (int i = i0; i < i1; ++i) cnt += wire16_decode(wire16_load(src, i), fv, tmp) ? 1 : 0;
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
    int pos = base + incl - cnt;
    T *dst = samples + win * ld_out;
    for (int i = i0; i < i1; ++i) {
        if (wire16_decode(wire16_load(src, i), fv, tmp)) dst[pos++] = (T)tmp;
    }
    if (tid == 0) n_valid[win] = total;
}

}  // namespace

template <typename T>
int launch_decode_wire16(apda_ctx *ctx, cudaStream_t st, const unsigned char *d_payload, int64_t n_max, int64_t ld_bytes,
                         int64_t batch, const double *d_first_value, T *d_samples, int64_t ld_out, int *d_n_valid) {
    if (batch == 0) return APDA_OK;
    decode_wire16_kernel<T><<<(unsigned)batch, 256, 0, st>>>(d_payload, (int)n_max, ld_bytes, d_first_value, d_samples,
                                                            ld_out, d_n_valid);
    ctx->launches++;
    APDA_CUDA(cudaGetLastError());
    return APDA_OK;
}
template int launch_decode_wire16<double>(apda_ctx *, cudaStream_t, const unsigned char *, int64_t, int64_t, int64_t,
                                          const double *, double *, int64_t, int *);
template int launch_decode_wire16<float>(apda_ctx *, cudaStream_t, const unsigned char *, int64_t, int64_t, int64_t,
                                         const double *, float *, int64_t, int *);
