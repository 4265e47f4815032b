/* ORACLE (test infrastructure, never the product path).
 *
 * Plain-C restatement of the reference FFT front end, batched, for the sizes
 * where the pure-Python port (oracle/ref_port.py) is too slow (10k-1M windows,
 * N = 2^20 .. 2^24).  Compile with -O2 -ffp-contract=off: every product and sum
 * below must round separately, as CPython float/complex arithmetic does.
 *
 * Reference behaviour followed (paths relative to the reference checkout):
 *   metrics/fft_iterativa.py:5-11   median centring (statistics.median)
 *   metrics/fft_iterativa.py:13-22  zero padding to 2^k after centring
 *   metrics/fft_iterativa.py:24-36  bit-reversal permutation
 *   metrics/fft_iterativa.py:38-70  DIT radix-2 butterflies, twiddle recurrence w *= w_m
 *   metrics/fft_iterativa.py:85     bin 0 forced to zero
 *   utils/get_peak_prominence.py:159 / get_peak_resolution.py:84  |X| via abs(complex) == glibc hypot
 *
 * Pinned bit-for-bit against the live reference by tests/golden/make_golden.py
 * and tests/test_oracle.py (KAT spectra SHA-256).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_PI 3.141592653589793 /* == cmath.pi */

static int cmp_double(const void *a, const void *b) {
    double x = *(const double *)a, y = *(const double *)b;
    return (x > y) - (x < y);
}

/* Flat twiddle table: stage with half-span h (h = 1,2,4,..,N/2) starts at offset h-1
 * and holds h entries; entry j is the value of w before the j-th butterfly of a block. */
void oracle_twiddle_table(int64_t N, double *tw /* 2*(N-1) doubles, interleaved */) {
    for (int64_t h = 1; h < N; h <<= 1) {
        double theta = (-2.0 * ORACLE_PI) / (double)(2 * h);
        double cr = cos(theta), ci = sin(theta);
        double wr = 1.0, wi = 0.0;
        double *t = tw + 2 * (h - 1);
        for (int64_t j = 0; j < h; ++j) {
            t[2 * j] = wr;
            t[2 * j + 1] = wi;
            double nr = wr * cr - wi * ci;
            double ni = wr * ci + wi * cr;
            wr = nr;
            wi = ni;
        }
    }
}

/* statistics.median: sort; odd -> middle, even -> (a+b)/2 */
double oracle_median(const double *x, int64_t n) {
    double *tmp = (double *)malloc(sizeof(double) * (size_t)n);
    memcpy(tmp, x, sizeof(double) * (size_t)n);
    qsort(tmp, (size_t)n, sizeof(double), cmp_double);
    double m = (n & 1) ? tmp[n / 2] : (tmp[n / 2 - 1] + tmp[n / 2]) / 2.0;
    free(tmp);
    return m;
}

static void butterflies(double *y /* interleaved, N complex */, int64_t N, const double *tw) {
    for (int64_t h = 1; h < N; h <<= 1) {
        const double *t = tw + 2 * (h - 1);
        for (int64_t base = 0; base < N; base += 2 * h) {
            for (int64_t j = 0; j < h; ++j) {
                double *lo = y + 2 * (base + j);
                double *hi = lo + 2 * h;
                double wr = t[2 * j], wi = t[2 * j + 1];
                double vr = hi[0] * wr - hi[1] * wi;
                double vi = hi[0] * wi + hi[1] * wr;
                double ur = lo[0], ui = lo[1];
                lo[0] = ur + vr;
                lo[1] = ui + vi;
                hi[0] = ur - vr;
                hi[1] = ui - vi;
            }
        }
    }
}

static inline int64_t bitrev(int64_t i, int bits) {
    int64_t r = 0;
    for (int b = 0; b < bits; ++b) r |= ((i >> b) & 1) << (bits - 1 - b);
    return r;
}

/* start_fft over a batch: samples[b*ld .. b*ld+n_samples) -> out[b*N*2 ..], N = 2^k >= n_samples.
 * tw = oracle_twiddle_table(N). */
void oracle_start_fft_batch(const double *samples, int64_t n_samples, int64_t ld, int64_t batch,
                            int64_t N, const double *tw, double *out) {
    int bits = 0;
    while (((int64_t)1 << bits) < N) ++bits;
    for (int64_t b = 0; b < batch; ++b) {
        const double *x = samples + b * ld;
        double *y = out + 2 * N * b;
        double med = n_samples > 0 ? oracle_median(x, n_samples) : 0.0;
        for (int64_t i = 0; i < N; ++i) {
            int64_t src = bitrev(i, bits);
            y[2 * i] = src < n_samples ? x[src] - med : 0.0;
            y[2 * i + 1] = 0.0;
        }
        butterflies(y, N, tw);
        y[0] = 0.0;
        y[1] = 0.0;
    }
}

/* fft() on complex input (reference fft(x) called directly), single transform, in place semantics on out. */
void oracle_fft_c2c(const double *in, int64_t N, const double *tw, double *out) {
    int bits = 0;
    while (((int64_t)1 << bits) < N) ++bits;
    for (int64_t i = 0; i < N; ++i) {
        int64_t src = bitrev(i, bits);
        out[2 * i] = in[2 * src];
        out[2 * i + 1] = in[2 * src + 1];
    }
    butterflies(out, N, tw);
}

/* magnitudes of bins [0, half) with glibc hypot (== CPython abs(complex)) */
void oracle_half_magnitudes(const double *spec, int64_t half, double *mags) {
    for (int64_t i = 0; i < half; ++i) mags[i] = hypot(spec[2 * i], spec[2 * i + 1]);
}
