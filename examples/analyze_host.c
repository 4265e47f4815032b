/* Minimal C host of libapda_b200.so: the reference's work_flow_fft body (GT_FFT_v5.py:635-642: start_fft + picker) for
 * a batch of windows held in host memory, through the C ABI only (no CUDA headers, no Python).
 *
 *   gcc -O2 -I include examples/analyze_host.c -L apda-fft_b200 -lapda_b200 -Wl,-rpath,$PWD/apda-fft_b200 -lm -o analyze_host
 *   ./analyze_host            # needs a B200; prints the peaks of a few synthetic 3-tone windows
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "apda_b200.h"

int main(void) {
    enum { N = 4096, B = 8 };
    const double fs = 125.0, two_pi = 6.283185307179586;
    float *x = (float *)malloc(sizeof(float) * N * B);
    apda_peak_rec *rec = (apda_peak_rec *)calloc(B, sizeof(apda_peak_rec));
    if (!x || !rec) return 2;
    for (int w = 0; w < B; ++w)
        for (int i = 0; i < N; ++i)
            x[(size_t)w * N + i] = (float)(0.5 * sin(two_pi * (101.6 + w) * i / N) + 0.3 * sin(two_pi * 252.4 * i / N + 0.3) +
                                           0.2 * sin(two_pi * 498.0 * i / N + 1.1));
    apda_ctx *ctx = NULL;
    if (apda_ctx_create(0, &ctx) != APDA_OK) {
        fprintf(stderr, "apda_ctx_create: %s\n", apda_last_error());
        return 1; /* no sm_100 device: there is no CPU fallback */
    }
    /* flexible-structure picker (get_top_peaks_prominence, k = 4), exact-median centring, 128-byte records */
    int rc = apda_analyze_f32_host(ctx, x, N, N, B, N, APDA_CENTER_MEDIAN, 1, fs, NULL, 4, 5, rec);
    if (rc != APDA_OK) {
        fprintf(stderr, "apda_analyze_f32_host: %s\n", apda_last_error());
        return 1;
    }
    for (int w = 0; w < B; ++w) {
        printf("window %d: %d peaks:", w, rec[w].count);
        for (int a = 0; a < rec[w].count; ++a)
            printf("  idx %d (%.4f Hz) mag %.4f", rec[w].pk[a].idx, rec[w].pk[a].idx * (fs / N), rec[w].pk[a].mag);
        printf("\n");
    }
    /* the same batch sharded over two contexts by one host process (several GPUs: one context per device; here two
     * contexts of device 0): every shard's records land in its rows of the caller's table - no collective, no Python */
    apda_ctx *second = NULL;
    apda_peak_rec *rec2 = (apda_peak_rec *)calloc(B, sizeof(apda_peak_rec));
    if (!rec2 || apda_ctx_create(0, &second) != APDA_OK) {
        fprintf(stderr, "second context: %s\n", apda_last_error());
        return 1;
    }
    apda_ctx *both[2] = {ctx, second};
    rc = apda_multi_analyze_f32_host(both, 2, x, N, N, B, N, APDA_CENTER_MEDIAN, 1, fs, NULL, 4, 5, rec2);
    if (rc != APDA_OK) {
        fprintf(stderr, "apda_multi_analyze_f32_host: %s\n", apda_last_error());
        return 1;
    }
    printf("multi: 2 contexts, table %s\n", memcmp(rec, rec2, sizeof(apda_peak_rec) * B) == 0 ? "identical" : "DIFFERENT");
    apda_ctx_destroy(second);
    apda_ctx_destroy(ctx);
    free(x);
    free(rec);
    free(rec2);
    return 0;
}
