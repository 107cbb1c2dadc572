/* host_pair.c -- the C ABI of librdvc_corr.so from plain C: one frame pair from HOST memory
 * (correlation volume + pyramid + `iters` lookups, copies included), no PyTorch, no CUDA headers.
 *
 *   gcc -std=c99 -Iinclude examples/host_pair.c -o host_pair \
 *       -L<pkg>/lib -lrdvc_corr -Wl,-rpath,<pkg>/lib -lm        # needs libcudart on the loader path
 *   ./host_pair [h w]                                            # default 46 80 (RDVC's 368x640 RAFT size)
 *
 * Pageable malloc() buffers work; pinned ones (cudaHostAlloc) make the copies asynchronous and are what
 * bench.py's e2e leg uses.  Two pairs can be kept in flight with rdvc_corr_pair_host_submit / _wait. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "rdvc_corr.h"

int main(int argc, char** argv) {
    const int B = 1, D = 256, L = 4, R = 4, iters = 12;
    const int h = argc > 2 ? atoi(argv[1]) : 46, w = argc > 2 ? atoi(argv[2]) : 80;
    const size_t N = (size_t)h * w, C = (size_t)L * (2 * R + 1) * (2 * R + 1);
    float* f1 = (float*)malloc(sizeof(float) * B * D * N);
    float* f2 = (float*)malloc(sizeof(float) * B * D * N);
    float* co = (float*)malloc(sizeof(float) * iters * B * 2 * N);
    float* out = (float*)malloc(sizeof(float) * iters * B * C * N);
    if (!f1 || !f2 || !co || !out) return 2;
    for (size_t i = 0; i < (size_t)B * D * N; ++i) {          /* any deterministic content */
        f1[i] = sinf(0.37f * (float)(i % 1009));
        f2[i] = cosf(0.11f * (float)(i % 2003));
    }
    for (int it = 0; it < iters; ++it)                          /* coords = pixel grid + a drifting offset */
        for (size_t q = 0; q < N; ++q) {
            co[((size_t)it * 2 + 0) * N + q] = (float)(q % w) + 0.3f * (float)it;   /* channel 0 = x */
            co[((size_t)it * 2 + 1) * N + q] = (float)(q / w) - 0.2f * (float)it;   /* channel 1 = y */
        }
    printf("librdvc_corr version %d\n", rdvc_corr_version());
    int rc = rdvc_corr_pair_host(f1, f2, co, out, B, D, h, w, L, R, iters, RDVC_DT_F32);
    if (rc != RDVC_OK) {
        fprintf(stderr, "rdvc_corr_pair_host failed (%d): %s\n", rc, rdvc_corr_last_error());
        return 1;
    }
    /* centre tap of level 0 at iteration 0 for pixel 0 is <fmap1[:, 0], fmap2[:, 0]> / sqrt(D) (bf16 operands) */
    double ref = 0.0;
    for (int c = 0; c < D; ++c) ref += (double)f1[(size_t)c * N] * (double)f2[(size_t)c * N];
    ref /= sqrt((double)D);
    printf("out[0, 4*9+4, 0, 0] = %.5f   (fp64 dot / sqrt(D) = %.5f)\n", out[(size_t)(4 * 9 + 4) * N], ref);
    rdvc_corr_release();
    free(f1); free(f2); free(co); free(out);
    return 0;
}
