/*
 * corr_oracle.c -- TEST INFRASTRUCTURE ONLY.  Not product code.
 *
 * Scalar CPU restatement of the RAFT correlation hot path that RDVC's encoder
 * executes through torchvision (the reference repo ships no CorrBlock of its
 * own).  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this file's
 * shared object.  The product path (librdvc_corr.so) never links or calls it.
 *
 * Follows, function by function (TV: = torchvision 0.26.0
 * models/optical_flow/):
 *   oracle_corr_volume     TV:raft.py:424-431  (_compute_corr_volume)
 *   oracle_build_pyramid   TV:raft.py:360-392  (build_pyramid, avg_pool2d k=2 s=2)
 *   oracle_index_pyramid   TV:raft.py:394-422  (index_pyramid)
 *                          TV:_utils.py:8-19   (grid_sample: absolute -> normalised)
 *                          aten grid_sampler_2d, bilinear / zeros / align_corners=True
 *
 * Pinning: checked against a live torchvision CorrBlock and the committed
 * fixtures in tests/golden/ by tests/test_oracle.py.  The reference's own
 * tests hold no golden vector for this path ("parity unpinned" by the
 * reference; pinned here against its dependency).
 *
 * Build: gcc -O2 -fopenmp -shared -fPIC corr_oracle.c -o _build/libcorr_oracle.so -lm
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#define ORACLE_MAX_LEVELS 8

/* Level l spatial size: floor halving at every level (avg_pool2d drops the
 * odd trailing row / column), TV:raft.py:390-392. */
static void level_dims(int h, int w, int level, int *hl, int *wl) {
    for (int l = 0; l < level; ++l) { h /= 2; w /= 2; }
    *hl = h; *wl = w;
}

size_t oracle_pyramid_elems(int B, int h, int w, int num_levels) {
    size_t total = 0;
    for (int l = 0; l < num_levels; ++l) {
        int hl, wl; level_dims(h, w, l, &hl, &wl);
        total += (size_t)B * h * w * hl * wl;
    }
    return total;
}

size_t oracle_level_offset(int B, int h, int w, int level) {
    return oracle_pyramid_elems(B, h, w, level);
}

/* corr[b, i, j] = sum_c f1[b, c, i] * f2[b, c, j] / sqrt(C)
 * f1, f2: (B, C, h*w) contiguous.  out: (B, N, N).
 * Accumulates in double, then rounds once: the tightest statement of the
 * mathematical result the fp32 GEMM approximates. */
void oracle_corr_volume(const float *f1, const float *f2, int B, int C, int h, int w,
                        float *out) {
    const size_t N = (size_t)h * w;
    const double scale = 1.0 / sqrt((double)C);
    for (int b = 0; b < B; ++b) {
        const float *a = f1 + (size_t)b * C * N;
        const float *bb = f2 + (size_t)b * C * N;
        float *o = out + (size_t)b * N * N;
#pragma omp parallel for schedule(static)
        for (long i = 0; i < (long)N; ++i) {
            for (size_t j0 = 0; j0 < N; j0 += 256) {
                double acc[256];
                size_t jn = (N - j0 < 256) ? (N - j0) : 256;
                for (size_t j = 0; j < jn; ++j) acc[j] = 0.0;
                for (int c = 0; c < C; ++c) {
                    const double av = a[(size_t)c * N + i];
                    const float *brow = bb + (size_t)c * N + j0;
                    for (size_t j = 0; j < jn; ++j) acc[j] += av * (double)brow[j];
                }
                for (size_t j = 0; j < jn; ++j) o[(size_t)i * N + j0 + j] = (float)(acc[j] * scale);
            }
        }
    }
}

/* 2x2 stride-2 average pool of `count` images (hi x wi) -> (hi/2 x wi/2). */
void oracle_avg_pool2(const float *in, size_t count, int hi, int wi, float *out) {
    const int ho = hi / 2, wo = wi / 2;
#pragma omp parallel for schedule(static)
    for (long q = 0; q < (long)count; ++q) {
        const float *src = in + (size_t)q * hi * wi;
        float *dst = out + (size_t)q * ho * wo;
        for (int y = 0; y < ho; ++y)
            for (int x = 0; x < wo; ++x) {
                /* aten avg_pool2d sums the window then divides by its size */
                float s = src[(2 * y) * wi + 2 * x] + src[(2 * y) * wi + 2 * x + 1] +
                          src[(2 * y + 1) * wi + 2 * x] + src[(2 * y + 1) * wi + 2 * x + 1];
                dst[y * wo + x] = s / 4.0f;
            }
    }
}

/* Whole pyramid, levels back to back: level l is (B*N, h_l, w_l) row-major. */
int oracle_build_pyramid(const float *f1, const float *f2, int B, int C, int h, int w,
                         int num_levels, float *pyramid) {
    if (num_levels < 1 || num_levels > ORACLE_MAX_LEVELS) return -1;
    const int min_size = 2 * (1 << (num_levels - 1)); /* TV:raft.py:376 */
    if (h < min_size || w < min_size) return -2;
    const size_t N = (size_t)h * w;
    oracle_corr_volume(f1, f2, B, C, h, w, pyramid);
    int hl = h, wl = w;
    float *cur = pyramid;
    for (int l = 1; l < num_levels; ++l) {
        float *nxt = cur + (size_t)B * N * hl * wl;
        oracle_avg_pool2(cur, (size_t)B * N, hl, wl, nxt);
        cur = nxt; hl /= 2; wl /= 2;
    }
    return 0;
}

/* One bilinear tap with zero padding, aten grid_sampler_2d semantics
 * (align_corners=True).  xn, yn are the NORMALISED coordinates produced by
 * TV:_utils.py:13-16; aten un-normalises them as ((c + 1) / 2) * (size - 1). */
static float bilinear_zeros(const float *img, int hl, int wl, float xn, float yn) {
    const float ix = ((xn + 1.0f) / 2.0f) * (float)(wl - 1);
    const float iy = ((yn + 1.0f) / 2.0f) * (float)(hl - 1);
    const float fx0 = floorf(ix), fy0 = floorf(iy);
    const long x0 = (long)fx0, y0 = (long)fy0;
    const long x1 = x0 + 1, y1 = y0 + 1;
    const float nw = ((float)x1 - ix) * ((float)y1 - iy);
    const float ne = (ix - (float)x0) * ((float)y1 - iy);
    const float sw = ((float)x1 - ix) * (iy - (float)y0);
    const float se = (ix - (float)x0) * (iy - (float)y0);
    float out = 0.0f;
    if (x0 >= 0 && x0 < wl && y0 >= 0 && y0 < hl) out += img[y0 * wl + x0] * nw;
    if (x1 >= 0 && x1 < wl && y0 >= 0 && y0 < hl) out += img[y0 * wl + x1] * ne;
    if (x0 >= 0 && x0 < wl && y1 >= 0 && y1 < hl) out += img[y1 * wl + x0] * sw;
    if (x1 >= 0 && x1 < wl && y1 >= 0 && y1 < hl) out += img[y1 * wl + x1] * se;
    return out;
}

/* coords: (B, 2, h, w), channel 0 = x, channel 1 = y (TV:_utils.py:22-26).
 * out:    (B, L*(2r+1)^2, h, w); channel = l*S*S + i*S + j, where the FIRST
 * window index i offsets x and the SECOND, j, offsets y: delta[i][j] =
 * (d[i], d[j]) from meshgrid(indexing="ij") and grid_sample reads
 * (x, y) = grid[..., 0], grid[..., 1]   (TV:raft.py:397-400, 406-411). */
int oracle_index_pyramid(const float *pyramid, const float *coords, int B, int h, int w,
                         int num_levels, int radius, float *out) {
    if (num_levels < 1 || num_levels > ORACLE_MAX_LEVELS || radius < 0) return -1;
    const size_t N = (size_t)h * w;
    const int S = 2 * radius + 1;
    const size_t C_out = (size_t)num_levels * S * S;
    for (int l = 0; l < num_levels; ++l) {
        int hl, wl; level_dims(h, w, l, &hl, &wl);
        const float *lvl = pyramid + oracle_level_offset(B, h, w, l);
#pragma omp parallel for schedule(static)
        for (long bq = 0; bq < (long)((size_t)B * N); ++bq) {
            const size_t b = (size_t)bq / N, q = (size_t)bq % N;
            float cx = coords[(b * 2 + 0) * N + q];
            float cy = coords[(b * 2 + 1) * N + q];
            /* centroids_coords = centroids_coords / 2, applied l times (TV:raft.py:412) */
            for (int k = 0; k < l; ++k) { cx = cx / 2.0f; cy = cy / 2.0f; }
            const float *img = lvl + (size_t)bq * hl * wl;
            for (int i = 0; i < S; ++i)
                for (int j = 0; j < S; ++j) {
                    const float xs = cx + (float)(i - radius);
                    const float ys = cy + (float)(j - radius);
                    const float xn = 2.0f * xs / (float)(wl - 1) - 1.0f;
                    const float yn = (hl > 1) ? (2.0f * ys / (float)(hl - 1) - 1.0f) : ys;
                    const size_t ch = (size_t)l * S * S + (size_t)i * S + j;
                    out[(b * C_out + ch) * N + q] = bilinear_zeros(img, hl, wl, xn, yn);
                }
        }
    }
    return 0;
}
