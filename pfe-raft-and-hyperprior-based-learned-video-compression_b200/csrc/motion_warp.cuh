// motion_warp.cuh -- the two steps between RAFT and the motion-compensation network of RDVC's
// P-frame path, fused ("next" row f-4 of SURVEY.md section 8):
//
//   resize_flow   R:codec_processing.py:772-818   bilinear resize of the RAFT-resolution flow to the
//                 frame size (aten upsample_bilinear2d, align_corners=False, antialias=False), then
//                 dx *= W / w_in, dy *= H / h_in                      (called at :1446)
//   WarpingLayer  R:codec_processing.py:322-367   warped[b,c,i,j] = bilinear(prev[b,c]; i + dy, j + dx),
//                 sample point clamped to the image (padding_mode='border'), pixel centres at
//                 integers (align_corners=True)                        (called at :1456)
//
// The reference runs ~15 elementwise / index kernels for this (resize, zeros_like + two slice copies,
// linspace x2, meshgrid, stack, repeat, two divides, stack, add, grid_sample) and round-trips the
// frame-resolution flow and a (B,H,W,2) sampling grid through HBM.  Here one thread produces one
// output pixel: it interpolates its own flow vector from the RAFT-resolution field (4 taps), writes
// it (the caller needs it for the MCN and the motion codec) and gathers the C channels of the
// previous frame around (i + dy, j + dx).  Bound by HBM traffic in principle: algorithmic bytes per frame =
// 4 * (2 h w + C H W [prev] + C H W [warped] + 2 H W [flow out]).
#pragma once
#include <cstdint>

namespace rdvc {

struct WarpParams {
    const float* prev;     // (B, C, H, W) or nullptr (resize only)
    const float* flow;     // (B, 2, h_in, w_in)
    float* warped;         // (B, C, H, W) or nullptr
    float* flow_out;       // (B, 2, H, W) or nullptr (do not materialise the resized flow)
    int B, C, H, W, h_in, w_in;
    float ry, rx;          // h_in / H, w_in / W  (aten's area_pixel_compute_scale, fp32)
    float sh, sw;          // H / h_in, W / w_in  (the flow-value scale factors, :808-809)
    int same_size;         // (h_in, w_in) == (H, W): resize_flow returns its input (:788)
};

// aten area_pixel_compute_source_index, align_corners=False, non-cubic: clamp below at 0
__device__ __forceinline__ void warp_src_index(int dst, float ratio, int n_in, int& i0, int& i1, float& lam) {
    float src = ratio * (static_cast<float>(dst) + 0.5f) - 0.5f;
    src = src < 0.f ? 0.f : src;
    i0 = min(static_cast<int>(src), n_in - 1);
    i1 = min(i0 + 1, n_in - 1);
    lam = src - static_cast<float>(i0);
}

constexpr int WARP_THREADS = 128;
constexpr int WARP_CB = 3;    // channels whose taps are in flight together (a frame has 3)
constexpr int WARP_PIX = 2;   // output pixels per thread, 128 apart (lanes stay on consecutive pixels, so every
                              // load and store instruction is coalesced): 4 independent flow -> taps chains
                              // in flight, because the kernel is latency-, not bandwidth-bound

// grid: (ceil(W / (128 * 4)), H, B)
__global__ void __launch_bounds__(WARP_THREADS)
motion_warp_kernel(const __grid_constant__ WarpParams p) {
    const int j0 = blockIdx.x * (WARP_THREADS * WARP_PIX) + threadIdx.x;   // pixel u of this thread: j0 + 128 u
    const int i = blockIdx.y, b = blockIdx.z;
    if (j0 >= p.W) return;
    constexpr int JS = WARP_THREADS;
    const size_t in_plane = static_cast<size_t>(p.h_in) * p.w_in;
    const float* fx_in = p.flow + static_cast<size_t>(b) * 2 * in_plane;
    const float* fy_in = fx_in + in_plane;
    const size_t plane = static_cast<size_t>(p.H) * p.W;
    const size_t pix = static_cast<size_t>(i) * p.W + j0;
    float dx[WARP_PIX], dy[WARP_PIX];
    if (p.same_size) {
#pragma unroll
        for (int u = 0; u < WARP_PIX; ++u) {
            const int j = min(j0 + u * JS, p.W - 1);
            dx[u] = __ldg(fx_in + static_cast<size_t>(i) * p.w_in + j);
            dy[u] = __ldg(fy_in + static_cast<size_t>(i) * p.w_in + j);
        }
    } else {
        int y0, y1;
        float ly;
        warp_src_index(i, p.ry, p.h_in, y0, y1, ly);
        const float hy = 1.f - ly;
        const size_t r0 = static_cast<size_t>(y0) * p.w_in, r1 = static_cast<size_t>(y1) * p.w_in;
        float a[WARP_PIX][8], lx[WARP_PIX];
#pragma unroll
        for (int u = 0; u < WARP_PIX; ++u) {    // all 32 taps in flight before the first is used
            int x0, x1;
            warp_src_index(min(j0 + u * JS, p.W - 1), p.rx, p.w_in, x0, x1, lx[u]);
            a[u][0] = __ldg(fx_in + r0 + x0); a[u][1] = __ldg(fx_in + r0 + x1);
            a[u][2] = __ldg(fx_in + r1 + x0); a[u][3] = __ldg(fx_in + r1 + x1);
            a[u][4] = __ldg(fy_in + r0 + x0); a[u][5] = __ldg(fy_in + r0 + x1);
            a[u][6] = __ldg(fy_in + r1 + x0); a[u][7] = __ldg(fy_in + r1 + x1);
        }
#pragma unroll
        for (int u = 0; u < WARP_PIX; ++u) {
            const float hx = 1.f - lx[u];
            // same association as aten's upsample_bilinear2d kernel
            dx[u] = (hy * (hx * a[u][0] + lx[u] * a[u][1]) + ly * (hx * a[u][2] + lx[u] * a[u][3])) * p.sw;
            dy[u] = (hy * (hx * a[u][4] + lx[u] * a[u][5]) + ly * (hx * a[u][6] + lx[u] * a[u][7])) * p.sh;
        }
    }
    if (p.flow_out) {
        float* fo = p.flow_out + static_cast<size_t>(b) * 2 * plane + pix;
#pragma unroll
        for (int u = 0; u < WARP_PIX; ++u)
            if (j0 + u * JS < p.W) { __stcs(fo + u * JS, dx[u]); __stcs(fo + plane + u * JS, dy[u]); }
    }
    if (!p.warped) return;
    // grid_sample(bilinear, border, align_corners=True) at absolute coordinates (j + dx, i + dy)
    float w[WARP_PIX][4];
    int o[WARP_PIX][4];
#pragma unroll
    for (int u = 0; u < WARP_PIX; ++u) {
        const float sx = fminf(fmaxf(static_cast<float>(j0 + u * JS) + dx[u], 0.f), static_cast<float>(p.W - 1));
        const float sy = fminf(fmaxf(static_cast<float>(i) + dy[u], 0.f), static_cast<float>(p.H - 1));
        const float fx0 = floorf(sx), fy0 = floorf(sy);
        const float ax = sx - fx0, ay = sy - fy0;
        const int xa = static_cast<int>(fx0), ya = static_cast<int>(fy0);
        const int xb = min(xa + 1, p.W - 1), yb = min(ya + 1, p.H - 1);   // weight is 0 when clamped
        w[u][0] = (1.f - ax) * (1.f - ay); w[u][1] = ax * (1.f - ay); w[u][2] = (1.f - ax) * ay; w[u][3] = ax * ay;
        o[u][0] = ya * p.W + xa; o[u][1] = ya * p.W + xb; o[u][2] = yb * p.W + xa; o[u][3] = yb * p.W + xb;
    }
    const float* src = p.prev + static_cast<size_t>(b) * p.C * plane;
    float* dst = p.warped + static_cast<size_t>(b) * p.C * plane + pix;
    // channels in batches of WARP_CB: all taps of a batch are in flight before the first is used
    for (int c0 = 0; c0 < p.C; c0 += WARP_CB) {
        float t[WARP_CB][WARP_PIX][4];
#pragma unroll
        for (int k = 0; k < WARP_CB; ++k) {
            const float* s = src + static_cast<size_t>(min(c0 + k, p.C - 1)) * plane;
#pragma unroll
            for (int u = 0; u < WARP_PIX; ++u) {
                t[k][u][0] = __ldg(s + o[u][0]); t[k][u][1] = __ldg(s + o[u][1]);
                t[k][u][2] = __ldg(s + o[u][2]); t[k][u][3] = __ldg(s + o[u][3]);
            }
        }
#pragma unroll
        for (int k = 0; k < WARP_CB; ++k) {
            if (c0 + k >= p.C) break;
            float* d = dst + static_cast<size_t>(c0 + k) * plane;
#pragma unroll
            for (int u = 0; u < WARP_PIX; ++u)
                if (j0 + u * JS < p.W)
                    __stcs(d + u * JS, t[k][u][0] * w[u][0] + t[k][u][1] * w[u][1] + t[k][u][2] * w[u][2] + t[k][u][3] * w[u][3]);
        }
    }
}

}  // namespace rdvc
