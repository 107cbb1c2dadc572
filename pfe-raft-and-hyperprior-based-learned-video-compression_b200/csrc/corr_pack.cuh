// corr_pack.cuh -- feature-map repack: (B, D, N) {fp32|bf16|fp16}, N contiguous
// (torchvision's NCHW fmaps, TV:raft.py:492-493) -> (B, N, D) bf16, D contiguous
// (K-major rows, what the tcgen05 operand descriptors in corr_build_sm100.cuh
// expect).  HBM-bound and tiny next to the volume write: 2 x 33 MB in, 2 x 17 MB
// out at 1080p.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cstdint>

namespace rdvc {

constexpr int PACK_TN = 32;       // pixels per block
constexpr int PACK_THREADS = 256;

template <typename T> __device__ __forceinline__ float pack_to_float(T v);
template <> __device__ __forceinline__ float pack_to_float<float>(float v) { return v; }
template <> __device__ __forceinline__ float pack_to_float<__nv_bfloat16>(__nv_bfloat16 v) {
    return __bfloat162float(v);
}
template <> __device__ __forceinline__ float pack_to_float<__half>(__half v) {
    return __half2float(v);
}

// grid: (ceil(N / 32), B, 2 maps); block: 256 threads; dynamic smem: D * 33 floats.
template <typename T>
__global__ void __launch_bounds__(PACK_THREADS)
corr_pack_kernel(const T* __restrict__ src1, const T* __restrict__ src2,
                 __nv_bfloat16* __restrict__ dst1, __nv_bfloat16* __restrict__ dst2, int D, int N) {
    extern __shared__ float tile[];  // [D][33]
    const T* src = (blockIdx.z == 0) ? src1 : src2;
    __nv_bfloat16* dst = (blockIdx.z == 0) ? dst1 : dst2;
    const int b = blockIdx.y;
    const int n0 = blockIdx.x * PACK_TN;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t plane = static_cast<size_t>(b) * D * N;

    // read: one channel row per warp iteration, 32 consecutive pixels per warp (128 B for fp32)
    const int n = n0 + lane;
    for (int d = warp; d < D; d += PACK_THREADS / 32) {
        float v = 0.f;
        if (n < N) v = pack_to_float<T>(src[plane + static_cast<size_t>(d) * N + n]);
        tile[d * 33 + lane] = v;
    }
    __syncthreads();
    // write: one pixel row (D bf16, contiguous) per warp iteration, 2 channels per lane per step
    for (int r = warp; r < PACK_TN; r += PACK_THREADS / 32) {
        const int nn = n0 + r;
        if (nn >= N) break;
        __nv_bfloat162* out =
            reinterpret_cast<__nv_bfloat162*>(dst + (static_cast<size_t>(b) * N + nn) * D);
        for (int d2 = lane; d2 < D / 2; d2 += 32) {
            const float lo = tile[(2 * d2) * 33 + r];
            const float hi = tile[(2 * d2 + 1) * 33 + r];
            out[d2] = __floats2bfloat162_rn(lo, hi);
        }
    }
}

// fmap2 repack WITH its pooled pyramid levels, for the linear build mode
// (corr_build_sm100.cuh): one CTA takes an 8x8 spatial block of fmap2 (all D
// channels, fp32 in smem) and emits the 64 level-0 rows, 16 level-1, 4 level-2 and
// 1 level-3 rows of the K-major bf16 operand.  Level l row (Y, X) is the mean over
// the 2^l x 2^l block of the source map (= l nested floor-cropped 2x2 means), taken
// in fp32 and rounded to bf16 ONCE.  fmap2 is read exactly once.
// grid: (ceil(w/8), ceil(h/8), B); block: 256 threads; dynamic smem: D * 65 floats.
constexpr int POOL_TS = 8;

struct PoolPackParams {
    __nv_bfloat16* dst[4];  // level l: (B, h_l * w_l, D)
    int D, h, w;
    int num_levels;
};

template <typename T>
__global__ void __launch_bounds__(PACK_THREADS)
corr_pack_pool_kernel(const T* __restrict__ src, const __grid_constant__ PoolPackParams p) {
    extern __shared__ float tile[];  // [D][65], column = py * 8 + px
    const int D = p.D, h = p.h, w = p.w, N = h * w;
    const int b = blockIdx.z;
    const int y0 = blockIdx.y * POOL_TS, x0 = blockIdx.x * POOL_TS;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t plane = static_cast<size_t>(b) * D * N;
    const int px = lane & 7, py = lane >> 3;  // lane covers (py, px) and (py + 4, px)
    for (int d = warp; d < D; d += PACK_THREADS / 32) {
        const T* base = src + plane + static_cast<size_t>(d) * N;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int yy = y0 + py + 4 * half, xx = x0 + px;
            float v = 0.f;
            if (yy < h && xx < w) v = pack_to_float<T>(base[static_cast<size_t>(yy) * w + xx]);
            tile[d * 65 + (py + 4 * half) * 8 + px] = v;
        }
    }
    __syncthreads();
    // 85 output rows: 64 (level 0) + 16 + 4 + 1; one warp per row, 2 channels per lane per step
    for (int r = warp; r < 85; r += PACK_THREADS / 32) {
        int l, idx;
        if (r < 64) { l = 0; idx = r; }
        else if (r < 80) { l = 1; idx = r - 64; }
        else if (r < 84) { l = 2; idx = r - 80; }
        else { l = 3; idx = 0; }
        if (l >= p.num_levels) break;
        const int side = POOL_TS >> l;          // pooled pixels per block side
        const int f = 1 << l;                   // source pixels per pooled pixel side
        const int oy = idx / side, ox = idx % side;
        const int Y = (y0 >> l) + oy, X = (x0 >> l) + ox;
        const int hl = h >> l, wl = w >> l;
        if (Y >= hl || X >= wl) continue;
        const float inv = 1.0f / static_cast<float>(f * f);
        __nv_bfloat162* out = reinterpret_cast<__nv_bfloat162*>(
            p.dst[l] + (static_cast<size_t>(b) * hl * wl + static_cast<size_t>(Y) * wl + X) * D);
        for (int d2 = lane; d2 < D / 2; d2 += 32) {
            const float* t0 = tile + (2 * d2) * 65 + (oy * f) * 8 + ox * f;
            float a0 = 0.f, a1 = 0.f;
            for (int dy = 0; dy < f; ++dy)
                for (int dx = 0; dx < f; ++dx) {
                    a0 += t0[dy * 8 + dx];
                    a1 += t0[65 + dy * 8 + dx];
                }
            out[d2] = __floats2bfloat162_rn(a0 * inv, a1 * inv);
        }
    }
}

}  // namespace rdvc
