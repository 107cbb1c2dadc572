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

// Pooled fmap2 levels for the linear build mode: level l row (Y, X) = mean over the
// 2^l x 2^l block of the fp32 feature map (equal to l successive floor-cropped 2x2
// means), rounded ONCE to bf16, written K-major like corr_pack_kernel.
// grid: (ceil(n_l / 32), B, 1); block: 256 threads; dynamic smem: D * 33 floats.
template <typename T>
__global__ void __launch_bounds__(PACK_THREADS)
corr_pool_pack_kernel(const T* __restrict__ src, __nv_bfloat16* __restrict__ dst, int D, int h, int w,
                      int level) {
    extern __shared__ float tile[];  // [D][33]
    const int hl = h >> level, wl = w >> level, nl = hl * wl, N = h * w;
    const int f = 1 << level;
    const float inv = 1.0f / static_cast<float>(f * f);
    const int b = blockIdx.y;
    const int o0 = blockIdx.x * PACK_TN;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t plane = static_cast<size_t>(b) * D * N;
    const int o = o0 + lane;
    const int Y = o / wl, X = o - Y * wl;
    for (int d = warp; d < D; d += PACK_THREADS / 32) {
        float acc = 0.f;
        if (o < nl) {
            const T* base = src + plane + static_cast<size_t>(d) * N + static_cast<size_t>(Y * f) * w + X * f;
            for (int dy = 0; dy < f; ++dy)
                for (int dx = 0; dx < f; ++dx) acc += pack_to_float<T>(base[dy * w + dx]);
        }
        tile[d * 33 + lane] = acc * inv;
    }
    __syncthreads();
    for (int r = warp; r < PACK_TN; r += PACK_THREADS / 32) {
        const int oo = o0 + r;
        if (oo >= nl) break;
        __nv_bfloat162* out =
            reinterpret_cast<__nv_bfloat162*>(dst + (static_cast<size_t>(b) * nl + oo) * D);
        for (int d2 = lane; d2 < D / 2; d2 += 32)
            out[d2] = __floats2bfloat162_rn(tile[(2 * d2) * 33 + r], tile[(2 * d2 + 1) * 33 + r]);
    }
}

}  // namespace rdvc
