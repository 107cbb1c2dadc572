// corr_pack.cuh -- feature-map repack, one launch for both maps:
//   (B, D, h, w) {fp32|bf16|fp16}, NCHW as torchvision's encoder emits them (TV:raft.py:492-493)
//   -> K-major bf16 rows [B][n_l][D] (D contiguous), what the tcgen05 operand descriptors in
//      corr_build_sm100.cuh expect: fmap1 at level 0; fmap2 at level 0 and, for the linear
//      build mode, at the pooled levels 1..3.
// Level l row (Y, X) is the mean over the 2^l x 2^l block of the source map (= l nested
// floor-cropped 2x2 means, TV:raft.py:390-392), taken in fp32 and rounded to bf16 ONCE.
//
// One CTA = an 8 x 32 spatial block x 64 channels, fp32 in shared memory: global reads are
// 128-byte rows of the source, global writes are 128-byte pieces (64 channels) of the K-major
// rows.  Each map is read exactly once.  HBM-bound and small next to the volume write
// (67 MB in, 39 MB out at 1080p).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cstdint>

namespace rdvc {

constexpr int PACK_THREADS = 256;
constexpr int PACK_TY = 8, PACK_TX = 32, PACK_CG = 64;   // block: 8 rows x 32 px x 64 channels
constexpr int PACK_PITCH = PACK_TY * PACK_TX + 1;        // 257 floats per channel (bank spread)
constexpr int PACK_SMEM_BYTES = PACK_CG * PACK_PITCH * 4;

struct PackParams {
    const void* src[2];          // fmap1, fmap2
    __nv_bfloat16* dst[2][4];    // [map][level]
    int levels[2];               // levels to emit per map (fmap1: 1)
    int tiled[2];                // per map: rows in RDVC_LAYOUT_TILED order instead of raster order
    int twl, thl;                // log2 tile width / height of the tiled order
    int B, D, h, w;
};

template <typename T> __device__ __forceinline__ float pack_to_float(T v);
template <> __device__ __forceinline__ float pack_to_float<float>(float v) { return v; }
template <> __device__ __forceinline__ float pack_to_float<__nv_bfloat16>(__nv_bfloat16 v) {
    return __bfloat162float(v);
}
template <> __device__ __forceinline__ float pack_to_float<__half>(__half v) {
    return __half2float(v);
}

// grid: (ceil(w/32), ceil(h/8), 2 * B * D/64); block: 256 threads; dynamic smem: PACK_SMEM_BYTES.
template <typename T>
__global__ void __launch_bounds__(PACK_THREADS)
corr_pack_kernel(const __grid_constant__ PackParams p) {
    extern __shared__ float tile[];  // [64][257], column = yy * 32 + xx
    const int D = p.D, h = p.h, w = p.w;
    const int cgn = D / PACK_CG;
    const int cg = blockIdx.z % cgn;
    const int b = (blockIdx.z / cgn) % p.B;
    const int map = blockIdx.z / (cgn * p.B);
    const int y0 = blockIdx.y * PACK_TY, x0 = blockIdx.x * PACK_TX;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const T* src = static_cast<const T*>(p.src[map]) +
                   (static_cast<size_t>(b) * D + static_cast<size_t>(cg) * PACK_CG) * h * w;

    // load: one 32-pixel source row (128 B of fp32) per warp instruction
    for (int i = warp; i < PACK_CG * PACK_TY; i += PACK_THREADS / 32) {
        const int c = i / PACK_TY, yy = i % PACK_TY;
        const int y = y0 + yy, x = x0 + lane;
        float v = 0.f;
        if (y < h && x < w) v = pack_to_float<T>(src[(static_cast<size_t>(c) * h + y) * w + x]);
        tile[c * PACK_PITCH + yy * PACK_TX + lane] = v;
    }
    __syncthreads();

    // store: one K-major row piece (64 channels = 128 B of bf16) per warp instruction
    const int L = p.levels[map];
    int row_base = 0;
    for (int l = 0; l < L; ++l) {
        const int sy = PACK_TY >> l, sx = PACK_TX >> l;   // pooled pixels in this block
        const int f = 1 << l;
        const int hl = h >> l, wl = w >> l;
        const float inv = 1.0f / static_cast<float>(f * f);
        __nv_bfloat16* dst = p.dst[map][l];
        const int nrows = sy * sx;
        // rows are dealt to warps continuing across levels so the heavy (pooled) rows spread out
        for (int r = (warp - row_base % 8 + 8) % 8; r < nrows; r += PACK_THREADS / 32) {
            const int oy = r / sx, ox = r % sx;
            const int Y = (y0 >> l) + oy, X = (x0 >> l) + ox;
            if (Y >= hl || X >= wl) continue;
            const float* t0 = tile + (2 * lane) * PACK_PITCH + (oy * f) * PACK_TX + ox * f;
            float a0 = 0.f, a1 = 0.f;
            for (int dy = 0; dy < f; ++dy)
                for (int dx = 0; dx < f; ++dx) {
                    a0 += t0[dy * PACK_TX + dx];
                    a1 += t0[PACK_PITCH + dy * PACK_TX + dx];
                }
            // operand row of pixel (Y, X): raster order, or tile by tile (see rdvc_corr.h); the
            // build's output columns follow the operand rows, so this IS the volume's layout
            size_t img = static_cast<size_t>(hl) * wl;
            size_t pix = static_cast<size_t>(Y) * wl + X;
            if (p.tiled[map]) {
                const int twl = p.twl, thl = p.thl;
                const int tiles_w = (wl + (1 << twl) - 1) >> twl, tiles_h = (hl + (1 << thl) - 1) >> thl;
                img = static_cast<size_t>(tiles_w * tiles_h) << (twl + thl);
                pix = (static_cast<size_t>((Y >> thl) * tiles_w + (X >> twl)) << (twl + thl)) +
                      ((Y & ((1 << thl) - 1)) << twl) + (X & ((1 << twl) - 1));
            }
            __nv_bfloat162* out = reinterpret_cast<__nv_bfloat162*>(
                dst + (static_cast<size_t>(b) * img + pix) * D + cg * PACK_CG);
            out[lane] = __floats2bfloat162_rn(a0 * inv, a1 * inv);
        }
        row_base += nrows;
    }
}

}  // namespace rdvc
