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
    int img[2][4];               // [map][level]: operand rows per batch item (pixels incl. layout padding)
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

// 4 consecutive elements (16-byte / 8-byte aligned) -> 4 floats
template <typename T> __device__ __forceinline__ void pack_ld4(const T* p, float* v);
template <> __device__ __forceinline__ void pack_ld4<float>(const float* p, float* v) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <> __device__ __forceinline__ void pack_ld4<__nv_bfloat16>(const __nv_bfloat16* p, float* v) {
    const uint2 t = __ldg(reinterpret_cast<const uint2*>(p));
    v[0] = __uint_as_float(t.x << 16); v[1] = __uint_as_float(t.x & 0xffff0000u);
    v[2] = __uint_as_float(t.y << 16); v[3] = __uint_as_float(t.y & 0xffff0000u);
}
template <> __device__ __forceinline__ void pack_ld4<__half>(const __half* p, float* v) {
    const uint2 t = __ldg(reinterpret_cast<const uint2*>(p));
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&t.x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&t.y));
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
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

    // load.  Vector path (w % 4 == 0): one warp instruction fetches 4 pixels per lane for 4
    // consecutive channels at one row (4 x 128 B of fp32), PACK_MLP of them in flight per
    // thread before the first is consumed; the 4 channels land 257 floats apart, so the four
    // scalar smem stores are conflict-free.  Scalar path: one 32-pixel row per instruction.
    constexpr int NWARP = PACK_THREADS / 32, PACK_MLP = 8;
    if ((w & 3) == 0) {
        constexpr int ROWS4 = PACK_CG * PACK_TY / 4;   // 128 four-channel row groups
        static_assert(ROWS4 % (NWARP * PACK_MLP) == 0, "load loop shape");
        const int cl = lane >> 3, xq = (lane & 7) * 4;
        const int x = x0 + xq;
        for (int i0 = warp; i0 < ROWS4; i0 += NWARP * PACK_MLP) {
            float v[PACK_MLP][4];
#pragma unroll
            for (int u = 0; u < PACK_MLP; ++u) {
                const int i = i0 + u * NWARP;
                const int c = (i / PACK_TY) * 4 + cl, y = y0 + i % PACK_TY;
                v[u][0] = v[u][1] = v[u][2] = v[u][3] = 0.f;
                if (y < h && x < w) pack_ld4<T>(src + (static_cast<size_t>(c) * h + y) * w + x, v[u]);
            }
#pragma unroll
            for (int u = 0; u < PACK_MLP; ++u) {
                const int i = i0 + u * NWARP;
                float* t = tile + ((i / PACK_TY) * 4 + cl) * PACK_PITCH + (i % PACK_TY) * PACK_TX + xq;
                t[0] = v[u][0]; t[1] = v[u][1]; t[2] = v[u][2]; t[3] = v[u][3];
            }
        }
    } else {
        static_assert((PACK_CG * PACK_TY) % (NWARP * PACK_MLP) == 0, "load loop shape");
        const int x = x0 + lane;
        for (int i0 = warp; i0 < PACK_CG * PACK_TY; i0 += NWARP * PACK_MLP) {
            float v[PACK_MLP];
#pragma unroll
            for (int u = 0; u < PACK_MLP; ++u) {
                const int i = i0 + u * NWARP;
                const int c = i / PACK_TY, y = y0 + i % PACK_TY;
                v[u] = (y < h && x < w) ? pack_to_float<T>(__ldg(src + (static_cast<size_t>(c) * h + y) * w + x)) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < PACK_MLP; ++u) {
                const int i = i0 + u * NWARP;
                tile[(i / PACK_TY) * PACK_PITCH + (i % PACK_TY) * PACK_TX + lane] = v[u];
            }
        }
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
            const size_t img = static_cast<size_t>(p.img[map][l]);
            size_t pix = static_cast<size_t>(Y) * wl + X;
            if (p.tiled[map]) {
                const int twl = p.twl, thl = p.thl;
                const int tiles_w = (wl + (1 << twl) - 1) >> twl;
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
