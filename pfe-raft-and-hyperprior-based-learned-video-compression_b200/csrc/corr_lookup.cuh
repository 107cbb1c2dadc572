// corr_lookup.cuh -- radius-R bilinear window gather over the correlation
// pyramid (replaces TV:raft.py:394-422 index_pyramid + TV:_utils.py:8-19).
//
//   out[b, l*S*S + i*S + j, y, x] = bilinear(pyr[l][b*N + y*w + x],
//        xs = cx / 2^l + (i - R),  ys = cy / 2^l + (j - R)),   S = 2R + 1,
//   zero outside the level, pixel centres at integer coordinates.
//
// All S*S taps of one (pixel, level) share one fractional offset (the window
// offsets are integers), so the window is a separable 2-tap filter over a
// (S+1) x (S+1) footprint: every footprint element is loaded once.
//
// Mapping: one CTA = 32 consecutive query pixels x all levels; warp = level,
// lane = pixel.  A lane walks the S+1 footprint rows of ITS pixel's image at
// that level, keeps the previous horizontally-filtered row in registers and
// emits one window row per footprint row.  Because lanes are consecutive
// pixels, every output store is a fully coalesced 128-byte line of the
// (B, L*S*S, h, w) tensor -- no staging, no transpose pass.
// The gather side reads (S+1) floats from each of S+1 short rows; the `VEC`
// variant fetches them as aligned 16-byte words (4 LDG.128 instead of 10
// LDG.32 per row) and shifts in registers.
#pragma once
#include <cuda_bf16.h>
#include <cstdint>

namespace rdvc {

constexpr int LKP_MAX_LEVELS = 4;

struct LookupParams {
    const void* lvl[LKP_MAX_LEVELS];
    int hl[LKP_MAX_LEVELS];
    int wl[LKP_MAX_LEVELS];
    const float* coords;  // (B, 2, N)
    float* out;           // (B, L*S*S, N)
    int B, N;
    int num_levels;
    long long total;      // B * N
};

template <typename VolT> __device__ __forceinline__ float vol_ld(const VolT* p);
template <> __device__ __forceinline__ float vol_ld<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float vol_ld<__nv_bfloat16>(const __nv_bfloat16* p) {
    return __bfloat162float(__ldg(p));
}

// Load footprint columns [xa, xa + R2) of image row `row` (wl wide) into t[],
// zero outside [0, wl).  Scalar version: one predicated load per element.
template <int R2, typename VolT>
__device__ __forceinline__ void load_row_scalar(const VolT* __restrict__ row, int xa, int wl,
                                                float* t) {
#pragma unroll
    for (int k = 0; k < R2; ++k) {
        const int x = xa + k;
        t[k] = (x >= 0 && x < wl) ? vol_ld<VolT>(row + x) : 0.f;
    }
}

// Vector version (fp32 volume): the R2 floats starting at element `ge` of the level
// buffer are covered by NV aligned 16-byte words; a word is fetched only if it
// overlaps the valid part of the row, so no access leaves the (16-byte padded)
// level buffer.  The sub-word shift s = ge & 3 is resolved with two rounds of
// predicated register moves.
template <int R2>
__device__ __forceinline__ void load_row_vec_f32(const float* __restrict__ lvl, long long row_ge,
                                                 int xa, int wl, float* t) {
    constexpr int NV = (R2 + 3 + 3) / 4;  // words needed for any shift 0..3
    constexpr int NC = NV * 4;
    const long long ge = row_ge + xa;
    const long long al = ge & ~3LL;
    const int s = static_cast<int>(ge - al);
    const long long lo = row_ge, hi = row_ge + wl;  // valid global element range of this row
    float c[NC];
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        const long long w0 = al + 4 * k;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (w0 + 3 >= lo && w0 < hi) v = __ldg(reinterpret_cast<const float4*>(lvl + w0));
        c[4 * k] = v.x; c[4 * k + 1] = v.y; c[4 * k + 2] = v.z; c[4 * k + 3] = v.w;
    }
    if (s & 2) {
#pragma unroll
        for (int k = 0; k + 2 < NC; ++k) c[k] = c[k + 2];
    }
    if (s & 1) {
#pragma unroll
        for (int k = 0; k + 1 < NC; ++k) c[k] = c[k + 1];
    }
#pragma unroll
    for (int k = 0; k < R2; ++k) {
        const int x = xa + k;
        t[k] = (x >= 0 && x < wl) ? c[k] : 0.f;
    }
}

// grid: ceil(B*N / 32) blocks; block: 32 * num_levels threads.
template <int R, typename VolT, bool VEC>
__global__ void __launch_bounds__(32 * LKP_MAX_LEVELS)
corr_lookup_kernel(const __grid_constant__ LookupParams p) {
    constexpr int S = 2 * R + 1;
    constexpr int R2 = S + 1;  // footprint side
    const int l = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const long long pix = static_cast<long long>(blockIdx.x) * 32 + lane;
    if (pix >= p.total) return;
    const int b = static_cast<int>(pix / p.N);
    const int q = static_cast<int>(pix - static_cast<long long>(b) * p.N);

    const int hl = p.hl[l], wl = p.wl[l];
    const float inv = 1.0f / static_cast<float>(1 << l);
    float cx = __ldg(p.coords + (static_cast<size_t>(b) * 2 + 0) * p.N + q) * inv;
    float cy = __ldg(p.coords + (static_cast<size_t>(b) * 2 + 1) * p.N + q) * inv;
    // keep float->int conversion defined for wild coordinates; anything this far
    // outside samples only zeros anyway
    cx = fminf(fmaxf(cx, -1.0e6f), 1.0e6f);
    cy = fminf(fmaxf(cy, -1.0e6f), 1.0e6f);
    const float fx0 = floorf(cx), fy0 = floorf(cy);
    const float fx = cx - fx0, fy = cy - fy0;
    const int xa = static_cast<int>(fx0) - R;  // first footprint column
    const int ya = static_cast<int>(fy0) - R;  // first footprint row

    const VolT* lvl = static_cast<const VolT*>(p.lvl[l]);
    const long long img_ge = pix * (static_cast<long long>(hl) * wl);  // element offset of this image
    const size_t C_out = static_cast<size_t>(p.num_levels) * S * S;
    float* outp = p.out + (static_cast<size_t>(b) * C_out + static_cast<size_t>(l) * S * S) * p.N + q;

    // whole window left/right of the level: every tap is zero
    const bool x_dead = (xa + R2 <= 0) || (xa >= wl);

    float prev[S];
#pragma unroll
    for (int rr = 0; rr < R2; ++rr) {
        const int y = ya + rr;
        float hrow[S];
        if (y >= 0 && y < hl && !x_dead) {
            float t[R2];
            const long long row_ge = img_ge + static_cast<long long>(y) * wl;
            if constexpr (VEC) {
                load_row_vec_f32<R2>(reinterpret_cast<const float*>(lvl), row_ge, xa, wl, t);
            } else {
                load_row_scalar<R2, VolT>(lvl + row_ge, xa, wl, t);
            }
#pragma unroll
            for (int i = 0; i < S; ++i) hrow[i] = t[i] * (1.0f - fx) + t[i + 1] * fx;
        } else {
#pragma unroll
            for (int i = 0; i < S; ++i) hrow[i] = 0.f;
        }
        if (rr > 0) {
            const int j = rr - 1;  // window row: ys = cy + (j - R)
#pragma unroll
            for (int i = 0; i < S; ++i) {
                const float v = prev[i] * (1.0f - fy) + hrow[i] * fy;
                __stcs(outp + static_cast<size_t>(i * S + j) * p.N, v);
            }
        }
#pragma unroll
        for (int i = 0; i < S; ++i) prev[i] = hrow[i];
    }
}

}  // namespace rdvc
