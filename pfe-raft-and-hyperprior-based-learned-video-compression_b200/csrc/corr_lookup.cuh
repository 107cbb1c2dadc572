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
// The gather side reads (S+1) elements from each of S+1 short rows as aligned
// 16-byte words (LDG.128) and shifts in registers.  It is DRAM-bound (no reuse
// between pixels: every query pixel owns its own image), so the pyramid layout
// decides the cost: in RDVC_LAYOUT_TILED (16-byte x 4-row tiles = one 64-byte
// DRAM atom) the footprint touches ~11 atoms per level instead of ~20 for
// row-major rows; padding pixels are zeros, so no per-element masking either.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cstdint>
#include <type_traits>

namespace rdvc {

constexpr int LKP_MAX_LEVELS = 4;

struct LookupParams {
    const void* lvl[LKP_MAX_LEVELS];
    int hl[LKP_MAX_LEVELS];
    int wl[LKP_MAX_LEVELS];
    long long img[LKP_MAX_LEVELS];   // elements of one level image (incl. layout padding)
    int tiles_w[LKP_MAX_LEVELS];     // RDVC_LAYOUT_TILED: tiles per image row
    int twl, thl;                    // RDVC_LAYOUT_TILED: log2 tile width / height
    const float* coords;  // (B, 2, N)
    void* out;            // LKP_OUT_NCHW_*: (B, L*S*S, N);  LKP_OUT_KM_*: [feat_pitch / 8][feat_rows][8] 16-bit chunks
    int feat_pitch;       // LKP_OUT_KM_*: K elements per query pixel (rdvc_corr_feat_pitch)
    long long feat_rows;  // LKP_OUT_KM_*: B*N rounded up to a multiple of 8 (rows of every chunk plane)
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
// buffer are covered by NV aligned 16-byte words; a word is fetched only if it holds a
// needed element inside the valid part of the row, so no access leaves the (16-byte
// padded) level buffer.  Split in two so the fetch of row r+1 can be issued before row r
// is consumed (software pipelining: the kernel is latency-bound, not issue-bound).
template <int R2>
struct RowWords {
    static constexpr int NV = (R2 + 3 + 3) / 4;  // words needed for any shift 0..3
    static constexpr int NC = NV * 4;
    float c[NC];
};

template <int R2>
__device__ __forceinline__ void fetch_row_vec_f32(const float* __restrict__ lvl, long long row_ge,
                                                  int xa, int wl, bool live, RowWords<R2>& rw) {
    constexpr int NV = RowWords<R2>::NV;
    const long long ge = row_ge + xa;
    const long long al = ge & ~3LL;
    const long long lo = row_ge, hi = row_ge + wl;  // valid global element range of this row
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        const long long w0 = al + 4 * k;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (live && w0 + 3 >= lo && w0 < hi && w0 < ge + R2)
            v = __ldg(reinterpret_cast<const float4*>(lvl + w0));
        rw.c[4 * k] = v.x; rw.c[4 * k + 1] = v.y; rw.c[4 * k + 2] = v.z; rw.c[4 * k + 3] = v.w;
    }
}

// The sub-word shift s = ge & 3 is resolved with two rounds of predicated register moves.
template <int R2>
__device__ __forceinline__ void select_row_vec_f32(RowWords<R2>& rw, long long row_ge, int xa, int wl,
                                                   float* t) {
    constexpr int NC = RowWords<R2>::NC;
    const int s = static_cast<int>((row_ge + xa) & 3LL);
    if (s & 2) {
#pragma unroll
        for (int k = 0; k + 2 < NC; ++k) rw.c[k] = rw.c[k + 2];
    }
    if (s & 1) {
#pragma unroll
        for (int k = 0; k + 1 < NC; ++k) rw.c[k] = rw.c[k + 1];
    }
#pragma unroll
    for (int k = 0; k < R2; ++k) {
        const int x = xa + k;
        t[k] = (x >= 0 && x < wl) ? rw.c[k] : 0.f;
    }
}

// 16-byte volume load of the tiled kernel.  RDVC_LKP_LD (experiment, compile-time): 0 = ld.global.nc (default),
// 1 / 2 = the same with an L2 prefetch-size hint of 128 / 256 bytes (pulls the neighbouring tiles of the same tile
// row into L2 with the missing one).  Measured at 1080p (tools/exp_lookup_ld.py, round 2): 128 B no change (29.7 vs
// 29.8 us fp32, 20.1 vs 20.1 bf16), 256 B 8 % slower -- bigger DRAM bursts do not buy back the extra bytes.
#ifndef RDVC_LKP_LD
#define RDVC_LKP_LD 0
#endif
__device__ __forceinline__ uint4 lkp_ld16(const void* p) {
#if RDVC_LKP_LD == 1
    uint4 v;
    asm volatile("ld.global.nc.L2::128B.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
#elif RDVC_LKP_LD == 2
    uint4 v;
    asm volatile("ld.global.nc.L2::256B.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
#else
    return __ldg(reinterpret_cast<const uint4*>(p));
#endif
}

// ---- RDVC_LAYOUT_TILED rows -------------------------------------------------
// Footprint columns [xa, xa + R2) of image row y are covered by NV aligned 16-byte words;
// word k starts at column x0 = (xa & ~(EPW-1)) + k * EPW, which never straddles a tile
// (tile_w is a multiple of EPW).  A word is fetched iff it lies inside the padded image and
// holds a needed column; everything else is zero -- and so are the padding pixels themselves.
template <int R2, typename VolT>
struct TileRow {
    static constexpr int EPW = 16 / static_cast<int>(sizeof(VolT));  // elements per word
    static constexpr int NV = (R2 + 2 * EPW - 2) / EPW;              // words for any shift
    static constexpr int NC = NV * EPW;
    uint4 wd[NV];
};

template <int R2>
__device__ __forceinline__ void tile_row_unpack(const TileRow<R2, float>& tr, float* c) {
#pragma unroll
    for (int k = 0; k < TileRow<R2, float>::NV; ++k) {
        c[4 * k] = __uint_as_float(tr.wd[k].x); c[4 * k + 1] = __uint_as_float(tr.wd[k].y);
        c[4 * k + 2] = __uint_as_float(tr.wd[k].z); c[4 * k + 3] = __uint_as_float(tr.wd[k].w);
    }
}
template <int R2>
__device__ __forceinline__ void tile_row_unpack(const TileRow<R2, __nv_bfloat16>& tr, float* c) {
#pragma unroll
    for (int k = 0; k < TileRow<R2, __nv_bfloat16>::NV; ++k) {
        const uint32_t w[4] = {tr.wd[k].x, tr.wd[k].y, tr.wd[k].z, tr.wd[k].w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            c[8 * k + 2 * i] = __uint_as_float(w[i] << 16);
            c[8 * k + 2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
        }
    }
}

// c[0..NC) -> c[s..s+R2) moved to the front, s in [0, EPW): log2(EPW) rounds of predicated moves
template <int NC, int EPW>
__device__ __forceinline__ void shift_front(float* c, int s) {
    if constexpr (EPW > 4) {
        if (s & 4) {
#pragma unroll
            for (int k = 0; k + 4 < NC; ++k) c[k] = c[k + 4];
        }
    }
    if (s & 2) {
#pragma unroll
        for (int k = 0; k + 2 < NC; ++k) c[k] = c[k + 2];
    }
    if (s & 1) {
#pragma unroll
        for (int k = 0; k + 1 < NC; ++k) c[k] = c[k + 1];
    }
}

// Per-thread setup shared by both kernels: the query pixel, its window origin and weights.
template <int R>
struct LookupSite {
    int b, q;        // batch item, pixel index inside it
    int xa, ya;      // first footprint column / row at this level
    float fx, fy;    // fractional offset (shared by every tap of the window)
    __device__ __forceinline__ void init(const LookupParams& p, long long pix, int l) {
        b = static_cast<int>(pix / p.N);
        q = static_cast<int>(pix - static_cast<long long>(b) * p.N);
        const float inv = 1.0f / static_cast<float>(1 << l);
        float cx = __ldg(p.coords + (static_cast<size_t>(b) * 2 + 0) * p.N + q) * inv;
        float cy = __ldg(p.coords + (static_cast<size_t>(b) * 2 + 1) * p.N + q) * inv;
        // keep float->int conversion defined for wild coordinates; anything this far
        // outside samples only zeros anyway
        cx = fminf(fmaxf(cx, -1.0e6f), 1.0e6f);
        cy = fminf(fmaxf(cy, -1.0e6f), 1.0e6f);
        const float fx0 = floorf(cx), fy0 = floorf(cy);
        fx = cx - fx0;
        fy = cy - fy0;
        xa = static_cast<int>(fx0) - R;
        ya = static_cast<int>(fy0) - R;
    }
};

// ---- output forms of the tiled kernel ---------------------------------------------------------
// NCHW: the (B, L*S*S, h, w) tensor torchvision's index_pyramid returns, fp32 (what RAFT.forward consumes) or
//       fp16 (what the consumer casts it to under the reference's default autocast, R:codec_processing.py:1436).
// KM  : 16-bit features as the K-major A operand of the 1x1 convolution that follows the lookup in MotionEncoder
//       (TV:raft.py:185,202; corr_conv1x1_sm100.cuh).  Feature k of query pixel m -- k = l * PL + j * S + i, PL = S*S
//       rounded up to 8, the order a thread produces its taps in (window row j outer, x index i inner); padding
//       features hold 0; the convolution's weight matrix is permuted to match on the host
//       (rdvc_conv1x1_pack_weights) -- lives at element ((k / 8) * feat_rows + m) * 8 + k % 8: CHUNK-major,
//       [K/8][rows][8].  A warp (32 consecutive pixels) then stores 8 taps of all its pixels as ONE contiguous
//       512-byte run (row-major rows, 704 bytes apart, cost 32 separate 16-byte segments per store instruction:
//       45 us per launch instead of 29), and a TMA box of 8 chunks x 128 rows lands in shared memory as exactly the
//       un-swizzled K-major core-matrix layout of tcgen05 (8 rows x 16 bytes contiguous).
constexpr int LKP_OUT_NCHW_F32 = 0, LKP_OUT_NCHW_F16 = 1, LKP_OUT_KM_BF16 = 2, LKP_OUT_KM_F16 = 3;

__host__ __device__ constexpr int lkp_level_pitch(int radius) {
    return ((2 * radius + 1) * (2 * radius + 1) + 7) / 8 * 8;
}
__host__ __device__ constexpr int lkp_feat_pitch(int num_levels, int radius) {
    return (num_levels * lkp_level_pitch(radius) + 15) / 16 * 16;      // whole UMMA k-steps
}

template <bool F16>
__device__ __forceinline__ uint4 lkp_pack8(const float* v) {
    uint4 w;
    if constexpr (F16) {
        // saturating: a correlation value beyond +-65504 (three orders above anything RAFT's features produce) must
        // not turn into an infinity that poisons the whole GEMM row
        auto sat = [](float x) { return fminf(fmaxf(x, -65504.f), 65504.f); };
        const __half2 a = __floats2half2_rn(sat(v[0]), sat(v[1])), b = __floats2half2_rn(sat(v[2]), sat(v[3]));
        const __half2 c = __floats2half2_rn(sat(v[4]), sat(v[5])), d = __floats2half2_rn(sat(v[6]), sat(v[7]));
        w.x = *reinterpret_cast<const uint32_t*>(&a); w.y = *reinterpret_cast<const uint32_t*>(&b);
        w.z = *reinterpret_cast<const uint32_t*>(&c); w.w = *reinterpret_cast<const uint32_t*>(&d);
    } else {
        const __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
        const __nv_bfloat162 c = __floats2bfloat162_rn(v[4], v[5]), d = __floats2bfloat162_rn(v[6], v[7]);
        w.x = *reinterpret_cast<const uint32_t*>(&a); w.y = *reinterpret_cast<const uint32_t*>(&b);
        w.z = *reinterpret_cast<const uint32_t*>(&c); w.w = *reinterpret_cast<const uint32_t*>(&d);
    }
    return w;
}

// ---- RDVC_LAYOUT_TILED kernel -------------------------------------------------------------
// One CTA = 32 consecutive query pixels x all levels: warp = level, lane = pixel.
// Measured at 1080p and rejected (tools/exp_lookup.py history): splitting a window's rows over
// 2-3 warps (+20 % time: the shared boundary rows cost more than the extra parallelism gives),
// 2-4 footprint rows in flight instead of 1 (no change), L2 evict_last hints on the volume loads
// (no change: the per-iteration footprint exceeds what L2 keeps), plain instead of streaming
// output stores (+15 %), a cp.async (LDGSTS) ring of 3-4 rows per thread in shared memory
// (3x SLOWER: the ring's shared memory takes the L1 capacity the gather lives on -- adjacent
// footprint rows share 32-byte sectors), L2 evict_first loads / evict_last stores in
// any combination (no change), de-phasing the CTAs with __nanosleep (CTAs delayed by up to 7 us still
// finish inside the same 28.8 us: the kernel is throughput-, not latency-bound), programmatic
// dependent launch for back-to-back lookups (fp32: 28.3 -> 33.8 us, the early CTAs take residency from the
// running kernel; bf16: no change).
// What bounds it: DRAM transactions.  In steady state (12 lookups back to back) one launch moves
// ~103 MB of 64-byte atoms in (10.6 atoms per window and level instead of the 6.25 its 400 bytes
// need) and its 42 MB result out: 145 MB in 28.5 us = 5.1 TB/s of scattered traffic, 0.78 of the
// measured copy peak.  (ncu's single cold launch shows only 13 MB written -- the rest is still
// dirty in L2 when it ends -- which is why loads-only 18.7 us + stores-only 15 us looked additive.)
// DBG (timing experiments, RDVC_EXPERIMENTS builds only): 1 = no volume loads, 2 = no output stores.
template <int R, typename VolT, int OUT, int DBG>
__global__ void __launch_bounds__(32 * LKP_MAX_LEVELS)
corr_lookup_tiled_kernel(const __grid_constant__ LookupParams p) {
    constexpr int S = 2 * R + 1;
    constexpr int R2 = S + 1;                       // footprint side
    using TRow = TileRow<R2, VolT>;
    constexpr int EPW = TRow::EPW, NV = TRow::NV, NC = TRow::NC;
    const int l = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const long long pix = static_cast<long long>(blockIdx.x) * 32 + lane;
    if (pix >= p.total) return;
    LookupSite<R> site;
    site.init(p, pix, l);
    const float fx = site.fx, fy = site.fy;
    const int xa = site.xa, ya = site.ya;

    const int hl = p.hl[l];
    const int twl = p.twl, thl = p.thl, tiles_w = p.tiles_w[l];
    const int wpad = tiles_w << twl;
    const int x_al = xa & ~(EPW - 1), s = xa & (EPW - 1);
    const VolT* img = static_cast<const VolT*>(p.lvl[l]) + pix * p.img[l];   // this pixel's image
    // element offsets inside the image fit 32 bits: one IMAD.WIDE per address
    int colpart[NV];
    bool okx[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        const int x0 = x_al + k * EPW;
        okx[k] = (x0 >= 0) && (x0 < wpad) && (x0 < xa + R2) && (DBG != 1);
        colpart[k] = ((x0 >> twl) << (twl + thl)) + (x0 & ((1 << twl) - 1));
    }
    auto fetch = [&](int y, TRow& tr) {
        const bool live = (y >= 0) && (y < hl);
        const int rowpart = (((y >> thl) * tiles_w) << (twl + thl)) + ((y & ((1 << thl) - 1)) << twl);
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if (live && okx[k]) v = lkp_ld16(img + (rowpart + colpart[k]));
            tr.wd[k] = v;
        }
    };

    constexpr bool KM = (OUT == LKP_OUT_KM_BF16 || OUT == LKP_OUT_KM_F16);
    using OutT = typename std::conditional<OUT == LKP_OUT_NCHW_F32, float, __half>::type;   // NCHW element type
    const size_t C_out = static_cast<size_t>(p.num_levels) * S * S;
    OutT* outp = static_cast<OutT*>(p.out) + (static_cast<size_t>(site.b) * C_out + static_cast<size_t>(l) * S * S) * p.N + site.q;
    const size_t chan_i = static_cast<size_t>(S) * p.N;   // stride of the window's x index
    // K-major form: this thread's S*S taps are chunks l * PL/8 .. of row `pix`; a chunk plane is feat_rows * 16 bytes
    constexpr int PL = lkp_level_pitch(R);
    const size_t chunk_stride = static_cast<size_t>(p.feat_rows) * 8;                       // elements
    uint16_t* featp = static_cast<uint16_t*>(p.out) + static_cast<size_t>(l) * (PL / 8) * chunk_stride +
                      static_cast<size_t>(pix) * 8;
    float stash[8];

    TRow buf[2];
    fetch(ya, buf[0]);
    float prev[S];
#pragma unroll
    for (int rr = 0; rr < R2; ++rr) {
        if (rr + 1 < R2) fetch(ya + rr + 1, buf[(rr + 1) & 1]);   // next row in flight first
        float c[NC];
        tile_row_unpack<R2>(buf[rr & 1], c);
        shift_front<NC, EPW>(c, s);
        float hrow[S];
#pragma unroll
        for (int i = 0; i < S; ++i) hrow[i] = c[i] * (1.0f - fx) + c[i + 1] * fx;
        if (rr > 0) {                               // window row j = rr - 1: ys = cy + (j - R)
            OutT* o = outp + static_cast<size_t>(rr - 1) * p.N;
#pragma unroll
            for (int i = 0; i < S; ++i) {
                const float v = prev[i] * (1.0f - fy) + hrow[i] * fy;
                if constexpr (KM) {
                    const int idx = (rr - 1) * S + i;          // a compile-time constant once unrolled
                    stash[idx & 7] = v;
                    if ((idx & 7) == 7 || idx == S * S - 1) {
#pragma unroll
                        for (int z = (idx & 7) + 1; z < 8; ++z) stash[z] = 0.f;    // padding columns of the level
                        const uint4 wv = lkp_pack8<OUT == LKP_OUT_KM_F16>(stash);
                        if (DBG != 2 || v == 12345.678f) *reinterpret_cast<uint4*>(featp + (idx >> 3) * chunk_stride) = wv;
                    }
                } else if constexpr (OUT == LKP_OUT_NCHW_F16) {
                    if (DBG != 2 || v == 12345.678f) *o = __float2half_rn(v);
                    o += chan_i;
                } else {
                    if (DBG != 2 || v == 12345.678f) __stcs(o, v);
                    o += chan_i;
                }
            }
        }
#pragma unroll
        for (int i = 0; i < S; ++i) prev[i] = hrow[i];
    }
    if constexpr (KM) {
        if (l == p.num_levels - 1)                       // feature-row tail: whole zero chunks up to feat_pitch
            for (int c = PL; l * PL + c < p.feat_pitch; c += 8)
                *reinterpret_cast<uint4*>(featp + (c >> 3) * chunk_stride) = make_uint4(0u, 0u, 0u, 0u);
    }
}

// ---- RDVC_LAYOUT_ROWMAJOR kernel (torchvision's element order; kept for interchange) -------
// VEC: 128-bit loads (fp32 volume only) instead of one predicated load per element.
// grid: ceil(B*N / 32) blocks; block: 32 * num_levels threads (warp = level, lane = pixel).
template <int R, typename VolT, bool VEC>
__global__ void __launch_bounds__(32 * LKP_MAX_LEVELS)
corr_lookup_kernel(const __grid_constant__ LookupParams p) {
    constexpr int S = 2 * R + 1;
    constexpr int R2 = S + 1;  // footprint side
    const int l = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const long long pix = static_cast<long long>(blockIdx.x) * 32 + lane;
    if (pix >= p.total) return;
    LookupSite<R> site;
    site.init(p, pix, l);
    const float fx = site.fx, fy = site.fy;
    const int xa = site.xa, ya = site.ya;
    const int hl = p.hl[l], wl = p.wl[l];

    const VolT* lvl = static_cast<const VolT*>(p.lvl[l]);
    const long long img_ge = pix * p.img[l];  // element offset of this pixel's image
    const size_t C_out = static_cast<size_t>(p.num_levels) * S * S;
    float* outp = static_cast<float*>(p.out) + (static_cast<size_t>(site.b) * C_out + static_cast<size_t>(l) * S * S) * p.N + site.q;

    // whole window left/right of the level: every tap is zero
    const bool x_dead = (xa + R2 <= 0) || (xa >= wl);

    float prev[S];
    if constexpr (VEC) {
        const float* lv32 = reinterpret_cast<const float*>(lvl);
        RowWords<R2> buf[2];
        fetch_row_vec_f32<R2>(lv32, img_ge + static_cast<long long>(ya) * wl, xa, wl,
                              ya >= 0 && ya < hl && !x_dead, buf[0]);
#pragma unroll
        for (int rr = 0; rr < R2; ++rr) {
            const int y = ya + rr;
            if (rr + 1 < R2) {  // issue the next row's loads before consuming this row
                const int yn = y + 1;
                fetch_row_vec_f32<R2>(lv32, img_ge + static_cast<long long>(yn) * wl, xa, wl,
                                      yn >= 0 && yn < hl && !x_dead, buf[(rr + 1) & 1]);
            }
            float hrow[S];
            if (y >= 0 && y < hl && !x_dead) {
                float t[R2];
                select_row_vec_f32<R2>(buf[rr & 1], img_ge + static_cast<long long>(y) * wl, xa, wl, t);
#pragma unroll
                for (int i = 0; i < S; ++i) hrow[i] = t[i] * (1.0f - fx) + t[i + 1] * fx;
            } else {
#pragma unroll
                for (int i = 0; i < S; ++i) hrow[i] = 0.f;
            }
            if (rr > 0) {
                const int j = rr - 1;  // window row: ys = cy + (j - R)
#pragma unroll
                for (int i = 0; i < S; ++i)
                    __stcs(outp + static_cast<size_t>(i * S + j) * p.N, prev[i] * (1.0f - fy) + hrow[i] * fy);
            }
#pragma unroll
            for (int i = 0; i < S; ++i) prev[i] = hrow[i];
        }
    } else {
#pragma unroll
        for (int rr = 0; rr < R2; ++rr) {
            const int y = ya + rr;
            float hrow[S];
            if (y >= 0 && y < hl && !x_dead) {
                float t[R2];
                const long long row_ge = img_ge + static_cast<long long>(y) * wl;
                load_row_scalar<R2, VolT>(lvl + row_ge, xa, wl, t);
#pragma unroll
                for (int i = 0; i < S; ++i) hrow[i] = t[i] * (1.0f - fx) + t[i + 1] * fx;
            } else {
#pragma unroll
                for (int i = 0; i < S; ++i) hrow[i] = 0.f;
            }
            if (rr > 0) {
                const int j = rr - 1;  // window row: ys = cy + (j - R)
#pragma unroll
                for (int i = 0; i < S; ++i)
                    __stcs(outp + static_cast<size_t>(i * S + j) * p.N, prev[i] * (1.0f - fy) + hrow[i] * fy);
            }
#pragma unroll
            for (int i = 0; i < S; ++i) prev[i] = hrow[i];
        }
    }
}

}  // namespace rdvc
