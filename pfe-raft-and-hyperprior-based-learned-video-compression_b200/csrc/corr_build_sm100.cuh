// corr_build_sm100.cuh -- all-pairs correlation volume + 4-level pyramid in ONE
// pass (replaces TV:raft.py:360-392 build_pyramid and :424-431
// _compute_corr_volume).
//
//   pyr[0][b*N + i][y][x] = (1/sqrt(D)) * sum_c fmap1[b,c,i] * fmap2[b,c,y*w+x]
//   pyr[l+1]              = 2x2 mean of pyr[l] over (y, x), floor-cropped
//
// Shape of the computation: a [N x D] x [D x n_l] GEMM per level and batch item with
// D = 256 (only 16 UMMA k-steps) whose fp32 output (5.7 GB at 1080p) dwarfs its inputs
// (2 x 16.7 MB bf16).  It is HBM-WRITE bound (0.87 ms at the measured 6.5 TB/s vs 0.43 ms
// of tensor time), so the kernel is organised around the epilogue:
//
//  * Persistent, one CTA per SM, warp-specialised:
//      warp 0      TMA producer   (elected lane)
//      warp 1      tcgen05.mma issuer (elected lane), M128 x N256 x K16, bf16 -> fp32
//      warp 2      TMEM allocator (512 columns = two 128x256 fp32 accumulators)
//      warp 3      idle
//      warps 4+    epilogue (4 or 8 warps, BuildCfg): TMEM -> registers -> scale (-> bf16) ->
//                  swizzled smem staging -> TMA tensor stores
//  * fmap2-STATIONARY: one N-tile (256 fmap2 pixels x full K = 128 KB of smem) stays
//    resident while 128-row fmap1 tiles stream through a 3- or 4-stage 16 KB ring, so
//    L2->SM operand traffic is half the output bytes instead of equal.
//  * Two accumulators ping-pong so the MMAs of tile t+1 run under the epilogue of tile t.
//  * Two ways to get the pyramid (MODE below): LINEAR (default) computes every level as
//    its own GEMM columns from pooled fmap2 rows, so all stores are full-width boxes;
//    FUSED pools levels 1-3 in the epilogue's registers from spatial fmap2 tiles (exact
//    pooling of the fp32 accumulators, but sub-line writes: 2.3x slower, kept for parity).
//  * LINEAR-mode stores are 4 KB TMA boxes of 16 query rows x 256 contiguous bytes (wide
//    path; 32 rows x 128 bytes or staged st.global for row pitches that do not allow it).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cstdint>

#include "ptx_sm100.cuh"

namespace rdvc {

constexpr int BLD_BLOCK_M = 128;
constexpr int BLD_BLOCK_N = 256;
constexpr int BLD_BLOCK_K = 64;  // bf16 per 128-byte swizzle row
constexpr int BLD_UMMA_K = 16;
#ifndef RDVC_EW4_A_STAGES      // experiment knobs for the 4-warp epilogue shape (see BuildCfg)
#define RDVC_EW4_A_STAGES 4
#endif
#ifndef RDVC_EW4_STG_BUFS
#define RDVC_EW4_STG_BUFS 2
#endif
// Round 2, 8-warp shape at 1080p (tools/exp_epi_shapes.py): 3 stages + 1 buffer fp32 0.958 / bf16 0.710 ms; 4 + 1:
// 1.011 / 0.650; 2 + 2: 1.187 / 0.990 (the ring starves) -- against 0.982 / 0.632 for the 4-warp shape (4 + 2).  So:
// fp32 volume -> 8 warps, 3 + 1; bf16 volume -> 4 warps, 4 + 2, as before.
#ifndef RDVC_EW8_A_STAGES      // the same knobs for the 8-warp shape (3 stages + 1 buffer by default)
#define RDVC_EW8_A_STAGES 3
#endif
#ifndef RDVC_EW8_STG_BUFS
#define RDVC_EW8_STG_BUFS 1
#endif
constexpr int BLD_MAX_KC = 4;    // D <= 256
constexpr int BLD_A_STAGE_BYTES = BLD_BLOCK_M * BLD_BLOCK_K * 2;  // 16 KB
constexpr int BLD_B_SLAB_BYTES = BLD_BLOCK_N * BLD_BLOCK_K * 2;   // 32 KB
constexpr int BLD_STG_BYTES = 4096;  // one TMA-store staging buffer (a 4 KB box)
constexpr int BLD_MAX_LEVELS = 4;

constexpr int BLD_SMEM_B = 0;
constexpr int BLD_SMEM_A = BLD_SMEM_B + BLD_MAX_KC * BLD_B_SLAB_BYTES;       // 131072

// Epilogue shape.  EW = 8 epilogue warps: two per TMEM lane quarter, each half a tile's width, one
// 4 KB staging buffer per warp (32 KB in flight) beside a 3-stage fmap1 ring.  EW = 4: one warp per
// quarter, the whole tile width, two buffers per warp (32 KB in flight) beside a 4-stage ring.  The
// stationary fmap2 tile takes 128 KB, so ring + staging share ~97 KB (2 ring stages starve the MMA:
// 1.18 ms).  Measured at 1080p: fp32 volume 0.98 ms (EW 8) vs 1.02-1.05 ms (EW 4: one warp per
// quarter cannot issue the boxes fast enough); bf16 volume 0.75 ms (EW 8: its two boxes per tile
// serialise on the single buffer, ~0.8 us each) vs 0.66 ms (EW 4 with 4 + 2; 3 + 3 gives 0.68, 4 + 1
// 0.675).  With the stores compiled out either shape runs in 0.50 ms (fmap1-load latency, not MMA
// rate), so the bf16 volume (0.46 ms of writes) sits between its two bounds.
template <int EW>
struct BuildCfg {
    static_assert(EW == 4 || EW == 8, "epilogue warps");
    static constexpr int EPI_WARPS = EW;
    static constexpr int SUBS = 8 / EW;                 // 128-column halves of a tile per epilogue warp
    static constexpr int STG_BUFS = (EW == 4) ? RDVC_EW4_STG_BUFS : RDVC_EW8_STG_BUFS;   // staging buffers per epilogue warp
    static constexpr int A_STAGES = (EW == 4) ? RDVC_EW4_A_STAGES : RDVC_EW8_A_STAGES;   // fmap1 ring stages (16 KB each)
    static constexpr int THREADS = 128 + EW * 32;
    static constexpr int SMEM_STG = BLD_SMEM_A + A_STAGES * BLD_A_STAGE_BYTES;
    static constexpr int SMEM_BAR = SMEM_STG + EW * BLD_STG_BYTES * STG_BUFS;
    static constexpr int SMEM_TOTAL = SMEM_BAR + 128;
    static constexpr int SMEM_LAUNCH = SMEM_TOTAL + 1024;  // slack for 1024-byte alignment
    static_assert(SMEM_LAUNCH <= 232448, "shared memory: 227 KB per CTA");
};

struct BuildParams {
    void* lvl[BLD_MAX_LEVELS];  // level base pointers (256-byte aligned)
    int hl[BLD_MAX_LEVELS];
    int wl[BLD_MAX_LEVELS];
    int nl[BLD_MAX_LEVELS];     // MODE_LINEAR: output columns (pixels incl. layout padding) of level l
    int B, h, w, N;
    int num_levels;
    int kc;             // D / 64
    int m_blks;         // ceil(N / 128)
    int nty, ntx;       // MODE_FUSED: fmap2 tiles along y / x
    int ntiles;         // N-tiles per batch item (fused: nty * ntx; linear: sum over levels)
    int tile_start[BLD_MAX_LEVELS];  // MODE_LINEAR: first N-tile of each level
    int msplit;         // m-range slices per fmap2 tile (work item = tile x slice)
    float scale;        // 1 / sqrt(D)
    int ab_format;      // tcgen05 kind::f16 operand format: 1 = bf16, 0 = fp16
    int dbg_store_mask; // RDVC_EXPERIMENTS builds only: bit l set = write level l (default 15)
    int dbg_policy;     // RDVC_EXPERIMENTS builds only: TMA-store L2 policy (0 default, 1 evict_last, 2 evict_first)
    int tma_out;        // MODE_LINEAR: 2 bits per level: how level l is written --
                        //   0 staged st.global, 1 = TMA boxes 32 rows x 128 B (3-D map {n_l, N, B}),
                        //   2 = TMA boxes 16 rows x 256 B (4-D map {128 B, n_l*es/128, N, B})
};

// ---- output element traits -------------------------------------------------
template <typename OutT> struct OutTraits;
template <> struct OutTraits<float> {
    static constexpr int EPC = 4;  // elements per 16-byte chunk
    static constexpr int CH = 8;   // chunks per 32-element staging row
    __device__ static __forceinline__ int swz(int c, int r) { return c ^ (r & 7); }
    __device__ static __forceinline__ uint4 pack(const float* v) {
        return make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]),
                          __float_as_uint(v[3]));
    }
    __device__ static __forceinline__ float cvt(float x) { return x; }
};
template <> struct OutTraits<__nv_bfloat16> {
    static constexpr int EPC = 8;
    static constexpr int CH = 4;
    __device__ static __forceinline__ int swz(int c, int r) { return c ^ ((r >> 1) & 3); }
    __device__ static __forceinline__ uint32_t pk(float a, float b) {
        __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
        return *reinterpret_cast<uint32_t*>(&t);
    }
    __device__ static __forceinline__ uint4 pack(const float* v) {
        return make_uint4(pk(v[0], v[1]), pk(v[2], v[3]), pk(v[4], v[5]), pk(v[6], v[7]));
    }
    __device__ static __forceinline__ __nv_bfloat16 cvt(float x) { return __float2bfloat16_rn(x); }
};

__device__ __forceinline__ void st_stream_16(void* p, uint4 v) {
    asm volatile("st.global.cs.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y),
                 "r"(v.z), "r"(v.w)
                 : "memory");
}

__device__ __forceinline__ void sts_16(uint32_t addr, uint4 v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z),
                 "r"(v.w)
                 : "memory");
}
__device__ __forceinline__ uint4 lds_16(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "r"(addr)
                 : "memory");
    return v;
}
template <typename OutT> __device__ __forceinline__ OutT lds_elem(uint32_t addr);
template <> __device__ __forceinline__ float lds_elem<float>(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
    return v;
}
template <> __device__ __forceinline__ __nv_bfloat16 lds_elem<__nv_bfloat16>(uint32_t addr) {
    uint16_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr) : "memory");
    return __ushort_as_bfloat16(v);
}

// One warp writes a [32 rows][32 elements] register block (lane = row) to global
// memory through its private swizzled staging buffer (`stg`: 32-bit shared
// address) so that global stores are 16-byte vectors over contiguous segments.
// The 32 elements of a row are 32/S segments of S elements: segment s goes to
// level row (yl0 + s), columns [xl0, xl0 + S).  `img0` points at the image of the
// warp's first row; rows >= rows_valid and pixels outside (hl, wl) are skipped.
template <typename OutT, int S>
__device__ __forceinline__ void staged_store(uint32_t stg, const float* v, int lane, OutT* img0,
                                             size_t img_elems, int rows_valid, int yl0, int xl0,
                                             int hl, int wl, bool vec) {
    using TR = OutTraits<OutT>;
    constexpr int EPC = TR::EPC, CH = TR::CH;
    constexpr int RPI = 32 / CH;  // rows covered per read-back iteration
#pragma unroll
    for (int c = 0; c < CH; ++c)
        sts_16(stg + (lane * CH + TR::swz(c, lane)) * 16, TR::pack(v + c * EPC));
    __syncwarp();
    if (vec) {
        const int c = lane % CH, r0 = lane / CH;
        const int e0 = c * EPC;
        const int y = yl0 + e0 / S, x = xl0 + e0 % S;
        const bool ok = (y < hl) && (x < wl);
        OutT* dst = img0 + static_cast<size_t>(r0) * img_elems + static_cast<size_t>(y) * wl + x;
        uint4 val[CH];
#pragma unroll
        for (int it = 0; it < CH; ++it) {
            const int r = it * RPI + r0;
            val[it] = lds_16(stg + (r * CH + TR::swz(c, r)) * 16);
        }
#pragma unroll
        for (int it = 0; it < CH; ++it) {
            const int r = it * RPI + r0;
            if (ok && r < rows_valid) st_stream_16(dst + static_cast<size_t>(it * RPI) * img_elems, val[it]);
        }
    } else {
        const int c = lane / EPC, e = lane % EPC;
        const int y = yl0 + lane / S, x = xl0 + lane % S;
        const bool ok = (y < hl) && (x < wl);
        OutT* dst = img0 + static_cast<size_t>(y) * wl + x;
        for (int r = 0; r < rows_valid; ++r) {
            const OutT val = lds_elem<OutT>(stg + ((r * CH + TR::swz(c, r)) * EPC + e) * sizeof(OutT));
            if (ok) dst[static_cast<size_t>(r) * img_elems] = val;
        }
    }
    __syncwarp();
}

// MODE_FUSED : an N-tile is a TILE_Y x TILE_X spatial block of fmap2 (4-D TMA box); the
//              epilogue pools levels 1..3 in registers (one thread = one query pixel = one
//              2-D patch) and writes every level.  Elegant, exact pooling of the fp32
//              accumulators -- but a CTA then owns only 64/32/16/8-byte pieces of each
//              output line and HBM punishes sub-line writes (measured: +1.3 ms at 1080p).
// MODE_LINEAR: pooling is linear, so level l = fmap1^T . avgpool_l(fmap2) / sqrt(D) exactly
//              (SURVEY.md 7.3).  The pack pass emits the pooled fmap2 levels as extra K-major
//              rows; an N-tile is 256 CONSECUTIVE pixels of one level, so every thread's
//              output row is 1 KB contiguous and every store is a full, aligned line.
//              +33 % tensor work (0.43 ms at peak, still under the 0.87 ms HBM bound).
constexpr int MODE_FUSED = 0;
constexpr int MODE_LINEAR = 1;

// CL = 2 (MODE_LINEAR only; launched as clusters of two CTAs): the two CTAs of a cluster hold fmap2 tiles 2j and 2j + 1
// of the SAME batch item and m-slice, so they stream the same fmap1 tiles in the same order -- each loads HALF of
// every 16 KB ring stage (64 of the 128 query rows) and TMA-multicasts it into both CTAs' shared memory: one L2 read
// feeds two SMs.  Why: at full MMA rate the fmap1 stream alone is 61 GB/s per SM = 9.1 TB/s chip-wide, above the
// ~7 TB/s L2 -> SM feed this part sustains, while the volume's writes cross the same L2.
// Protocol: A_FULL[s] stays per CTA (armed by the CTA's own producer for the full 16 KB; the peer's half arrives on
// it through the multicast's mbarrier signal); A_EMPTY[s] takes TWO arrivals per phase -- each CTA's tcgen05.commit is
// multicast to both CTAs -- because a stage may be overwritten by the peer's next multicast only when BOTH consumers
// are done with it.  A CTA whose tile index falls off the end (odd tile count) still loads, multiplies and commits
// (its partner depends on it) but stores nothing.
template <int MODE, int TILE_Y, int TILE_X, typename OutT, int EW, int CL = 1>
__global__ void __launch_bounds__(BuildCfg<EW>::THREADS, 1)
corr_build_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b0,
                  const __grid_constant__ CUtensorMap tm_b1, const __grid_constant__ CUtensorMap tm_b2,
                  const __grid_constant__ CUtensorMap tm_b3, const __grid_constant__ CUtensorMap tm_o0,
                  const __grid_constant__ CUtensorMap tm_o1, const __grid_constant__ CUtensorMap tm_o2,
                  const __grid_constant__ CUtensorMap tm_o3, const BuildParams p) {
    using Cfg = BuildCfg<EW>;
    constexpr int BLD_EPI_WARPS = Cfg::EPI_WARPS, BLD_SUBS = Cfg::SUBS, BLD_STG_BUFS = Cfg::STG_BUFS;
    constexpr int BLD_SMEM_BAR = Cfg::SMEM_BAR, BLD_SMEM_STG = Cfg::SMEM_STG, BLD_A_STAGES = Cfg::A_STAGES;
    static_assert(TILE_Y * TILE_X == BLD_BLOCK_N, "tile must hold 256 fmap2 pixels");
    static_assert(CL == 1 || (CL == 2 && MODE == MODE_LINEAR), "clusters: linear mode, pairs");
    static_assert(TILE_Y % 8 == 0 && TILE_X % 16 == 0, "sub-tiles are 8 x 16");
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    const uint32_t s_b = ptx::smem_u32(smem + BLD_SMEM_B);
    const uint32_t s_a = ptx::smem_u32(smem + BLD_SMEM_A);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BLD_SMEM_BAR);
    const uint32_t bar0 = ptx::smem_u32(bars);
    // barrier indices
    constexpr int A_FULL = 0, A_EMPTY = 4, B_FULL = 8, B_EMPTY = 9, T_FULL = 10, T_EMPTY = 12;
    static_assert(BLD_A_STAGES <= 4, "barrier slots");
    auto bar = [&](int i) { return bar0 + 8u * i; };
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(bars + 14);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&tm_a);
        ptx::prefetch_tensormap(&tm_b0);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < BLD_A_STAGES; ++i) {
            ptx::mbar_init(bar(A_FULL + i), 1);
            ptx::mbar_init(bar(A_EMPTY + i), CL);      // CL = 2: this CTA's commit and the peer's
        }
        ptx::mbar_init(bar(B_FULL), 1);
        ptx::mbar_init(bar(B_EMPTY), 1);
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(bar(T_FULL + i), 1);
            ptx::mbar_init(bar(T_EMPTY + i), BLD_EPI_WARPS);
        }
        ptx::fence_mbar_init();
    }
    if (warp == 2) {
        ptx::tmem_alloc(ptx::smem_u32(const_cast<uint32_t*>(tmem_slot)), 512);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    if constexpr (CL == 2) ptx::cluster_sync_relaxed_arrive();   // the peer's barriers exist before anything signals them
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // Work decomposition.  A work item is one stationary fmap2 tile (batch b, tile nt)
    // times one slice of the m-blocks; items are dealt round-robin, slice-major, so the
    // CTAs running at the same time sweep the SAME query rows: their stores land in one
    // moving window of the volume (a few 2 MB pages) instead of 148 scattered ones
    // (measured 2x), and they share the streamed fmap1 tiles in L2.
    const int ntiles = p.ntiles;       // N-tiles per batch item
    const int kc_n = p.kc;
    // CL = 1: items = (fmap2 tile of any batch item) x slice, dealt over the CTAs.  CL = 2: items = (PAIR of tiles
    // 2j, 2j + 1 of one batch item) x slice, dealt over the clusters; rank r of the cluster takes tile 2j + r.
    const uint32_t crank = (CL == 2) ? ptx::cluster_ctarank() : 0u;
    const int pairs_per_b = (ntiles + 1) / 2;
    const int units = (CL == 2) ? p.B * pairs_per_b : p.B * ntiles;
    const int n_items = units * p.msplit;
    const int item0 = (CL == 2) ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
    const int item_step = (CL == 2) ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
    // item -> (batch item, tile, slice, does this CTA own a real tile)
    auto decode = [&](int item, int& b, int& nt, int& sl) -> bool {
        const int u = item % units;
        sl = item / units;
        if constexpr (CL == 2) {
            b = u / pairs_per_b;
            nt = 2 * (u % pairs_per_b) + static_cast<int>(crank);
            if (nt >= ntiles) { nt = ntiles - 1; return false; }
            return true;
        } else {
            b = u / ntiles;
            nt = u % ntiles;
            return true;
        }
    };

    if (warp == 0) {
        // ===================== TMA producer =====================
        // The whole warp runs the loops and the barrier waits (converged: tensor-map pointers, coordinates and
        // barrier addresses stay in uniform registers); only the arrive / TMA instructions sit under elect_one().
        {
            uint32_t a_it = 0, b_it = 0;
            for (int item = item0; item < n_items; item += item_step) {
                int b, nt, sl;
                decode(item, b, nt, sl);
                const int mb0 = static_cast<int>(static_cast<long long>(sl) * p.m_blks / p.msplit);
                const int mb1 = static_cast<int>(static_cast<long long>(sl + 1) * p.m_blks / p.msplit);
                if (mb0 == mb1) continue;
                // new stationary fmap2 tile: wait until the MMAs reading the old one retired
                ptx::mbar_wait(bar(B_EMPTY), (b_it & 1) ^ 1);
                if constexpr (MODE == MODE_FUSED) {
                    const int y0 = (nt / p.ntx) * TILE_Y, x0 = (nt % p.ntx) * TILE_X;
                    if (ptx::elect_one()) {
                        ptx::mbar_arrive_expect_tx(bar(B_FULL), kc_n * BLD_B_SLAB_BYTES);
                        for (int kc = 0; kc < kc_n; ++kc)
                            ptx::tma_load_4d(s_b + kc * BLD_B_SLAB_BYTES, &tm_b0, bar(B_FULL),
                                             kc * BLD_BLOCK_K, x0, y0, b);
                    }
                } else {
                    int l = 0;
                    while (l + 1 < p.num_levels && nt >= p.tile_start[l + 1]) ++l;
                    const CUtensorMap* tm = (l == 0) ? &tm_b0 : (l == 1) ? &tm_b1 : (l == 2) ? &tm_b2 : &tm_b3;
                    const int c0 = (nt - p.tile_start[l]) * BLD_BLOCK_N;
                    if (ptx::elect_one()) {
                        ptx::mbar_arrive_expect_tx(bar(B_FULL), kc_n * BLD_B_SLAB_BYTES);
                        for (int kc = 0; kc < kc_n; ++kc)
                            ptx::tma_load_3d(s_b + kc * BLD_B_SLAB_BYTES, tm, bar(B_FULL),
                                             kc * BLD_BLOCK_K, c0, b);
                    }
                }
                __syncwarp();
                ++b_it;
                for (int mb = mb0; mb < mb1; ++mb) {
                    for (int kc = 0; kc < kc_n; ++kc, ++a_it) {
                        const uint32_t st = a_it % BLD_A_STAGES, ph = (a_it / BLD_A_STAGES) & 1;
                        ptx::mbar_wait(bar(A_EMPTY + st), ph ^ 1);
                        if (ptx::elect_one()) {
#ifdef RDVC_EXPERIMENTS
                            if ((p.dbg_store_mask & 16) && a_it >= BLD_A_STAGES) {  // debug: no A traffic
                                ptx::mbar_arrive(bar(A_FULL + st));
                            } else
#endif
                            {
                                ptx::mbar_arrive_expect_tx(bar(A_FULL + st), BLD_A_STAGE_BYTES);
                                if constexpr (CL == 2)      // my half of the stage (64 query rows), into BOTH CTAs
                                    ptx::tma_load_3d_multicast(s_a + st * BLD_A_STAGE_BYTES + crank * (BLD_A_STAGE_BYTES / 2), &tm_a,
                                                               bar(A_FULL + st), kc * BLD_BLOCK_K,
                                                               mb * BLD_BLOCK_M + static_cast<int>(crank) * (BLD_BLOCK_M / 2), b, 3);
                                else
                                    ptx::tma_load_3d(s_a + st * BLD_A_STAGE_BYTES, &tm_a, bar(A_FULL + st),
                                                     kc * BLD_BLOCK_K, mb * BLD_BLOCK_M, b);
                            }
                        }
                        __syncwarp();
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        // Converged warp, tcgen05 instructions under elect_one() (the form the MCN kernels use): with
        // `if (lane == 0)` around the loops every tcgen05.mma sat in an elect / R2UR.BROADCAST / branch wrapper
        // of ~9 instructions; here the descriptors live in uniform registers and the four MMAs of a k-block
        // are issued back to back.  Every commit is issued inside the SAME elect block as the MMAs it covers
        // (tcgen05.commit tracks the executing thread's operations).
        {
            const uint32_t idesc = ptx::umma_idesc(BLD_BLOCK_M, BLD_BLOCK_N, p.ab_format);   // 1 = bf16, 0 = fp16 operands
            uint32_t a_it = 0, b_it = 0, tile_it = 0;
            for (int item = item0; item < n_items; item += item_step) {
                const int sl = item / units;
                const int mb0 = static_cast<int>(static_cast<long long>(sl) * p.m_blks / p.msplit);
                const int mb1 = static_cast<int>(static_cast<long long>(sl + 1) * p.m_blks / p.msplit);
                if (mb0 == mb1) continue;
                ptx::mbar_wait(bar(B_FULL), b_it & 1);
                ++b_it;
                for (int mb = mb0; mb < mb1; ++mb, ++tile_it) {
                    const uint32_t acc = tile_it & 1, acc_ph = (tile_it >> 1) & 1;
                    ptx::mbar_wait(bar(T_EMPTY + acc), acc_ph ^ 1);
                    ptx::tc_fence_after();
                    const uint32_t d_tmem = tmem_base + acc * BLD_BLOCK_N;
                    for (int kc = 0; kc < kc_n; ++kc, ++a_it) {
                        const uint32_t st = a_it % BLD_A_STAGES, ph = (a_it / BLD_A_STAGES) & 1;
                        ptx::mbar_wait(bar(A_FULL + st), ph);
                        ptx::tc_fence_after();
                        const uint64_t a_desc0 = ptx::umma_desc_k_sw128(s_a + st * BLD_A_STAGE_BYTES);
                        const uint64_t b_desc0 = ptx::umma_desc_k_sw128(s_b + kc * BLD_B_SLAB_BYTES);
                        if (ptx::elect_one()) {
#pragma unroll
                            for (int k = 0; k < BLD_BLOCK_K / BLD_UMMA_K; ++k)
                                ptx::umma_bf16(d_tmem, a_desc0 + ((k * BLD_UMMA_K * 2) >> 4),
                                               b_desc0 + ((k * BLD_UMMA_K * 2) >> 4), idesc, (kc | k) != 0 ? 1u : 0u);
                            // frees the ring slot when the MMAs retire (CL = 2: in both CTAs -- the peer's multicast writes here too)
                            if constexpr (CL == 2) ptx::umma_commit_multicast(bar(A_EMPTY + st), 3);
                            else ptx::umma_commit(bar(A_EMPTY + st));
                            if (kc == kc_n - 1) {
                                ptx::umma_commit(bar(T_FULL + acc));      // accumulator ready for the epilogue
                                // every MMA that reads this fmap2 tile has been issued
                                if (mb == mb1 - 1) ptx::umma_commit(bar(B_EMPTY));
                            }
                        }
                        __syncwarp();
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp >= 4) {
        // ===================== epilogue =====================
        using TR = OutTraits<OutT>;
        const int e = warp - 4;
        const int q = e & 3;          // TMEM lane quarter this warp may read (warp_id % 4)
        const int sub0 = (e >> 2) * BLD_SUBS;   // first 128-column half of a tile this warp handles
        const uint32_t stg = ptx::smem_u32(smem + BLD_SMEM_STG) + e * BLD_STG_BYTES * BLD_STG_BUFS;
        uint32_t box_it = 0;  // TMA boxes issued by this warp (selects the staging buffer)
        const float scale = p.scale;
        const int L = p.num_levels;
#ifdef RDVC_EXPERIMENTS
        const int smask = p.dbg_store_mask;     // experiments only: levels to write, bit 5 = skip the TMEM reads
        const int dbg_policy = p.dbg_policy;
#else
        constexpr int smask = 15;
        constexpr int dbg_policy = 0;
#endif
        uint32_t tile_it = 0;

        if constexpr (MODE == MODE_LINEAR) {
            for (int item = item0; item < n_items; item += item_step) {
                int b, nt, sl;
                const bool active = decode(item, b, nt, sl);
                const int mb0 = static_cast<int>(static_cast<long long>(sl) * p.m_blks / p.msplit);
                const int mb1 = static_cast<int>(static_cast<long long>(sl + 1) * p.m_blks / p.msplit);
                int l = 0;
                while (l + 1 < L && nt >= p.tile_start[l + 1]) ++l;
                const int n_l = p.nl[l];                                 // pixels of this level
                const int tile_col0 = (nt - p.tile_start[l]) * BLD_BLOCK_N;
                OutT* const lv = static_cast<OutT*>(p.lvl[l]);
                const bool vec = (n_l % TR::EPC) == 0;
                const bool wr = ((smask >> l) & 1) && active;
                const int omode = (p.tma_out >> (2 * l)) & 3;
                const CUtensorMap* tmo = (l == 0) ? &tm_o0 : (l == 1) ? &tm_o1 : (l == 2) ? &tm_o2 : &tm_o3;
                for (int mb = mb0; mb < mb1; ++mb, ++tile_it) {
                    const int m0 = mb * BLD_BLOCK_M + q * 32;
                    int rows_valid = p.N - m0;
                    rows_valid = rows_valid < 0 ? 0 : (rows_valid > 32 ? 32 : rows_valid);
                    OutT* const img0 = lv + (static_cast<size_t>(b) * p.N + m0) * n_l;
                    const uint32_t acc = tile_it & 1, acc_ph = (tile_it >> 1) & 1;
                    ptx::mbar_wait(bar(T_FULL + acc), acc_ph);
                    ptx::tc_fence_after();
#pragma unroll
                    for (int sb_i = 0; sb_i < BLD_SUBS; ++sb_i) {
                    const int sub = sub0 + sb_i;
                    const bool last_sub = (sb_i == BLD_SUBS - 1);
                    const int col0 = tile_col0 + sub * (BLD_BLOCK_N / 2);
                    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                                           acc * BLD_BLOCK_N + sub * (BLD_BLOCK_N / 2);
                    if (omode == 2) {
                        // ---- wide path: every row visit of a TMA box is 256 contiguous bytes (measured
                        // 6.1 TB/s vs 5.5 TB/s for 128-byte visits, tools/micro/wbench3.cu).  A box is
                        // 16 query rows x two 128-byte column blocks = 4 KB, smem order [row][block][128 B]
                        // with TMA's SWIZZLE_128B; the warp's rows go out as two boxes (lanes 0-15, 16-31).
                        // BLD_STG_BUFS boxes per warp are in flight: a store takes ~0.8 us to drain, so the
                        // bytes in flight per SM (warps x buffers x 4 KB) bound the write rate.
                        constexpr int EPB = 128 / static_cast<int>(sizeof(OutT));   // elements per column block
                        constexpr int GC = 2 * EPB;                                  // columns per 256-byte group
                        constexpr int GROUPS = (BLD_BLOCK_N / 2) / GC;               // 2 (fp32) / 1 (bf16)
#pragma unroll
                        for (int g = 0; g < GROUPS; ++g) {
                            uint32_t pk[64];  // this thread's 256 bytes: block 0 = pk[0..31], block 1 = pk[32..63]
                            if constexpr (sizeof(OutT) == 4) {
                                float v[64];
#pragma unroll
                                for (int k = 0; k < 4; ++k) ptx::tmem_ld_x16(taddr + g * GC + k * 16, v + k * 16);
                                ptx::tmem_ld_wait();
#pragma unroll
                                for (int i = 0; i < 64; ++i) pk[i] = __float_as_uint(v[i] * scale);
                            } else {
#pragma unroll
                                for (int hb = 0; hb < 2; ++hb) {
                                    float v[64];
#pragma unroll
                                    for (int k = 0; k < 4; ++k)
                                        ptx::tmem_ld_x16(taddr + g * GC + hb * 64 + k * 16, v + k * 16);
                                    ptx::tmem_ld_wait();
#pragma unroll
                                    for (int i = 0; i < 32; ++i)
                                        pk[hb * 32 + i] = OutTraits<__nv_bfloat16>::pk(v[2 * i] * scale, v[2 * i + 1] * scale);
                                }
                            }
                            if (last_sub && g == GROUPS - 1) {
                                // every TMEM read of this tile is done: hand the accumulator back early
                                ptx::tc_fence_before();
                                __syncwarp();
                                if (lane == 0) ptx::mbar_arrive(bar(T_EMPTY + acc));
                            }
                            if (!wr) continue;
                            const int r = lane & 15;
                            const int flip = (r >> 2) & 1;   // lanes 4-7 / 12-15 write the other block first:
                                                             // the 8 lanes of a store phase then hit 8 different
                                                             // swizzled 16-byte columns (conflict-free)
#pragma unroll
                            for (int hh = 0; hh < 2; ++hh) {
                                // the box that used this buffer BLD_STG_BUFS boxes ago has left smem
                                if (lane == 0) ptx::bulk_wait_read<BLD_STG_BUFS - 1>();
                                __syncwarp();
                                const uint32_t sb = stg + (box_it % BLD_STG_BUFS) * BLD_STG_BYTES;
                                ++box_it;
                                if ((lane >> 4) == hh) {
#pragma unroll
                                    for (int ps = 0; ps < 2; ++ps) {
                                        const int cb = ps ^ flip;
                                        const int rr = r * 2 + cb;          // 128-byte row of the box in smem
#pragma unroll
                                        for (int c = 0; c < 8; ++c) {
                                            uint4 w;
                                            w.x = flip ? pk[(1 - ps) * 32 + c * 4 + 0] : pk[ps * 32 + c * 4 + 0];
                                            w.y = flip ? pk[(1 - ps) * 32 + c * 4 + 1] : pk[ps * 32 + c * 4 + 1];
                                            w.z = flip ? pk[(1 - ps) * 32 + c * 4 + 2] : pk[ps * 32 + c * 4 + 2];
                                            w.w = flip ? pk[(1 - ps) * 32 + c * 4 + 3] : pk[ps * 32 + c * 4 + 3];
                                            sts_16(sb + rr * 128 + ((c ^ (rr & 7)) << 4), w);
                                        }
                                    }
                                }
                                ptx::fence_proxy_async_smem();
                                __syncwarp();
                                if (lane == 0) {
                                    if (dbg_policy == 0) ptx::tma_store_4d(tmo, sb, 0, (col0 + g * GC) / EPB, m0 + 16 * hh, b);
                                    else ptx::tma_store_4d_hint(tmo, sb, 0, (col0 + g * GC) / EPB, m0 + 16 * hh, b,
                                                                dbg_policy == 1 ? ptx::policy_evict_last()
                                                                                  : ptx::policy_evict_first());
                                    ptx::bulk_commit();
                                }
                            }
                        }
                        continue;
                    }
                    // CW columns per pass: a staged row is always 128 bytes (32 fp32 or 64 bf16)
                    constexpr int CW = 128 / static_cast<int>(sizeof(OutT));
                    constexpr int PASSES = (BLD_BLOCK_N / 2) / CW;
#pragma unroll
                    for (int j = 0; j < PASSES; ++j) {
                        float v[CW];  // CW consecutive pixels of this thread's query row
                        if (!(smask & 32)) {  // debug bit 5: skip the TMEM reads
#pragma unroll
                            for (int k = 0; k < CW / 16; ++k) ptx::tmem_ld_x16(taddr + j * CW + k * 16, v + k * 16);
                            ptx::tmem_ld_wait();
                        } else {
#pragma unroll
                            for (int i = 0; i < CW; ++i) v[i] = 0.f;
                        }
                        if (last_sub && j == PASSES - 1) {
                            // every TMEM read of this tile is done: hand the accumulator back early
                            ptx::tc_fence_before();
                            __syncwarp();
                            if (lane == 0) ptx::mbar_arrive(bar(T_EMPTY + acc));
                        }
#pragma unroll
                        for (int i = 0; i < CW; ++i) v[i] *= scale;
                        if (wr) {
                            if (omode == 1) {
                                // staged rows are 128 B (8 x 16-byte chunks) with the chunk index XORed by
                                // (row & 7): exactly TMA's SWIZZLE_128B, so the engine un-swizzles on the way out
                                if (lane == 0) ptx::bulk_wait_read<BLD_STG_BUFS - 1>();
                                __syncwarp();
                                const uint32_t sb = stg + (box_it % BLD_STG_BUFS) * BLD_STG_BYTES;
                                ++box_it;
#pragma unroll
                                for (int c = 0; c < 8; ++c)
                                    sts_16(sb + (lane * 8 + (c ^ (lane & 7))) * 16, TR::pack(v + c * TR::EPC));
                                ptx::fence_proxy_async_smem();
                                __syncwarp();
                                if (lane == 0) {
                                    if (dbg_policy == 0) ptx::tma_store_3d(tmo, sb, col0 + j * CW, m0, b);
                                    else ptx::tma_store_3d_hint(tmo, sb, col0 + j * CW, m0, b,
                                                                dbg_policy == 1 ? ptx::policy_evict_last()
                                                                                  : ptx::policy_evict_first());
                                    ptx::bulk_commit();
                                }
                            } else {
                                // TMA boxes of a previous (aligned) level may still be reading the buffers
                                if (lane == 0) ptx::bulk_wait_read<0>();
                                __syncwarp();
#pragma unroll
                                for (int k = 0; k < CW / 32; ++k)
                                    staged_store<OutT, 32>(stg, v + 32 * k, lane, img0, n_l, rows_valid, 0,
                                                           col0 + j * CW + 32 * k, 1, n_l, vec);
                            }
                        }
                    }
                    }  // 128-column halves of this warp
                }
            }
            if (lane == 0) ptx::bulk_wait<0>();  // all TMA stores of this warp have landed
        } else {
        constexpr int SUBS_X = TILE_X / 16;
        OutT* const l0 = static_cast<OutT*>(p.lvl[0]);
        OutT* const l1 = static_cast<OutT*>(p.lvl[1]);
        OutT* const l2 = static_cast<OutT*>(p.lvl[2]);
        OutT* const l3 = static_cast<OutT*>(p.lvl[3]);
        const bool vec0 = (p.wl[0] % TR::EPC) == 0;
        const bool vec1 = (L > 1) && (p.wl[1] % TR::EPC) == 0;
        const bool vec2 = (L > 2) && (p.wl[2] % 4) == 0;
        const bool vec3 = (L > 3) && (p.wl[3] % 2) == 0;
        const size_t img0 = static_cast<size_t>(p.hl[0]) * p.wl[0];
        const size_t img1 = static_cast<size_t>(p.hl[1]) * p.wl[1];
        const size_t img2 = static_cast<size_t>(p.hl[2]) * p.wl[2];
        const size_t img3 = static_cast<size_t>(p.hl[3]) * p.wl[3];

        for (int item = item0; item < n_items; item += item_step) {
          int b, nt, sl;
          decode(item, b, nt, sl);
          const int mb0 = static_cast<int>(static_cast<long long>(sl) * p.m_blks / p.msplit);
          const int mb1 = static_cast<int>(static_cast<long long>(sl + 1) * p.m_blks / p.msplit);
          for (int mb = mb0; mb < mb1; ++mb, ++tile_it) {
            const int m0 = mb * BLD_BLOCK_M + q * 32;   // first query pixel of this warp
            int rows_valid = p.N - m0;
            rows_valid = rows_valid < 0 ? 0 : (rows_valid > 32 ? 32 : rows_valid);
            const size_t row0 = static_cast<size_t>(b) * p.N + m0;

            const uint32_t acc = tile_it & 1, acc_ph = (tile_it >> 1) & 1;
            ptx::mbar_wait(bar(T_FULL + acc), acc_ph);
            ptx::tc_fence_after();
            const uint32_t taddr =
                tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BLD_BLOCK_N;
#pragma unroll
            for (int sb_i = 0; sb_i < BLD_SUBS; ++sb_i) {
            const int sub = sub0 + sb_i;
            const bool last_sub = (sb_i == BLD_SUBS - 1);
            const int sy = (sub / SUBS_X) * 8, sx = (sub % SUBS_X) * 16;
            const int Y0 = (nt / p.ntx) * TILE_Y + sy;  // level-0 origin of this warp's 8x16 patch
            const int X0 = (nt % p.ntx) * TILE_X + sx;

            float p1[32];  // level 1 of this patch: 4 rows x 8
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float v[32];  // level-0 rows 2j, 2j+1 of the patch, 16 columns each
                ptx::tmem_ld_x16(taddr + (sy + 2 * j) * TILE_X + sx, v);
                ptx::tmem_ld_x16(taddr + (sy + 2 * j + 1) * TILE_X + sx, v + 16);
                ptx::tmem_ld_wait();
                if (last_sub && j == 3) {
                    // every TMEM read of this tile is done: hand the accumulator back early
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(bar(T_EMPTY + acc));
                }
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] *= scale;
#pragma unroll
                for (int x = 0; x < 8; ++x)
                    p1[j * 8 + x] =
                        (v[2 * x] + v[2 * x + 1] + v[16 + 2 * x] + v[16 + 2 * x + 1]) * 0.25f;
                if (smask & 1)
                    staged_store<OutT, 16>(stg, v, lane, l0 + row0 * img0, img0, rows_valid, Y0 + 2 * j,
                                           X0, p.hl[0], p.wl[0], vec0);
            }
            if (L > 1 && (smask & 2))
                staged_store<OutT, 8>(stg, p1, lane, l1 + row0 * img1, img1, rows_valid, Y0 >> 1,
                                      X0 >> 1, p.hl[1], p.wl[1], vec1);
            if (L > 2) {
                float p2[8];  // 2 rows x 4
#pragma unroll
                for (int y = 0; y < 2; ++y)
#pragma unroll
                    for (int x = 0; x < 4; ++x)
                        p2[y * 4 + x] = (p1[(2 * y) * 8 + 2 * x] + p1[(2 * y) * 8 + 2 * x + 1] +
                                         p1[(2 * y + 1) * 8 + 2 * x] + p1[(2 * y + 1) * 8 + 2 * x + 1]) *
                                        0.25f;
                const bool row_ok = lane < rows_valid;
                const int y2 = Y0 >> 2, x2 = X0 >> 2;
                if (row_ok && (smask & 4)) {
                    OutT* img = l2 + (row0 + lane) * img2;
#pragma unroll
                    for (int y = 0; y < 2; ++y) {
                        if (y2 + y >= p.hl[2]) continue;
                        OutT* dst = img + static_cast<size_t>(y2 + y) * p.wl[2] + x2;
                        if (vec2 && x2 + 3 < p.wl[2]) {
                            if constexpr (sizeof(OutT) == 4) {
                                *reinterpret_cast<float4*>(dst) = make_float4(
                                    p2[y * 4], p2[y * 4 + 1], p2[y * 4 + 2], p2[y * 4 + 3]);
                            } else {
                                *reinterpret_cast<uint2*>(dst) =
                                    make_uint2(OutTraits<__nv_bfloat16>::pk(p2[y * 4], p2[y * 4 + 1]),
                                               OutTraits<__nv_bfloat16>::pk(p2[y * 4 + 2], p2[y * 4 + 3]));
                            }
                        } else {
#pragma unroll
                            for (int x = 0; x < 4; ++x)
                                if (x2 + x < p.wl[2]) dst[x] = TR::cvt(p2[y * 4 + x]);
                        }
                    }
                }
                if (L > 3) {
                    const float p3a = (p2[0] + p2[1] + p2[4] + p2[5]) * 0.25f;
                    const float p3b = (p2[2] + p2[3] + p2[6] + p2[7]) * 0.25f;
                    const int y3 = Y0 >> 3, x3 = X0 >> 3;
                    if (row_ok && y3 < p.hl[3] && (smask & 8)) {
                        OutT* dst = l3 + (row0 + lane) * img3 + static_cast<size_t>(y3) * p.wl[3] + x3;
                        if (vec3 && x3 + 1 < p.wl[3]) {
                            if constexpr (sizeof(OutT) == 4) {
                                *reinterpret_cast<float2*>(dst) = make_float2(p3a, p3b);
                            } else {
                                *reinterpret_cast<uint32_t*>(dst) = OutTraits<__nv_bfloat16>::pk(p3a, p3b);
                            }
                        } else {
                            if (x3 < p.wl[3]) dst[0] = TR::cvt(p3a);
                            if (x3 + 1 < p.wl[3]) dst[1] = TR::cvt(p3b);
                        }
                    }
                }
            }
            }  // 8x16 patches of this warp
          }  // m-blocks of this item
        }      // items
        }      // MODE_FUSED
    }

    ptx::tc_fence_before();
    __syncthreads();
    if constexpr (CL == 2) ptx::cluster_sync_all();    // the peer's commits still arrive on this CTA's barriers until it is done
    if (warp == 2) ptx::tmem_dealloc(tmem_base, 512);
}

}  // namespace rdvc
