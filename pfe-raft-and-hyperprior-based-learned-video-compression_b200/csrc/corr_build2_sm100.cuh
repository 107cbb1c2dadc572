// corr_build2_sm100.cuh -- the linear-mode build (see corr_build_sm100.cuh) on CTA PAIRS.
//
// Same arithmetic, same tile (256 fmap2 pixels stationary, fmap1 streamed), same epilogue -- but the
// tcgen05.mma is `cta_group::2`: two CTAs on the two SMs of a TPC form one M = 256 x N = 256 MMA.
// Each CTA streams its own 128 query rows (fmap1) and holds only HALF of the stationary fmap2 tile
// (128 pixels x 256 channels = 64 KB instead of 128 KB); the tensor core reads both halves.  That
// frees 64 KB of shared memory per CTA, which is what the build is short of (BuildCfg in
// corr_build_sm100.cuh): here a 4-stage fmap1 ring AND three 4 KB staging boxes for each of 8
// epilogue warps fit (96 KB of stores in flight per SM instead of 32).
//
// Protocol (all barriers exist at the same shared-memory offset in both CTAs; "leader" = cluster rank 0):
//   A_FULL[s], B_FULL   leader only.  Leader's producer arms them (expect_tx for BOTH CTAs' bytes);
//                       each CTA's TMA load completes its bytes on the leader's barrier.
//   A_EMPTY[s], B_EMPTY both CTAs, signalled by the leader's tcgen05.commit multicast to {0, 1}.
//   T_FULL[a]           both CTAs, tcgen05.commit multicast: accumulator a is ready in both TMEMs.
//   T_EMPTY[a]          leader only, 16 arrivals: every epilogue warp of both CTAs (remote arrive).
// Only the leader issues MMAs; both CTAs run a producer and 8 epilogue warps.
#pragma once
#include "corr_build_sm100.cuh"

namespace rdvc {

constexpr int B2_A_STAGES = 4;
constexpr int B2_STG_BUFS = 3;
constexpr int B2_EPI_WARPS = 8;
constexpr int B2_THREADS = 128 + B2_EPI_WARPS * 32;
constexpr int B2_BHALF_SLAB = (BLD_BLOCK_N / 2) * BLD_BLOCK_K * 2;          // 16 KB: 128 pixels x 64 channels
constexpr int B2_SMEM_B = 0;                                                // 4 slabs = 64 KB
constexpr int B2_SMEM_A = B2_SMEM_B + BLD_MAX_KC * B2_BHALF_SLAB;           // 65536
constexpr int B2_SMEM_STG = B2_SMEM_A + B2_A_STAGES * BLD_A_STAGE_BYTES;    // 131072
constexpr int B2_SMEM_BAR = B2_SMEM_STG + B2_EPI_WARPS * B2_STG_BUFS * BLD_STG_BYTES;   // 229376
constexpr int B2_SMEM_TOTAL = B2_SMEM_BAR + 256;
constexpr int B2_SMEM_LAUNCH = B2_SMEM_TOTAL + 1024;

// BuildParams as for the single-CTA kernel, except: m_blks = ceil(N / 256) (row blocks of a PAIR),
// msplit chosen for gridDim.x / 2 clusters, every level written with the wide (omode 2) boxes.
template <typename OutT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(B2_THREADS, 1)
corr_build2_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b0,
                   const __grid_constant__ CUtensorMap tm_b1, const __grid_constant__ CUtensorMap tm_b2,
                   const __grid_constant__ CUtensorMap tm_b3, const __grid_constant__ CUtensorMap tm_o0,
                   const __grid_constant__ CUtensorMap tm_o1, const __grid_constant__ CUtensorMap tm_o2,
                   const __grid_constant__ CUtensorMap tm_o3, const BuildParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    const uint32_t s_b = ptx::smem_u32(smem + B2_SMEM_B);
    const uint32_t s_a = ptx::smem_u32(smem + B2_SMEM_A);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + B2_SMEM_BAR);
    const uint32_t bar0 = ptx::smem_u32(bars);
    constexpr int A_FULL = 0, A_EMPTY = 4, B_FULL = 8, B_EMPTY = 9, T_FULL = 10, T_EMPTY = 12;
    auto bar = [&](int i) { return bar0 + 8u * i; };
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(bars + 14);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = ptx::cluster_ctarank();
    const bool leader = (rank == 0);
    const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&tm_a);
        ptx::prefetch_tensormap(&tm_b0);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < B2_A_STAGES; ++i) {
            ptx::mbar_init(bar(A_FULL + i), 1);
            ptx::mbar_init(bar(A_EMPTY + i), 1);
        }
        ptx::mbar_init(bar(B_FULL), 1);
        ptx::mbar_init(bar(B_EMPTY), 1);
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(bar(T_FULL + i), 1);
            ptx::mbar_init(bar(T_EMPTY + i), 2 * B2_EPI_WARPS);
        }
        ptx::fence_mbar_init();
    }
    if (warp == 2) {
        ptx::tmem_alloc_2sm(ptx::smem_u32(const_cast<uint32_t*>(tmem_slot)), 512);
        ptx::tmem_relinquish_2sm();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::cluster_sync_all();          // the peer's barriers are initialised before anyone signals them
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int ntiles = p.ntiles;
    const int units = p.B * ntiles;
    const int n_items = units * p.msplit;
    const int kc_n = p.kc;

    if (warp == 0) {
        // ===================== TMA producer (both CTAs) =====================
        if (lane == 0) {
            uint32_t a_it = 0, b_it = 0;
            for (int item = cluster_id; item < n_items; item += n_clusters) {
                const int u = item % units, sl = item / units;
                const int b = u / ntiles, nt = u % ntiles;
                const int mb0 = static_cast<int>(static_cast<long long>(sl) * p.m_blks / p.msplit);
                const int mb1 = static_cast<int>(static_cast<long long>(sl + 1) * p.m_blks / p.msplit);
                if (mb0 == mb1) continue;
                ptx::mbar_wait(bar(B_EMPTY), (b_it & 1) ^ 1);
                if (leader) ptx::mbar_arrive_expect_tx(bar(B_FULL), 2 * kc_n * B2_BHALF_SLAB);
                int l = 0;
                while (l + 1 < p.num_levels && nt >= p.tile_start[l + 1]) ++l;
                const CUtensorMap* tm = (l == 0) ? &tm_b0 : (l == 1) ? &tm_b1 : (l == 2) ? &tm_b2 : &tm_b3;
                const int c0 = (nt - p.tile_start[l]) * BLD_BLOCK_N + static_cast<int>(rank) * (BLD_BLOCK_N / 2);
                for (int kc = 0; kc < kc_n; ++kc)
                    ptx::tma_load_3d_2sm(s_b + kc * B2_BHALF_SLAB, tm, bar(B_FULL), kc * BLD_BLOCK_K, c0, b);
                ++b_it;
                for (int mb = mb0; mb < mb1; ++mb) {
                    const int m0 = mb * (2 * BLD_BLOCK_M) + static_cast<int>(rank) * BLD_BLOCK_M;
                    for (int kc = 0; kc < kc_n; ++kc, ++a_it) {
                        const uint32_t st = a_it % B2_A_STAGES, ph = (a_it / B2_A_STAGES) & 1;
                        ptx::mbar_wait(bar(A_EMPTY + st), ph ^ 1);
                        if (leader) ptx::mbar_arrive_expect_tx(bar(A_FULL + st), 2 * BLD_A_STAGE_BYTES);
                        ptx::tma_load_3d_2sm(s_a + st * BLD_A_STAGE_BYTES, &tm_a, bar(A_FULL + st),
                                             kc * BLD_BLOCK_K, m0, b);
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (leader && lane == 0) {
            const uint32_t idesc = ptx::umma_idesc(2 * BLD_BLOCK_M, BLD_BLOCK_N, p.ab_format);   // 1 = bf16, 0 = fp16 operands
            uint32_t a_it = 0, b_it = 0, tile_it = 0;
            for (int item = cluster_id; item < n_items; item += n_clusters) {
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
                        const uint32_t st = a_it % B2_A_STAGES, ph = (a_it / B2_A_STAGES) & 1;
                        ptx::mbar_wait(bar(A_FULL + st), ph);
                        ptx::tc_fence_after();
                        const uint32_t a_addr = s_a + st * BLD_A_STAGE_BYTES;
                        const uint32_t b_addr = s_b + kc * B2_BHALF_SLAB;
#pragma unroll
                        for (int k = 0; k < BLD_BLOCK_K / BLD_UMMA_K; ++k) {
                            ptx::umma_bf16_2sm(d_tmem, ptx::umma_desc_k_sw128(a_addr + k * BLD_UMMA_K * 2),
                                               ptx::umma_desc_k_sw128(b_addr + k * BLD_UMMA_K * 2), idesc,
                                               (kc | k) != 0 ? 1u : 0u);
                        }
                        ptx::umma_commit_2sm(bar(A_EMPTY + st), 3);   // ring slot free in both CTAs
                    }
                    ptx::umma_commit_2sm(bar(T_FULL + acc), 3);       // accumulator ready in both CTAs
                }
                ptx::umma_commit_2sm(bar(B_EMPTY), 3);
            }
        }
        __syncwarp();
    } else if (warp >= 4) {
        // ===================== epilogue (both CTAs): wide boxes only =====================
        const int e = warp - 4;
        const int q = e & 3;          // TMEM lane quarter this warp may read (warp_id % 4)
        const int sub = e >> 2;       // which 128-column half of the tile
        const uint32_t stg = ptx::smem_u32(smem + B2_SMEM_STG) + e * BLD_STG_BYTES * B2_STG_BUFS;
        const uint32_t t_empty_leader0 = ptx::mapa(bar(T_EMPTY), 0);
        uint32_t box_it = 0, tile_it = 0;
        const float scale = p.scale;
        const int L = p.num_levels;
        constexpr int EPB = 128 / static_cast<int>(sizeof(OutT));   // elements per 128-byte column block
        constexpr int GC = 2 * EPB;                                  // columns per 256-byte group
        constexpr int GROUPS = (BLD_BLOCK_N / 2) / GC;               // 2 (fp32) / 1 (bf16)
        for (int item = cluster_id; item < n_items; item += n_clusters) {
            const int u = item % units, sl = item / units;
            const int b = u / ntiles, nt = u % ntiles;
            const int mb0 = static_cast<int>(static_cast<long long>(sl) * p.m_blks / p.msplit);
            const int mb1 = static_cast<int>(static_cast<long long>(sl + 1) * p.m_blks / p.msplit);
            int l = 0;
            while (l + 1 < L && nt >= p.tile_start[l + 1]) ++l;
            const int col0 = (nt - p.tile_start[l]) * BLD_BLOCK_N + sub * (BLD_BLOCK_N / 2);
            const bool wr = (p.dbg_store_mask >> l) & 1;
            const CUtensorMap* tmo = (l == 0) ? &tm_o0 : (l == 1) ? &tm_o1 : (l == 2) ? &tm_o2 : &tm_o3;
            for (int mb = mb0; mb < mb1; ++mb, ++tile_it) {
                const int m0 = mb * (2 * BLD_BLOCK_M) + static_cast<int>(rank) * BLD_BLOCK_M + q * 32;
                const uint32_t acc = tile_it & 1, acc_ph = (tile_it >> 1) & 1;
                ptx::mbar_wait(bar(T_FULL + acc), acc_ph);
                ptx::tc_fence_after();
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                                       acc * BLD_BLOCK_N + sub * (BLD_BLOCK_N / 2);
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
                    if (g == GROUPS - 1) {
                        // every TMEM read of this tile is done: hand the accumulator back (to the leader)
                        ptx::tc_fence_before();
                        __syncwarp();
                        if (lane == 0) ptx::mbar_arrive_cluster_relaxed(t_empty_leader0 + 8u * acc);   // relaxed: see ptx_sm100.cuh
                    }
                    if (!wr) continue;
                    const int r = lane & 15;
                    const int flip = (r >> 2) & 1;   // see corr_build_sm100.cuh: conflict-free swizzled staging
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        if (lane == 0) ptx::bulk_wait_read<B2_STG_BUFS - 1>();
                        __syncwarp();
                        const uint32_t sb = stg + (box_it % B2_STG_BUFS) * BLD_STG_BYTES;
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
                            ptx::tma_store_4d(tmo, sb, 0, (col0 + g * GC) / EPB, m0 + 16 * hh, b);
                            ptx::bulk_commit();
                        }
                    }
                }
            }
        }
        if (lane == 0) ptx::bulk_wait<0>();  // all TMA stores of this warp have landed
    }

    ptx::tc_fence_before();
    __syncthreads();
    ptx::cluster_sync_all();          // the leader's MMAs read the peer's shared memory until the very end
    if (warp == 2) ptx::tmem_dealloc_2sm(tmem_base, 512);
}

}  // namespace rdvc
