// mcn_convx_sm100.cuh -- the MCN convolution with ONE activation box per tile (x halo included).
//
// mcn_conv_sm100.cuh fetches three boxes per tile (column offsets dsx = -1, 0, +1) because a one-super-pixel
// shift in x is a one-row shift of the swizzled operand, and measured that the L2 -> SM feed those boxes need
// (3.75x the activation bytes, ~7.5 TB/s) bounds the layer.  Here the shift is moved from the INPUT to the OUTPUT:
//
//   out[sp] = sum_dy  A[sp]   . Wc(dy)          (column offset  0: "centre" accumulator C)
//           + sum_dy  A[sp-1] . Wl(dy)          (what super-pixel sp-1 sends to its right neighbour: accumulator L)
//           + sum_dy  A[sp+1] . Wr(dy)          (what super-pixel sp+1 sends to its left neighbour:  accumulator R)
//
// All three products use the SAME unshifted box rows; C, L and R are separate TMEM column ranges, and the
// epilogue adds L from the lane on the left and R from the lane on the right with warp shuffles (a warp holds
// two tile rows of 16 consecutive super-pixels).  The first and last column of a tile row are halo columns --
// computed, never written -- so a tile outputs 14 of its 16 columns (12.5 % redundant MMAs) and reads one
// 20 / 24 KB box per 224 pixels instead of three per 256: 2.6x less L2 -> SM traffic.  The tap matrices are the ones
// rdvc_mcn_pack_weights already produces: Wl(dy) is tap (dy, dsx = -1), Wr(dy) is tap (dy, dsx = +1); for 3 x 3 only
// the q = 0 rows of Wl and the q = 1 rows of Wr are non-zero, so those MMAs are N = 32.
// One ring stage is now one whole tile, so six (3 x 3) tiles of loads are in flight instead of two.
#pragma once
#include "mcn_conv_sm100.cuh"

namespace rdvc {

constexpr int MCNX_TXO = 14;   // output columns (super-pixels) of a tile row; the box is 16 wide

template <int R, int NOUT>
struct McnXCfg {
    using Base = McnCfg<R, NOUT>;
    static constexpr int NTAPS = Base::NTAPS;
    static constexpr int W_TAP_BYTES = Base::W_TAP_BYTES;
    static constexpr int W_BYTES = Base::W_BYTES;
    static constexpr int A_BYTES = Base::A_BYTES;
    static constexpr int NLR = (R == 1) ? NOUT / 2 : NOUT;       // columns of the L and of the R accumulator
    static constexpr int ACC_COLS = (NOUT == 16) ? 64 : NOUT + 2 * NLR;   // 128 (3x3), 192 (5x5 32->32), 48 -> 64 (last layer)
    static constexpr int NACC = (R == 2 && NOUT == 64) ? 2 : 4;
    static constexpr int TMEM_COLS = (ACC_COLS * NACC <= 256) ? 256 : 512;
    static constexpr int STAGES = (R == 2 && NOUT == 64) ? 3 : 6;
    static constexpr int ISSUERS = (STAGES % 2 == 0) ? 2 : 1;    // a ring slot must belong to one issuer (parity waits)
    static constexpr int SMEM_W = 0;
    static constexpr int SMEM_A = W_BYTES;
    static constexpr int SMEM_STG = SMEM_A + STAGES * A_BYTES;
    static constexpr int SMEM_BAR = SMEM_STG + 8 * MCN_STG_BYTES;
    static constexpr int SMEM_TOTAL = SMEM_BAR + 256;
    static constexpr int SMEM_LAUNCH = SMEM_TOTAL + 1024;
    static_assert(ACC_COLS * NACC <= 512, "TMEM columns");
    static_assert(SMEM_LAUNCH <= 227 * 1024, "shared memory");
};

template <int R, int NOUT, int KPAT>
__global__ void __launch_bounds__(MCN_THREADS, 1)
mcn_convx_kernel(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_w,
                 const __grid_constant__ CUtensorMap tm_out, const __grid_constant__ CUtensorMap tm_res,
                 const McnConvParams p) {
    using Cfg = McnXCfg<R, NOUT>;
    constexpr int NTAPS = Cfg::NTAPS, STAGES = Cfg::STAGES, NLR = Cfg::NLR, NACC = Cfg::NACC;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    const uint32_t s_w = ptx::smem_u32(smem + Cfg::SMEM_W);
    const uint32_t s_a = ptx::smem_u32(smem + Cfg::SMEM_A);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::SMEM_BAR);
    const uint32_t bar0 = ptx::smem_u32(bars);
    constexpr int A_FULL = 0, A_EMPTY = 6, W_FULL = 12, T_FULL = 13, T_EMPTY = 17, R_FULL = 21;   // 6, 6, 1, 4, 4, 8
    auto bar = [&](int i) { return bar0 + 8u * i; };
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(bars + 30);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&tm_in);
        ptx::prefetch_tensormap(&tm_w);
        ptx::prefetch_tensormap(&tm_out);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < STAGES; ++i) {
            ptx::mbar_init(bar(A_FULL + i), 1);
            ptx::mbar_init(bar(A_EMPTY + i), 1);
        }
        ptx::mbar_init(bar(W_FULL), 1);
        for (int i = 0; i < 4; ++i) {
            ptx::mbar_init(bar(T_FULL + i), 1);
            ptx::mbar_init(bar(T_EMPTY + i), 4);
        }
        for (int i = 0; i < 8; ++i) ptx::mbar_init(bar(R_FULL + i), 1);
        ptx::fence_mbar_init();
    }
    if (warp == 2) {
        ptx::tmem_alloc(ptx::smem_u32(const_cast<uint32_t*>(tmem_slot)), Cfg::TMEM_COLS);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int tiles_per_img = p.ntx * p.nty;
    const int n_tiles = p.B * tiles_per_img;

    if (warp == 0) {
        // ===================== TMA producer: one box per tile =====================
        if (lane == 0) {
            ptx::mbar_arrive_expect_tx(bar(W_FULL), Cfg::W_BYTES);
            for (int t = 0; t < NTAPS; ++t)
                ptx::tma_load_3d(s_w + t * Cfg::W_TAP_BYTES, &tm_w, bar(W_FULL), 0, 0, t);
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
                const int tw = p.reverse ? n_tiles - 1 - tile : tile;
                const int b = tw / tiles_per_img, rem = tw % tiles_per_img;
                const int y0 = (rem / p.ntx) * MCN_TY, xo = (rem % p.ntx) * MCNX_TXO;
                const uint32_t st = it % STAGES, ph = (it / STAGES) & 1;
                ptx::mbar_wait(bar(A_EMPTY + st), ph ^ 1);
                ptx::mbar_arrive_expect_tx(bar(A_FULL + st), Cfg::A_BYTES);
                ptx::tma_load_4d(s_a + st * Cfg::A_BYTES, &tm_in, bar(A_FULL + st), 0, xo - 1, y0 - R, b);
            }
        }
        __syncwarp();
    } else if (warp == 1 || warp == 3) {
        // ===================== MMA issuers (even / odd tiles) =====================
        const uint32_t mma_id = warp >> 1;
        if (mma_id < Cfg::ISSUERS) {      // the whole warp runs the loop; only the tcgen05 instructions are elected (see mcn_conv_kernel)
            const uint32_t idesc_c = ptx::umma_idesc(128, NOUT, 0);   // fp16 operands, fp32 accumulate
            const uint32_t idesc_s = ptx::umma_idesc(128, NLR, 0);
            ptx::mbar_wait(bar(W_FULL), 0);
            ptx::tc_fence_after();
            const uint64_t b_desc0 = ptx::umma_desc_k_sw128(s_w);
            for (uint32_t tile_it = mma_id; blockIdx.x + static_cast<long long>(tile_it) * gridDim.x < n_tiles;
                 tile_it += Cfg::ISSUERS) {
                const uint32_t acc = tile_it % NACC, acc_ph = (tile_it / NACC) & 1;
                ptx::mbar_wait(bar(T_EMPTY + acc), acc_ph ^ 1);
                ptx::tc_fence_after();
                const uint32_t d_c = tmem_base + acc * Cfg::ACC_COLS;
                const uint32_t d_l = d_c + NOUT, d_r = d_c + NOUT + NLR;
                const uint32_t st = tile_it % STAGES, ph = (tile_it / STAGES) & 1;
                ptx::mbar_wait(bar(A_FULL + st), ph);
                ptx::tc_fence_after();
                const uint64_t a_desc0 = ptx::umma_desc_k_sw128(s_a + st * Cfg::A_BYTES);
                // rows of a side tap that can be non-zero: 3x3 -> q = 0 (first half) for L, q = 1 (second half) for R
                constexpr int R_ROW0 = (R == 1) ? NOUT / 2 : 0;
                if (ptx::elect_one()) {
#pragma unroll
                for (int dy = 0; dy < 2 * R + 1; ++dy) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t a_desc = a_desc0 + ((dy * MCN_ROW_BYTES + k * 32) >> 4);
                        if (mcn_kstep_on(KPAT, 1, k)) {     // centre
                            constexpr int k0 = 0;
                            ptx::umma_bf16(d_c, a_desc, b_desc0 + (((dy * 3 + 1) * Cfg::W_TAP_BYTES + k * 32) >> 4), idesc_c,
                                           (dy == 0 && k == k0) ? 0u : 1u);
                        }
                        if (mcn_kstep_on(KPAT, 0, k)) {     // tap dsx = -1: this super-pixel's contribution to its RIGHT neighbour
                            constexpr int k0 = (KPAT == MCN_K_3X3) ? 2 : 0;
                            ptx::umma_bf16(d_l, a_desc, b_desc0 + (((dy * 3 + 0) * Cfg::W_TAP_BYTES + k * 32) >> 4), idesc_s,
                                           (dy == 0 && k == k0) ? 0u : 1u);
                        }
                        if (mcn_kstep_on(KPAT, 2, k)) {     // tap dsx = +1: contribution to the LEFT neighbour
                            constexpr int k0 = 0;
                            ptx::umma_bf16(d_r, a_desc,
                                           b_desc0 + (((dy * 3 + 2) * Cfg::W_TAP_BYTES + R_ROW0 * 128 + k * 32) >> 4), idesc_s,
                                           (dy == 0 && k == k0) ? 0u : 1u);
                        }
                    }
                }
                ptx::umma_commit(bar(A_EMPTY + st));
                ptx::umma_commit(bar(T_FULL + acc));
                }
                __syncwarp();
            }
        }
        __syncwarp();
    } else if (warp >= 4) {
        // ===================== epilogue: 8 warps = lane quarter x tile parity =====================
        const int e = warp - 4;
        const int q = e & 3;
        const uint32_t par = e >> 2;
        const int ty = q * 2 + (lane >> 4), tx = lane & 15;
        const bool col_ok = (tx >= 1) && (tx <= MCNX_TXO);          // columns 0 and 15 are halo
        const int srow = (lane >> 4) * MCNX_TXO + (tx - 1);         // this thread's row of the warp's (14, 2) staging box
        const uint32_t sb = ptx::smem_u32(smem + Cfg::SMEM_STG) + e * MCN_STG_BYTES;
        for (uint32_t tile_it = par; blockIdx.x + static_cast<long long>(tile_it) * gridDim.x < n_tiles; tile_it += 2) {
            const int tile = blockIdx.x + tile_it * gridDim.x;
            const int tw = p.reverse ? n_tiles - 1 - tile : tile;
            const int b = tw / tiles_per_img, rem = tw % tiles_per_img;
            const int y0 = (rem / p.ntx) * MCN_TY, xo = (rem % p.ntx) * MCNX_TXO;
            const int y = y0 + ty, sp = xo - 1 + tx;
            const bool inside = col_ok && (y < p.H) && (sp < p.Wsp);
            const uint32_t acc = tile_it % NACC, acc_ph = (tile_it / NACC) & 1;
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * Cfg::ACC_COLS;

            if constexpr (NOUT == 64) {
                const bool use_res = (p.residual != nullptr);
                if (lane == 0) {
                    ptx::bulk_wait_read<0>();   // this warp's previous store has finished reading `sb`
                    if (use_res) {
                        ptx::mbar_arrive_expect_tx(bar(R_FULL + e), 2 * MCNX_TXO * 128);
                        ptx::tma_load_4d(sb, &tm_res, bar(R_FULL + e), 0, xo, y0 + 2 * q, b);
                    }
                }
                __syncwarp();
                ptx::mbar_wait(bar(T_FULL + acc), acc_ph);
                ptx::tc_fence_after();
                float v[64];
#pragma unroll
                for (int k = 0; k < 4; ++k) ptx::tmem_ld_x16(taddr + k * 16, v + k * 16);
                ptx::tmem_ld_wait();
                // L comes from the lane on the left, R from the lane on the right (lanes 0-15 and 16-31 are two tile
                // rows; what crosses between them lands in halo columns only)
#pragma unroll
                for (int c = 0; c < NLR / 16; ++c) {
                    float l[16], r[16];
                    ptx::tmem_ld_x16(taddr + NOUT + c * 16, l);
                    ptx::tmem_ld_x16(taddr + NOUT + NLR + c * 16, r);
                    ptx::tmem_ld_wait();
                    constexpr int RBASE = (R == 1) ? 32 : 0;        // 3x3: L feeds pixel q = 0, R feeds pixel q = 1
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        v[c * 16 + i] += __shfl_up_sync(0xffffffffu, l[i], 1);
                        v[RBASE + c * 16 + i] += __shfl_down_sync(0xffffffffu, r[i], 1);
                    }
                }
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(bar(T_EMPTY + acc));   // accumulator back to the MMA warp
#pragma unroll
                for (int i = 0; i < 64; ++i) v[i] += p.bias[i & 31];
                if (use_res) {
                    ptx::mbar_wait(bar(R_FULL + e), (tile_it >> 1) & 1);
                    if (col_ok) {
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            const uint4 r4 = mcn_lds_16(sb + srow * 128 + ((c ^ (srow & 7)) << 4));
                            const uint32_t w4[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w4[j]));
                                v[c * 8 + 2 * j] += f.x;
                                v[c * 8 + 2 * j + 1] += f.y;
                            }
                        }
                    }
                }
                if (p.act == 1) {
#pragma unroll
                    for (int i = 0; i < 64; ++i) v[i] = v[i] > 0.f ? v[i] : 0.2f * v[i];
                }
                if (2 * sp + 1 >= p.W) {   // the padding pixel of an odd-width row stays zero
#pragma unroll
                    for (int i = 32; i < 64; ++i) v[i] = 0.f;
                }
                if (col_ok) {
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        uint4 w;
                        w.x = mcn_pack_h2(v[c * 8 + 0], v[c * 8 + 1]);
                        w.y = mcn_pack_h2(v[c * 8 + 2], v[c * 8 + 3]);
                        w.z = mcn_pack_h2(v[c * 8 + 4], v[c * 8 + 5]);
                        w.w = mcn_pack_h2(v[c * 8 + 6], v[c * 8 + 7]);
                        mcn_sts_16(sb + srow * 128 + ((c ^ (srow & 7)) << 4), w);   // TMA's SWIZZLE_128B
                    }
                }
                ptx::fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    ptx::tma_store_4d(&tm_out, sb, 0, xo, y0 + 2 * q, b);   // (64, 14, 2) box, clipped to the tensor
                    ptx::bulk_commit();
                }
            } else {
                // last layer: 2 pixels x 8 padded channels per thread; out = warped * sigmoid(conv + bias)
                const int x = 2 * sp;
                const bool ok0 = inside && (x < p.W), ok1 = inside && (x + 1 < p.W);
                const bool pair = ok1 && ((p.W & 1) == 0);
                float wv[8][2];
                const size_t plane = static_cast<size_t>(p.H) * p.W;
                const size_t off0 = static_cast<size_t>(b) * p.cout * plane + static_cast<size_t>(y) * p.W + x;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    wv[c][0] = wv[c][1] = 0.f;
                    if (c < p.cout) {
                        if (pair) {
                            const float2 f = __ldg(reinterpret_cast<const float2*>(p.warped + off0 + c * plane));
                            wv[c][0] = f.x; wv[c][1] = f.y;
                        } else {
                            if (ok0) wv[c][0] = __ldg(p.warped + off0 + c * plane);
                            if (ok1) wv[c][1] = __ldg(p.warped + off0 + c * plane + 1);
                        }
                    }
                }
                ptx::mbar_wait(bar(T_FULL + acc), acc_ph);
                ptx::tc_fence_after();
                float v[16], l[16], r[16];
                ptx::tmem_ld_x16(taddr, v);
                ptx::tmem_ld_x16(taddr + NOUT, l);
                ptx::tmem_ld_x16(taddr + NOUT + NLR, r);
                ptx::tmem_ld_wait();
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(bar(T_EMPTY + acc));
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    v[i] += __shfl_up_sync(0xffffffffu, l[i], 1) + __shfl_down_sync(0xffffffffu, r[i], 1);
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    if (c < p.cout) {
                        const float s0 = 1.f / (1.f + __expf(-(v[c] + p.bias[c])));
                        const float s1 = 1.f / (1.f + __expf(-(v[8 + c] + p.bias[c])));
                        if (pair) {
                            *reinterpret_cast<float2*>(p.out + off0 + c * plane) = make_float2(wv[c][0] * s0, wv[c][1] * s1);
                        } else {
                            if (ok0) p.out[off0 + c * plane] = wv[c][0] * s0;
                            if (ok1) p.out[off0 + c * plane + 1] = wv[c][1] * s1;
                        }
                    }
                }
            }
        }
        if constexpr (NOUT == 64) {
            if (lane == 0) ptx::bulk_wait<0>();   // every TMA store of this warp has landed
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 2) ptx::tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

}  // namespace rdvc
