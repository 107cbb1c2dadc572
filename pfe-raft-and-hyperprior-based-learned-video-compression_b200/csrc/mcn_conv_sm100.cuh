// mcn_conv_sm100.cuh -- the motion-compensation network's convolutions ("next" row f-4:
// MotionCompensationNetwork, R:codec_processing.py:369-406, called per P-frame at :1458) as
// tcgen05 implicit GEMMs over fp16 NHWC activations.
//
//   network = conv5x5(8 -> 32) + BN + LeakyReLU(0.2)
//             3 x [ conv3x3 + BN + LeakyReLU ; conv3x3 + BN ; + block input ; LeakyReLU ]
//             conv5x5(32 -> 3) + bias ; sigmoid ;  out = warped_ref * sigmoid(...)
//
// BatchNorm (inference) is folded into the weights and a per-channel bias on the host, so every
// layer is  out = act(conv(in) + bias [+ residual]).
//
// Activation layout in HBM: [B][H][Wsp][64] fp16, Wsp = ceil(W / 2): plain NHWC with 32 channels
// where two horizontally adjacent pixels ("a super-pixel") share one 128-byte row -- exactly one
// SWIZZLE_128B row of a K-major UMMA operand.  A k x k convolution over pixels is then a sum over
// taps (dy in [-R, R], dsx in {-1, 0, 1} super-pixels) of [128 super-pixels x 64] x [64 x NOUT]
// GEMMs, NOUT = 2 pixels x 32 output channels; the weight matrix of a tap holds w[co][ci][dy][dx]
// at row (q, co), column (p, ci) with dx = 2 dsx + p - q, zero where |dx| > R (packed once on the
// host by rdvc_mcn_pack_weights; k-steps whose 16 columns are all zero are left out of the
// compile-time MMA schedule, so a 3 x 3 layer issues 24 instead of 36 MMAs per tile).  The A operand of
// a tap is the activation tensor itself, shifted.  One 4-D TMA box (64 ch', 16 super-pixels,
// 8 + 2R rows) at (x0 + dsx, y0 - R) serves all 2R + 1 vertical taps of a column offset dsx: a
// tile row is 16 x 128 B = two whole 1 KB swizzle atoms, so the box shifted down by dy rows is
// again a valid SWIZZLE_128B operand (descriptor start address + dy * 2 KB) -- 3 boxes of 20 / 24 KB
// per tile instead of 9 / 15 of 16 KB (measured: L2->SM traffic was the limiter, 174 us per 3x3
// layer at 1080p with one box per tap).  The zero padding of the convolution is TMA's
// out-of-bounds fill.
//
// Persistent, one CTA per SM, warp-specialised like the correlation build:
//   warp 0  TMA producer: all tap matrices once (stationary, 72 / 120 KB), then the activation
//           boxes through a ring of 6 (3 for the 5x5 32 -> 32 shape) stages
//   warps 1, 3  tcgen05.mma issuers (even / odd tiles): M128 x NOUT x K16, fp16 x fp16 -> fp32 in TMEM,
//           two accumulators each
//   warp 2  TMEM allocator
//   warps 4-11 epilogue (lane quarter x tile parity): tcgen05.ld (thread = super-pixel) -> + bias (+ residual, a TMA load of the
//           warp's 4 KB box into its staging buffer) -> LeakyReLU -> fp16 -> swizzled smem row
//           -> one TMA store per warp of (64, 16, 2) = 4 KB;
//           the last layer instead applies sigmoid x warped_ref and writes NCHW fp32.
// HBM sees each activation once in and once out (266 MB per layer at 1080p); L2 -> SM traffic is
// 3 x (8 + 2R) / 8 of that -- measured to be what bounds this kernel.  mcn_convx_sm100.cuh is the variant
// with ONE box per tile (the x shift moved to the output side); the ABI picks per layer (option key 14).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cstdint>

#include "ptx_sm100.cuh"

namespace rdvc {

constexpr int MCN_C = 32;        // channels per pixel of the activation layout
constexpr int MCN_TX = 16;       // tile width in super-pixels (32 pixels)
constexpr int MCN_TY = 8;        // tile height in rows
constexpr int MCN_ROW_BYTES = MCN_TX * 128;           // 2 KB: one tile row = two 1 KB swizzle atoms
constexpr int MCN_STG_BYTES = 4096;                  // one epilogue warp's store box (32 rows x 128 B)
constexpr int MCN_THREADS = 384;                      // 4 service warps + 8 epilogue warps

// Which 16-column k-steps of a tap's 64-column matrix the MMA warp issues -- a COMPILE-TIME schedule, so the
// issue loop is straight-line code (two 64-bit adds + one tcgen05.mma per step).  With M128 x N64 x K16 MMAs
// (~32 tensor cycles each) a loop that tests a run-time mask and rebuilds descriptors (~25 dependent
// single-thread instructions per MMA) made the ISSUE the bottleneck: 132 us per 3x3 layer at 1080p, tensor
// pipe 19 % busy, producer and epilogue both waiting.
//   MCN_K_FULL: every k-step.
//   MCN_K_3X3 : 3x3 over all 32 channels: the left / right super-pixel column contributes through one of its
//               two pixels only (dx = 2 dsx + p - q must lie in [-1, 1]).
//   MCN_K_C16 : at most 16 input channels (the network's first layer has 8): k-steps 0 and 2.
constexpr int MCN_K_FULL = 0, MCN_K_3X3 = 1, MCN_K_C16 = 2;
__host__ __device__ constexpr bool mcn_kstep_on(int kpat, int dsx_idx, int k) {
    return kpat == MCN_K_FULL  ? true
           : kpat == MCN_K_3X3 ? (dsx_idx == 0 ? k >= 2 : dsx_idx == 1 ? true : k < 2)
                               : (k == 0 || k == 2);
}
// the run-time mask (rdvc_mcn_pack_weights) a schedule covers: bit (4 t + k), t = dy_idx * 3 + dsx_idx
inline unsigned long long mcn_pattern_mask(int kpat, int ksize) {
    unsigned long long m = 0;
    for (int dy = 0; dy < ksize; ++dy)
        for (int dsx = 0; dsx < 3; ++dsx)
            for (int k = 0; k < 4; ++k)
                if (mcn_kstep_on(kpat, dsx, k)) m |= 1ull << (4 * (dy * 3 + dsx) + k);
    return m;
}

template <int R, int NOUT>
struct McnCfg {
    static_assert(R == 1 || R == 2, "3x3 or 5x5");
    static_assert(NOUT == 64 || NOUT == 16, "2 pixels x 32 channels, or 2 pixels x 8 (last layer)");
    static constexpr int NTAPS = (2 * R + 1) * 3;
    static constexpr int W_TAP_BYTES = NOUT * 128;
    static constexpr int W_BYTES = NTAPS * W_TAP_BYTES;          // 72 KB (3x3), 120 KB (5x5), 30 KB (last)
    static constexpr int BOX_ROWS = MCN_TY + 2 * R;             // a box carries its dy halo
    static constexpr int A_BYTES = BOX_ROWS * MCN_ROW_BYTES;    // 20 KB (3x3) / 24 KB (5x5)
    static constexpr int STAGES = (R == 2 && NOUT == 64) ? 3 : 6;
    static constexpr int SMEM_W = 0;
    static constexpr int SMEM_A = W_BYTES;
    static constexpr int SMEM_STG = SMEM_A + STAGES * A_BYTES;
    static constexpr int SMEM_BAR = SMEM_STG + 4 * 2 * MCN_STG_BYTES;
    static constexpr int SMEM_TOTAL = SMEM_BAR + 256;
    static constexpr int SMEM_LAUNCH = SMEM_TOTAL + 1024;        // slack for 1024-byte alignment
    static constexpr int TMEM_COLS = 4 * NOUT;                   // four accumulators: 256 / 64 columns
    static_assert(W_BYTES % 1024 == 0, "tap matrices keep the ring 1024-byte aligned");
    static_assert(SMEM_LAUNCH <= 227 * 1024, "shared memory");
};

struct McnConvParams {
    int B, H, W, Wsp;
    int ntx, nty;                 // tiles along x (super-pixels / 16) and y (rows / 8)
    int act;                      // 0 none, 1 LeakyReLU(0.2)
    int prefetch_dist;            // L2 prefetch distance in tiles of this CTA (0 = off)
    int reverse;                  // walk the tiles last-to-first (see rdvc_mcn_forward: L2 reuse between layers)
    int cout;                     // last layer: real output channels (<= 8)
    const __half* residual;       // optional, activation layout; added before the activation
    const float* warped;          // last layer: (B, cout, H, W) fp32 multiplied by the sigmoid
    float* out;                   // last layer: (B, cout, H, W) fp32
    float bias[MCN_C];
};

__device__ __forceinline__ void mcn_sts_16(uint32_t addr, uint4 v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z),
                 "r"(v.w)
                 : "memory");
}
__device__ __forceinline__ uint4 mcn_lds_16(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "r"(addr)
                 : "memory");
    return v;
}
__device__ __forceinline__ uint32_t mcn_pack_h2(float a, float b) {
    __half2 t = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&t);
}

template <int R, int NOUT, int KPAT>
__global__ void __launch_bounds__(MCN_THREADS, 1)
mcn_conv_kernel(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_w,
                const __grid_constant__ CUtensorMap tm_out, const __grid_constant__ CUtensorMap tm_res,
                const McnConvParams p) {
    using Cfg = McnCfg<R, NOUT>;
    constexpr int NTAPS = Cfg::NTAPS, STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    const uint32_t s_w = ptx::smem_u32(smem + Cfg::SMEM_W);
    const uint32_t s_a = ptx::smem_u32(smem + Cfg::SMEM_A);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::SMEM_BAR);
    const uint32_t bar0 = ptx::smem_u32(bars);
    constexpr int A_FULL = 0, A_EMPTY = 6, W_FULL = 12, T_FULL = 13, T_EMPTY = 17, R_FULL = 21;   // 6, 6, 1, 4, 4, 8
    static_assert(STAGES <= 6, "barrier slots");
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
        // ===================== TMA producer =====================
        if (lane == 0) {
            ptx::mbar_arrive_expect_tx(bar(W_FULL), Cfg::W_BYTES);
            for (int t = 0; t < NTAPS; ++t)
                ptx::tma_load_3d(s_w + t * Cfg::W_TAP_BYTES, &tm_w, bar(W_FULL), 0, 0, t);
            // Experiment (option key 13, off by default): pull a tile's boxes into L2 `prefetch_dist` tiles ahead with
            // TMA prefetches (no shared memory needed; the boxes at dsx = -1 and +1 cover the whole footprint).
            // Measured at 1080p: 0.681 ms (off), 0.675-0.683 (1-4 ahead), 0.77 (6+) -- load latency is not what
            // bounds the layer; the L2 -> SM feed (0.51 GB per 3x3 layer at ~7.5 TB/s) is.
            auto prefetch_tile = [&](long long tile) {
                if (tile >= n_tiles) return;
                const int tw = p.reverse ? n_tiles - 1 - static_cast<int>(tile) : static_cast<int>(tile);
                const int b = tw / tiles_per_img, rem = tw % tiles_per_img;
                const int y0 = (rem / p.ntx) * MCN_TY, x0 = (rem % p.ntx) * MCN_TX;
                ptx::tma_prefetch_4d(&tm_in, 0, x0 - 1, y0 - R, b);
                ptx::tma_prefetch_4d(&tm_in, 0, x0 + 1, y0 - R, b);
            };
            if (p.prefetch_dist > 0)
                for (int d = 0; d < p.prefetch_dist; ++d) prefetch_tile(blockIdx.x + static_cast<long long>(d) * gridDim.x);
            uint32_t a_it = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                if (p.prefetch_dist > 0) prefetch_tile(tile + static_cast<long long>(p.prefetch_dist) * gridDim.x);
                const int tw = p.reverse ? n_tiles - 1 - tile : tile;
                const int b = tw / tiles_per_img, rem = tw % tiles_per_img;
                const int y0 = (rem / p.ntx) * MCN_TY, x0 = (rem % p.ntx) * MCN_TX;
                for (int dsx = -1; dsx <= 1; ++dsx, ++a_it) {
                    const uint32_t st = a_it % STAGES, ph = (a_it / STAGES) & 1;
                    ptx::mbar_wait(bar(A_EMPTY + st), ph ^ 1);
                    ptx::mbar_arrive_expect_tx(bar(A_FULL + st), Cfg::A_BYTES);
                    ptx::tma_load_4d(s_a + st * Cfg::A_BYTES, &tm_in, bar(A_FULL + st), 0, x0 + dsx, y0 - R, b);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1 || warp == 3) {
        // Two issuing warps: warp 1 takes the even tiles of this CTA, warp 3 the odd ones, each with two
        // TMEM accumulators (four in all), so the next tile's MMAs do not wait for the epilogue of the last.  An M128 x N64 x K16 MMA occupies the tensor pipe for ~32 cycles but costs one thread ~65
        // cycles to issue, so a single issuer left the pipe two-thirds idle (measured).
        // (With the 3-stage ring of the 5x5 32 -> 32 shape both parities would share every ring slot and a
        // parity wait could alias across the other warp's phase: that shape keeps one issuer.  With 6 stages
        // even tiles own slots 0-2 and odd tiles slots 3-5.)
        constexpr uint32_t ISSUERS = (STAGES == 6) ? 2 : 1;
        const uint32_t mma_id = warp >> 1;
        // ===================== MMA issuer =====================
        // The WHOLE warp runs the loops and the barrier waits; only the tcgen05 instructions sit under
        // elect_one().  With `if (lane == 0)` around everything the compiler keeps the descriptors in per-thread
        // registers and wraps every MMA in an elect / R2UR.BROADCAST / branch sequence (~9 instructions, ~65
        // cycles per MMA for a lone thread); in converged code they live in uniform registers and the MMAs of a
        // box are issued back to back, one instruction each (checked with cuobjdump).
        {
            const uint32_t idesc = ptx::umma_idesc(128, NOUT, 0);   // fp16 operands, fp32 accumulate
            ptx::mbar_wait(bar(W_FULL), 0);
            ptx::tc_fence_after();
            const uint64_t b_desc0 = ptx::umma_desc_k_sw128(s_w);
            for (uint32_t tile_it = mma_id; mma_id < ISSUERS && blockIdx.x + static_cast<long long>(tile_it) * gridDim.x < n_tiles;
                 tile_it += ISSUERS) {
                const uint32_t acc = tile_it & 3, acc_ph = (tile_it >> 2) & 1;   // four accumulators, two per issuer
                ptx::mbar_wait(bar(T_EMPTY + acc), acc_ph ^ 1);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * NOUT;
                uint32_t a_it = 3 * tile_it;
#pragma unroll
                for (int dsx = 0; dsx < 3; ++dsx, ++a_it) {
                    const uint32_t st = a_it % STAGES, ph = (a_it / STAGES) & 1;
                    ptx::mbar_wait(bar(A_FULL + st), ph);
                    ptx::tc_fence_after();
                    const uint64_t a_desc0 = ptx::umma_desc_k_sw128(s_a + st * Cfg::A_BYTES);
                    if (ptx::elect_one()) {
#pragma unroll
                    for (int dy = 0; dy < 2 * R + 1; ++dy) {
                        // the tap's A operand is the box shifted down by dy rows: 2 KB = two whole swizzle
                        // atoms, so the 128-byte swizzle phase of every row is unchanged.  Descriptor
                        // start addresses are in 16-byte units in the low bits: offsets are plain adds.
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            if (mcn_kstep_on(KPAT, dsx, k)) {
                                constexpr int first_k = (KPAT == MCN_K_3X3) ? 2 : 0;   // first step issued at dsx = 0, dy = 0
                                const bool first = (dsx == 0 && dy == 0 && k == first_k);
                                ptx::umma_bf16(d_tmem, a_desc0 + ((dy * MCN_ROW_BYTES + k * 32) >> 4),
                                               b_desc0 + (((dy * 3 + dsx) * Cfg::W_TAP_BYTES + k * 32) >> 4), idesc,
                                               first ? 0u : 1u);
                            }
                        }
                    }
                    ptx::umma_commit(bar(A_EMPTY + st));   // ring slot free when these MMAs retire
                    if (dsx == 2) ptx::umma_commit(bar(T_FULL + acc));
                    }
                    __syncwarp();
                }
            }
        }
        __syncwarp();
    } else if (warp >= 4) {
        // ===================== epilogue =====================
        // Eight warps: warp 4 + e reads TMEM lane quarter q = e % 4 (rows 32 q .. 32 q + 31 of a tile) and
        // takes the tiles of parity e / 4, i.e. one accumulator -- two warps per scheduler hide each other's
        // latencies (with four warps the epilogue was the second bottleneck, ~1500 cycles per tile).
        const int e = warp - 4;
        const int q = e & 3;
        const uint32_t par = e >> 2;
        const int ty = q * 2 + (lane >> 4), tx = lane & 15;
        const uint32_t sb = ptx::smem_u32(smem + Cfg::SMEM_STG) + e * MCN_STG_BYTES;   // this warp's staging box
        for (uint32_t tile_it = par; blockIdx.x + static_cast<long long>(tile_it) * gridDim.x < n_tiles; tile_it += 2) {
            const int tile = blockIdx.x + tile_it * gridDim.x;
            const int tw = p.reverse ? n_tiles - 1 - tile : tile;
            const int b = tw / tiles_per_img, rem = tw % tiles_per_img;
            const int y0 = (rem / p.ntx) * MCN_TY, x0 = (rem % p.ntx) * MCN_TX;
            const int y = y0 + ty, sp = x0 + tx;
            const bool inside = (y < p.H) && (sp < p.Wsp);
            const uint32_t acc = tile_it & 3, acc_ph = (tile_it >> 2) & 1;
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * NOUT;

            if constexpr (NOUT == 64) {
                // The residual rows of this warp (the same 4 KB box the result goes out through) are fetched
                // by TMA into the staging buffer before the accumulator is waited for: per-thread 128-byte
                // global reads here cost 35 us per layer at 1080p.
                const bool use_res = (p.residual != nullptr);
                if (lane == 0) {
                    ptx::bulk_wait_read<0>();   // this warp's previous store has finished reading `sb`
                    if (use_res) {
                        ptx::mbar_arrive_expect_tx(bar(R_FULL + e), MCN_STG_BYTES);
                        ptx::tma_load_4d(sb, &tm_res, bar(R_FULL + e), 0, x0, y0 + 2 * q, b);
                    }
                }
                __syncwarp();
                ptx::mbar_wait(bar(T_FULL + acc), acc_ph);
                ptx::tc_fence_after();
                float v[64];
#pragma unroll
                for (int k = 0; k < 4; ++k) ptx::tmem_ld_x16(taddr + k * 16, v + k * 16);
                ptx::tmem_ld_wait();
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(bar(T_EMPTY + acc));   // accumulator back to the MMA warp
#pragma unroll
                for (int i = 0; i < 64; ++i) v[i] += p.bias[i & 31];
                if (use_res) {
                    ptx::mbar_wait(bar(R_FULL + e), (tile_it >> 1) & 1);
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const uint4 r4 = mcn_lds_16(sb + lane * 128 + ((c ^ (lane & 7)) << 4));
                        const uint32_t w4[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w4[j]));
                            v[c * 8 + 2 * j] += f.x;
                            v[c * 8 + 2 * j + 1] += f.y;
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
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    uint4 w;
                    w.x = mcn_pack_h2(v[c * 8 + 0], v[c * 8 + 1]);
                    w.y = mcn_pack_h2(v[c * 8 + 2], v[c * 8 + 3]);
                    w.z = mcn_pack_h2(v[c * 8 + 4], v[c * 8 + 5]);
                    w.w = mcn_pack_h2(v[c * 8 + 6], v[c * 8 + 7]);
                    mcn_sts_16(sb + lane * 128 + ((c ^ (lane & 7)) << 4), w);   // TMA's SWIZZLE_128B
                }
                ptx::fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    ptx::tma_store_4d(&tm_out, sb, 0, x0, y0 + 2 * q, b);   // clipped to the tensor
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
                float v[16];
                ptx::tmem_ld_x16(taddr, v);
                ptx::tmem_ld_wait();
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(bar(T_EMPTY + acc));
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

// The network's input: concat(warped_ref (3), flow (2), ref_frame (3)) -- R:codec_processing.py:402 --
// NCHW fp32 -> activation layout (fp16, channels 8..31 and the odd-width padding pixel zero).
struct McnInputParams {
    const float* src[3];
    int ch[3];          // channels of each source (3, 2, 3)
    __half* dst;
    int B, H, W, Wsp;
};

__global__ void __launch_bounds__(128) mcn_pack_input_kernel(const McnInputParams p) {
    const int x = blockIdx.x * 128 + threadIdx.x;
    const int y = blockIdx.y, b = blockIdx.z;
    if (x >= 2 * p.Wsp) return;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = 0.f;
    if (x < p.W) {
        const size_t plane = static_cast<size_t>(p.H) * p.W;
        const size_t pix = static_cast<size_t>(y) * p.W + x;
        const int c1 = p.ch[0], c2 = c1 + p.ch[1], c3 = c2 + p.ch[2];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            if (c < c3) {
                const int s = (c < c1) ? 0 : (c < c2) ? 1 : 2;
                const int k = (c < c1) ? c : (c < c2) ? c - c1 : c - c2;
                const float* src = (s == 0) ? p.src[0] : (s == 1) ? p.src[1] : p.src[2];
                const int chs = (s == 0) ? p.ch[0] : (s == 1) ? p.ch[1] : p.ch[2];
                v[c] = __ldg(src + (static_cast<size_t>(b) * chs + k) * plane + pix);
            }
        }
    }
    uint4 w0;
    w0.x = mcn_pack_h2(v[0], v[1]);
    w0.y = mcn_pack_h2(v[2], v[3]);
    w0.z = mcn_pack_h2(v[4], v[5]);
    w0.w = mcn_pack_h2(v[6], v[7]);
    uint4* d = reinterpret_cast<uint4*>(p.dst + ((static_cast<size_t>(b) * p.H + y) * (2 * p.Wsp) + x) * MCN_C);
    const uint4 z = make_uint4(0, 0, 0, 0);
    d[0] = w0; d[1] = z; d[2] = z; d[3] = z;
}

}  // namespace rdvc
