// corr_encoder_tail_sm100.cuh -- "next" row f-2, last sub-item: the feature encoder's final 1x1 convolution
// (FeatureEncoder.conv = Conv2d(128, 256, kernel_size=1), TV:raft.py:139 construction, :150 call) computed straight
// into the K-major 16-bit operand rows the correlation build multiplies.
//
// Stock order of work: conv (cuDNN: reads 33 MB, writes the (2B, 256, h, w) fp32 maps, 67 MB at 1080p) ->
// corr_pack_kernel (reads the 67 MB, transposes, pools, rounds to bf16) -> build.  A 1x1 convolution commutes with
// the transpose and, being linear, with the 2^l x 2^l mean pooling (the bias included: a mean of (W x + b) is
// W mean(x) + b).  So: corr_pack_kernel runs on the encoder's 128-channel activation instead (half the bytes in,
// half the bytes out: K-major rows [n][128] of image 1 and of image 2 at every pyramid level, in the pyramid's row
// order), and THIS kernel turns every packed row into the build's operand row:
//
//     out[r][n] = round16( sum_k in[r][k] * W[n][k] + bias[n] * valid(r) ),    n < 256, k < 128
//
// valid(r) = 0 for the layout-padding rows of a level (they must stay exact zeros: the volume's padding pixels are
// what the lookup's zero padding reads), 1 otherwise.  The fp32 feature maps never exist.
//
// One launch covers all five row segments (image 1; image 2 at levels 0..3): the packed input and the operand
// output are single buffers of 256- / 512-byte rows, so ONE 2-D tensor map addresses every input tile by its row
// number.  Persistent CTAs, one per SM: the whole weight matrix (256 x 128 x 2 B = 64 KB, two 128B-swizzled
// k-blocks) stays in shared memory; 128-row tiles stream through a 4-stage 16 KB ring; tcgen05.mma M128 x N256 x
// K16, fp32 accumulators double-buffered in TMEM; sixteen epilogue warps (the epilogue is a latency chain -- TMEM
// load, convert, shared-memory transpose, store -- so it wants warps, not instructions: 8 warps x 2 passes measured
// 20.5 us per launch under ncu): tcgen05.ld -> + bias -> 16-bit pack -> swizzled 4 KB shared-memory transpose per
// warp -> 16-byte stores that cover whole 128-byte lines of the rows.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cstdint>

#include "corr_conv1x1_sm100.cuh"   // sts/lds helpers
#include "ptx_sm100.cuh"

namespace rdvc {

constexpr int ET_BLOCK_M = 128;
constexpr int ET_K = 128;                        // input channels (FeatureEncoder layers[3])
constexpr int ET_N = 256;                        // output channels = the build's D
constexpr int ET_KB = ET_K / 64;                 // 2 k-blocks
constexpr int ET_A_STAGES = 4;
constexpr int ET_A_STAGE_BYTES = ET_BLOCK_M * 64 * 2;     // 16 KB
constexpr int ET_W_SLAB_BYTES = ET_N * 64 * 2;            // 32 KB
constexpr int ET_EPI_WARPS = 16;                // four per TMEM lane quarter, 64 output channels each
constexpr int ET_THREADS = 128 + ET_EPI_WARPS * 32;
constexpr int ET_STG_BYTES = 32 * 128;                    // per epilogue warp: 32 rows x 64 channels x 2 B
constexpr int ET_SMEM_W = 0;
constexpr int ET_SMEM_A = ET_SMEM_W + ET_KB * ET_W_SLAB_BYTES;            // 65536
constexpr int ET_SMEM_STG = ET_SMEM_A + ET_A_STAGES * ET_A_STAGE_BYTES;   // 131072
constexpr int ET_SMEM_BIAS = ET_SMEM_STG + ET_EPI_WARPS * ET_STG_BYTES;   // 163840
constexpr int ET_SMEM_BAR = ET_SMEM_BIAS + ET_N * 4;
constexpr int ET_SMEM_TOTAL = ET_SMEM_BAR + 256;
constexpr int ET_SMEM_LAUNCH = ET_SMEM_TOTAL + 1024;
constexpr int ET_MAX_SEGS = 5;

struct EncoderTailParams {
    // segment s: `rows[s]` operand rows starting at row `in_row0[s]` of the packed input buffer (256-byte rows) and at
    // row `out_row0[s]` of the output buffer (512-byte rows); `tile0[s]` = first 128-row tile of the segment
    long long in_row0[ET_MAX_SEGS], out_row0[ET_MAX_SEGS], rows[ET_MAX_SEGS];
    int tile0[ET_MAX_SEGS + 1];
    // validity of a row inside a level image (padding rows get no bias): per segment, image rows `img[s]` (rows of one
    // batch item), level size and tiling; tiles_w == 0 means "every row is a real pixel"
    int img[ET_MAX_SEGS], hl[ET_MAX_SEGS], wl[ET_MAX_SEGS], tiles_w[ET_MAX_SEGS];
    int twl, thl;
    int n_segs, n_tiles;
    void* out;              // [rows][256] 16-bit
    const float* bias;      // 256 floats on the device, or nullptr
    int ab_format;          // 1 = bf16, 0 = fp16 (operands and output)
    unsigned long long* dbg_timeline;   // RDVC_EXPERIMENTS builds only: 16 globaltimer stamps per CTA (nullptr = off)
};

#ifdef RDVC_EXPERIMENTS
#define ET_STAMP(slot) do { if (p.dbg_timeline) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_) :: "memory"); \
                                                   p.dbg_timeline[blockIdx.x * 16 + (slot)] = t_; } } while (0)
#else
#define ET_STAMP(slot) ((void)0)
#endif

// grid: min(SMs, n_tiles) CTAs; block: ET_THREADS; dynamic smem: ET_SMEM_LAUNCH.
// tm_a: the packed input as a {128, total rows, 1} tensor, box {64, 128, 1}; tm_w: weights {128, 256, 1}, box {64, 256, 1}.
__global__ void __launch_bounds__(ET_THREADS, 1)
corr_encoder_tail_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_w,
                         const EncoderTailParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    const uint32_t s_w = ptx::smem_u32(smem + ET_SMEM_W);
    const uint32_t s_a = ptx::smem_u32(smem + ET_SMEM_A);
    const uint32_t s_stg = ptx::smem_u32(smem + ET_SMEM_STG);
    const uint32_t s_bias = ptx::smem_u32(smem + ET_SMEM_BIAS);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + ET_SMEM_BAR);
    const uint32_t bar0 = ptx::smem_u32(bars);
    constexpr int A_FULL = 0, A_EMPTY = ET_A_STAGES, W_FULL = 2 * ET_A_STAGES, T_FULL = W_FULL + 1, T_EMPTY = T_FULL + 2;
    auto bar = [&](int i) { return bar0 + 8u * i; };
    const uint32_t tmem_slot_addr = bar0 + 8u * (T_EMPTY + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        ET_STAMP(0);                                     // entry
        ptx::prefetch_tensormap(&tm_a);
        ptx::prefetch_tensormap(&tm_w);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < ET_A_STAGES; ++i) {
            ptx::mbar_init(bar(A_FULL + i), 1);
            ptx::mbar_init(bar(A_EMPTY + i), 1);
        }
        ptx::mbar_init(bar(W_FULL), 1);
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(bar(T_FULL + i), 1);
            ptx::mbar_init(bar(T_EMPTY + i), ET_EPI_WARPS);
        }
        ptx::fence_mbar_init();
    }
    if (warp == 2) {
        ptx::tmem_alloc(tmem_slot_addr, 512);
        ptx::tmem_relinquish();
    }
    if (warp >= 4 && threadIdx.x - 128 < ET_N)          // one element per thread: ONE global round trip (a single warp
        sts_f32(s_bias + (threadIdx.x - 128) * 4, p.bias ? __ldg(p.bias + (threadIdx.x - 128)) : 0.f);   // looping cost 2 us)
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot_addr) : "memory");
    if (warp == 0 && lane == 0) ET_STAMP(1);             // set-up done

    auto seg_of = [&](int tile) {
        int s = 0;
        while (s + 1 < p.n_segs && tile >= p.tile0[s + 1]) ++s;
        return s;
    };

    if (warp == 0) {
        // ===================== TMA producer (converged warp, instructions under elect_one) =====================
        if (ptx::elect_one()) {
            ptx::mbar_arrive_expect_tx(bar(W_FULL), ET_KB * ET_W_SLAB_BYTES);
            for (int kb = 0; kb < ET_KB; ++kb)
                ptx::tma_load_3d(s_w + kb * ET_W_SLAB_BYTES, &tm_w, bar(W_FULL), kb * 64, 0, 0);
        }
        __syncwarp();
        uint32_t a_it = 0;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
            const int s = seg_of(tile);
            const long long r0 = p.in_row0[s] + static_cast<long long>(tile - p.tile0[s]) * ET_BLOCK_M;
            for (int kb = 0; kb < ET_KB; ++kb, ++a_it) {
                const uint32_t st = a_it % ET_A_STAGES, ph = (a_it / ET_A_STAGES) & 1;
                ptx::mbar_wait(bar(A_EMPTY + st), ph ^ 1);
                if (ptx::elect_one()) {
                    ptx::mbar_arrive_expect_tx(bar(A_FULL + st), ET_A_STAGE_BYTES);
                    ptx::tma_load_3d(s_a + st * ET_A_STAGE_BYTES, &tm_a, bar(A_FULL + st), kb * 64, static_cast<int>(r0), 0);
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (converged warp) =====================
        const uint32_t idesc = ptx::umma_idesc(ET_BLOCK_M, ET_N, p.ab_format);
        ptx::mbar_wait(bar(W_FULL), 0);
        ptx::tc_fence_after();
        if (lane == 0) ET_STAMP(2);                      // weights landed
        const uint64_t b_desc0 = ptx::umma_desc_k_sw128(s_w);
        uint32_t a_it = 0, tile_it = 0;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++tile_it) {
            const uint32_t acc = tile_it & 1, acc_ph = (tile_it >> 1) & 1;
            ptx::mbar_wait(bar(T_EMPTY + acc), acc_ph ^ 1);
            ptx::tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * ET_N;
#pragma unroll
            for (int kb = 0; kb < ET_KB; ++kb, ++a_it) {
                const uint32_t st = a_it % ET_A_STAGES, ph = (a_it / ET_A_STAGES) & 1;
                ptx::mbar_wait(bar(A_FULL + st), ph);
                ptx::tc_fence_after();
                if (lane == 0 && kb == ET_KB - 1 && tile_it < 5) ET_STAMP(3 + tile_it);    // tile's operands landed (3..7)
                const uint64_t a_desc0 = ptx::umma_desc_k_sw128(s_a + st * ET_A_STAGE_BYTES);
                const uint64_t b_desc = b_desc0 + ((kb * ET_W_SLAB_BYTES) >> 4);
                if (ptx::elect_one()) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        ptx::umma_bf16(d_tmem, a_desc0 + ((k * 32) >> 4), b_desc + ((k * 32) >> 4), idesc, (kb | k) != 0 ? 1u : 0u);
                    ptx::umma_commit(bar(A_EMPTY + st));
                    if (kb == ET_KB - 1) ptx::umma_commit(bar(T_FULL + acc));
                }
                __syncwarp();
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue =====================
        // warp 4 + e: TMEM lane quarter q = e % 4 (rows 32 q .. 32 q + 31 of the tile), channel quarter e / 4 (64 channels,
        // one pass).  Out of TMEM a thread is one row; its 64 channels (128 bytes of 16-bit values) go
        // to row `lane` of the warp's buffer with the 16-byte chunk index XORed by (lane & 7) (conflict-free both
        // ways); then lane l re-reads chunk l % 8 of rows 4 k + l / 8 and stores it: a warp-wide store = four whole
        // 128-byte lines of four consecutive output rows.
        const int e = warp - 4;
        const int q = e & 3;
        const int ch_half = (e >> 2) * (ET_N / 4);        // first channel of this warp's quarter
        const uint32_t stg = s_stg + e * ET_STG_BYTES;
        uint8_t* const out = static_cast<uint8_t*>(p.out);
        uint32_t tile_it = 0;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++tile_it) {
            const int s = seg_of(tile);
            const long long rseg0 = static_cast<long long>(tile - p.tile0[s]) * ET_BLOCK_M + q * 32;   // first row of this warp in the segment
            // validity of MY row (the one my TMEM lane holds): bias only for real pixels
            const long long rme = rseg0 + lane;
            float bias_on = 1.f;
            if (p.tiles_w[s] > 0) {
                const int r_img = static_cast<int>(static_cast<unsigned>(rme) % static_cast<unsigned>(p.img[s]));   // rows < 2^31 (ABI check)
                const int tl = r_img >> (p.twl + p.thl), in = r_img & ((1 << (p.twl + p.thl)) - 1);
                const int ty = tl / p.tiles_w[s], tx = tl - ty * p.tiles_w[s];
                const int y = (ty << p.thl) + (in >> p.twl), x = (tx << p.twl) + (in & ((1 << p.twl) - 1));
                bias_on = (y < p.hl[s] && x < p.wl[s]) ? 1.f : 0.f;
            }
            const uint32_t acc = tile_it & 1, acc_ph = (tile_it >> 1) & 1;
            ptx::mbar_wait(bar(T_FULL + acc), acc_ph);
            ptx::tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * ET_N + ch_half;
            if (e == 0 && lane == 0 && tile_it < 4) ET_STAMP(8 + 2 * tile_it);            // accumulator ready (8, 10, 12, 14)
#pragma unroll
            for (int pass = 0; pass < 1; ++pass) {
                float v[64];
#pragma unroll
                for (int k = 0; k < 4; ++k) ptx::tmem_ld_x16(taddr + pass * 64 + k * 16, v + k * 16);
                ptx::tmem_ld_wait();
                if (pass == 0) {
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive_relaxed(bar(T_EMPTY + acc));   // global stores are in flight: no MEMBAR
                }
                uint32_t pk[32];
#pragma unroll
                for (int i4 = 0; i4 < 16; ++i4) {               // bias: 16 vector loads the compiler is free to batch
                    const float4 bb = lds_ro_f32x4(s_bias + (ch_half + pass * 64 + 4 * i4) * 4);
                    v[4 * i4] += bb.x * bias_on; v[4 * i4 + 1] += bb.y * bias_on;
                    v[4 * i4 + 2] += bb.z * bias_on; v[4 * i4 + 3] += bb.w * bias_on;
                }
                if (p.ab_format) {                               // one uniform branch, not 64 predicated clamps
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const __nv_bfloat162 t2 = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
                        pk[i] = *reinterpret_cast<const uint32_t*>(&t2);
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const __half2 t2 = __floats2half2_rn(fminf(fmaxf(v[2 * i], -65504.f), 65504.f),
                                                             fminf(fmaxf(v[2 * i + 1], -65504.f), 65504.f));
                        pk[i] = *reinterpret_cast<const uint32_t*>(&t2);
                    }
                }
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg + lane * 128 + ((c ^ (lane & 7)) << 4)),
                                 "r"(pk[4 * c]), "r"(pk[4 * c + 1]), "r"(pk[4 * c + 2]), "r"(pk[4 * c + 3]) : "memory");
                __syncwarp();
                const int cch = lane & 7, rsub = lane >> 3;
                // one 64-bit base per tile; a store is base + r * 512 under a 32-bit row bound
                uint8_t* const obase = out + (p.out_row0[s] + rseg0) * (ET_N * 2) + (ch_half + pass * 64) * 2 + cch * 16;
                const long long left = p.rows[s] - rseg0;
                const int nvalid = left > 32 ? 32 : (left < 0 ? 0 : static_cast<int>(left));
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int r = 4 * k + rsub;
                    uint4 w;
                    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w.x), "=r"(w.y), "=r"(w.z), "=r"(w.w)
                                 : "r"(stg + r * 128 + ((cch ^ (r & 7)) << 4)) : "memory");
                    if (r < nvalid) *reinterpret_cast<uint4*>(obase + r * (ET_N * 2)) = w;
                }
                __syncwarp();
            }
            if (e == 0 && lane == 0 && tile_it < 4) ET_STAMP(9 + 2 * tile_it);            // this warp's stores issued (9, 11, 13, 15)
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 2) ptx::tmem_dealloc(tmem_base, 512);
}

}  // namespace rdvc
