// wbench2.cu -- developer microbenchmark (round 1b): write a [32640 x 32640] fp32 matrix (level 0 of the
// 1080p volume, 4.26 GB) the way the build kernel's epilogue could, from SHARED memory through the TMA engine:
//   mode 0: per-lane 1-D bulk copies (cp.async.bulk.global.shared::cta), one image row piece of W bytes per lane
//   mode 1: plain st.global.v4 with a warp covering `W` contiguous bytes per row (reference)
// CTA = 8 warps; CTA owns a 256-column tile (1 KB per row) and sweeps 128-row blocks like the real kernel;
// warp (q, sub): rows q*32..+31; with W = 512 `sub` picks the half, with W = 1024 sub picks alternate m-blocks,
// with W = 256 / 128 a warp loops over the pieces of its half.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o wbench2 wbench2.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <cstdint>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <int W, int NBUF>
__global__ void __launch_bounds__(256) wk(float* out, long long rows, long long cols, int ntiles, int mblks, int mode,
                                          int order) {
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr int PITCH = W + 16;                       // conflict-free for lane = row writes
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = warp & 3, sub = warp >> 2;
    uint8_t* my = smem + warp * (32 * PITCH * NBUF);
    const long long total = (long long)ntiles * mblks;
    uint32_t it = 0;
    for (long long t = blockIdx.x; t < total; t += gridDim.x) {
        int nt, mb;
        if (order == 0) { nt = (int)(t % ntiles); mb = (int)(t / ntiles); }     // all CTAs on the same rows
        else { nt = (int)(t / mblks); mb = (int)(t % mblks); }                  // CTA sweeps rows of its tile
        constexpr int PIECES = 512 / W > 0 ? 512 / W : 1;  // pieces of this warp's 512-byte half
        if (W == 1024 && ((mb & 1) != sub)) continue;
        const long long row = (long long)mb * 128 + q * 32 + lane;
        for (int pc = 0; pc < PIECES; ++pc) {
            const long long col = (long long)nt * 256 + (W == 1024 ? 0 : sub * 128 + pc * (W / 4));
            if (mode == 0) {
                uint8_t* buf = my + (it % NBUF) * (32 * PITCH);
                ++it;
                bulk_wait_read<NBUF - 1>();            // per-lane groups: this lane's copy NBUF ago has left smem
                // stage: lane = row writes W bytes of its row
                float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
#pragma unroll
                for (int c = 0; c < W / 16; ++c) *reinterpret_cast<float4*>(buf + lane * PITCH + c * 16) = v;
                fence_async();
                if (row < rows && col + W / 4 <= cols) bulk_s2g(out + row * cols + col, smem_u32(buf + lane * PITCH), W);
                bulk_commit();
            } else {
                // plain stores: the warp writes its 32 rows, W bytes each, W/16 lanes per row
                constexpr int LPR = W / 16 > 32 ? 32 : W / 16;  // lanes per row
                constexpr int RPI = 32 / LPR;                    // rows per instruction
                for (int r0 = 0; r0 < 32; r0 += RPI)
                    for (int c0 = 0; c0 < W / 16; c0 += LPR) {
                        const long long rr = (long long)mb * 128 + q * 32 + r0 + lane / LPR;
                        const long long cc = col + (c0 + lane % LPR) * 4;
                        if (rr < rows && cc < cols) {
                            float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
                            asm volatile("st.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(out + rr * cols + cc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
                        }
                    }
            }
        }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <int W, int NBUF>
void run(float* out, long long rows, long long cols, int mode, int order) {
    const int ntiles = (int)(cols / 256), mblks = (int)(rows / 128);
    const int smem = 8 * 32 * (W + 16) * NBUF;
    cudaFuncSetAttribute(wk<W, NBUF>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; ++rep) wk<W, NBUF><<<148, 256, smem>>>(out, rows, cols, ntiles, mblks, mode, order);
    cudaEventRecord(e0);
    for (int rep = 0; rep < 5; ++rep) wk<W, NBUF><<<148, 256, smem>>>(out, rows, cols, ntiles, mblks, mode, order);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
    printf("mode %d (%s) order %d W %4d nbuf %d smem %6d: %.3f ms  %.0f GB/s  (%s)\n", mode, mode ? "st.global" : "bulk s2g",
           order, W, NBUF, smem, ms, (double)ntiles * 256 * mblks * 128 * 4 / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    const long long rows = 32640, cols = 32640;
    float* out;
    cudaMalloc(&out, rows * cols * 4);
    for (int order = 0; order < 2; ++order) {
        run<128, 2>(out, rows, cols, 0, order);
        run<256, 2>(out, rows, cols, 0, order);
        run<512, 1>(out, rows, cols, 0, order);
        run<512, 2>(out, rows, cols, 0, order);
        run<1024, 1>(out, rows, cols, 0, order);
        run<128, 1>(out, rows, cols, 1, order);
        run<512, 1>(out, rows, cols, 1, order);
        run<1024, 1>(out, rows, cols, 1, order);
    }
    return 0;
}
