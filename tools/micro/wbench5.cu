// wbench5.cu -- developer microbenchmark (round 1e): can the build's TMA stores form ONE sequential write front?
// A non-persistent elementwise kernel writes at 7.5 TB/s because the CTA dispatch order makes all SMs write one
// narrow, address-ordered front (wbench4 mode 3); the build's tile order reaches 6.1 TB/s (wbench3).  Here the same
// persistent CTAs, tiles and 4 KB boxes (16 rows x 256 B) as the build kernel write into a BOX-INTERLEAVED layout
//   [m-block][step (g, hh)][warp e][n-tile][16 rows][256 B]
// so that the boxes all SMs write at about the same time are adjacent in memory.
//   layout 0: row-major volume rows (what the build writes today)
//   layout 1: box-interleaved, n-tile fastest
//   layout 2: box-interleaved, [m-block][step][n-tile][warp e]
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o wbench5 wbench5.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, uint32_t src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(m), "r"(src),
                 "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__global__ void __launch_bounds__(256) wk(const __grid_constant__ CUtensorMap tm, int ntiles, int mblks, int layout) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = warp & 3, sub = warp >> 2;
    uint8_t* buf = smem + warp * 4096;
    const long long total = (long long)ntiles * mblks;
    for (long long t = blockIdx.x; t < total; t += gridDim.x) {
        const int nt = (int)(t % ntiles), mb = (int)(t / ntiles);     // all CTAs on the same m-block
        for (int g = 0; g < 2; ++g)
            for (int hh = 0; hh < 2; ++hh) {
                if (lane == 0) bulk_wait_read<0>();
                __syncwarp();
                for (int o = lane * 16; o < 4096; o += 32 * 16) *reinterpret_cast<float4*>(buf + o) = make_float4(1.f, 2.f, 3.f, 4.f);
                fence_async();
                __syncwarp();
                if (lane == 0) {
                    if (layout == 0) {
                        // {32 floats, column block, row}: row-major rows of 32640 floats
                        tma_store_3d(&tm, smem_u32(buf), 0, nt * 8 + sub * 4 + g * 2, mb * 128 + q * 32 + hh * 16);
                    } else {
                        const int step = g * 2 + hh;
                        long long box = (layout == 1) ? ((((long long)mb * 4 + step) * 8 + warp) * ntiles + nt)
                                                      : ((((long long)mb * 4 + step) * ntiles + nt) * 8 + warp);
                        tma_store_3d(&tm, smem_u32(buf), 0, 0, (int)(box * 16));
                    }
                    bulk_commit();
                }
            }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    const long long rows = 32640, cols = 32640;
    const int ntiles = 128, mblks = 255;                       // 127.5 -> 128 tiles in the interleaved layouts
    float* out;
    cudaMalloc(&out, (size_t)mblks * 128 * ntiles * 256 * 4);
    void* p = nullptr; cudaDriverEntryPointQueryResult qr;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr);
    EncodeFn enc = (EncodeFn)p;
    cudaFuncSetAttribute(wk, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 4096 + 1024);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int layout = 0; layout < 3; ++layout) {
        CUtensorMap tm;
        cuuint32_t es[3] = {1, 1, 1}, box[3] = {32, 2, 16};
        CUresult r;
        if (layout == 0) {
            cuuint64_t dims[3] = {32, (cuuint64_t)(cols / 32), (cuuint64_t)rows};
            cuuint64_t str[2] = {128, (cuuint64_t)cols * 4};
            r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, out, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        } else {
            cuuint64_t dims[3] = {32, 2, (cuuint64_t)mblks * 32 * ntiles * 16};    // every box = 16 dense rows of 256 B
            cuuint64_t str[2] = {128, 256};
            r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, out, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        }
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
        const int nt_used = layout == 0 ? 127 : ntiles;
        for (int rep = 0; rep < 2; ++rep) wk<<<148, 256, 8 * 4096 + 1024>>>(tm, nt_used, mblks, layout);
        cudaEventRecord(e0);
        for (int rep = 0; rep < 5; ++rep) wk<<<148, 256, 8 * 4096 + 1024>>>(tm, nt_used, mblks, layout);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
        printf("layout %d: %.3f ms  %.0f GB/s  (%s)\n", layout, ms, (double)nt_used * 256 * mblks * 128 * 4 / ms / 1e6,
               cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
