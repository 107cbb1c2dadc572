// wbench3.cu -- developer microbenchmark (round 1c): TMA TENSOR stores of level 0 of the 1080p volume
// ([32640 x 32640] fp32, 4.26 GB) from shared memory, box shapes compared.  The global tensor is described as
// {32 floats (128 B), CB column blocks, rows}: a box {32, cb, r} writes r rows x (cb * 128) contiguous bytes,
// traversed column-block-fastest, so cb > 1 makes every row visit cb*128 bytes wide.
// CTA = 8 warps, tile = 128 rows x 256 columns as in corr_build; warp (q, sub) owns 32 rows x 128 columns.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o wbench3 wbench3.cu -lcuda
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

// CB column blocks x RB rows per box; NBUF boxes in flight per warp
template <int CB, int RB, int NBUF>
__global__ void __launch_bounds__(256) wk(const __grid_constant__ CUtensorMap tm, int ntiles, int mblks, int order) {
    extern __shared__ __align__(1024) uint8_t smem[];
    constexpr int BOX = CB * RB * 128;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = warp & 3, sub = warp >> 2;
    uint8_t* my = smem + warp * (BOX * NBUF);
    const long long total = (long long)ntiles * mblks;
    uint32_t it = 0;
    for (long long t = blockIdx.x; t < total; t += gridDim.x) {
        int nt, mb;
        if (order == 0) { nt = (int)(t % ntiles); mb = (int)(t / ntiles); }
        else { nt = (int)(t / mblks); mb = (int)(t % mblks); }
        const int row0 = mb * 128 + q * 32;
        const int cb0 = nt * 8 + sub * 4;                 // first 128-byte column block of this warp's half
        for (int c = 0; c < 4; c += CB)
            for (int r = 0; r < 32; r += RB) {
                uint8_t* buf = my + (it % NBUF) * BOX;
                ++it;
                if (lane == 0) bulk_wait_read<NBUF - 1>();
                __syncwarp();
                // stage BOX bytes with 16-byte stores (pattern irrelevant for the write-side question)
                for (int o = lane * 16; o < BOX; o += 32 * 16) *reinterpret_cast<float4*>(buf + o) = make_float4(1.f, 2.f, 3.f, 4.f);
                fence_async();
                __syncwarp();
                if (lane == 0) {
                    tma_store_3d(&tm, smem_u32(buf), 0, cb0 + c, row0 + r);
                    bulk_commit();
                }
            }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int CB, int RB, int NBUF>
void run(EncodeFn enc, float* out, long long rows, long long cols, int order, long long pitch = 0) {
    if (!pitch) pitch = cols;
    CUtensorMap tm;
    cuuint64_t dims[3] = {32, (cuuint64_t)(cols / 32), (cuuint64_t)rows};
    cuuint64_t str[2] = {128, (cuuint64_t)pitch * 4};
    cuuint32_t box[3] = {32, CB, RB};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, out, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return; }
    const int ntiles = (int)(cols / 256), mblks = (int)(rows / 128);
    const int smem = 8 * CB * RB * 128 * NBUF + 1024;
    cudaFuncSetAttribute(wk<CB, RB, NBUF>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; ++rep) wk<CB, RB, NBUF><<<148, 256, smem>>>(tm, ntiles, mblks, order);
    cudaEventRecord(e0);
    for (int rep = 0; rep < 5; ++rep) wk<CB, RB, NBUF><<<148, 256, smem>>>(tm, ntiles, mblks, order);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
    printf("pitch %lld order %d box %4d B x %2d rows (%5d B) nbuf %d smem %6d: %.3f ms  %.0f GB/s  (%s)\n", pitch, order, CB * 128, RB,
           CB * RB * 128, NBUF, smem, ms, (double)ntiles * 256 * mblks * 128 * 4 / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    const long long rows = 32640, cols = 32640;
    float* out;
    cudaMalloc(&out, rows * 34816LL * 4);
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    EncodeFn enc = (EncodeFn)p;
    for (long long pitch : {32640LL, 32768LL, 32704LL, 33024LL, 32672LL, 34816LL}) {
        run<2, 16, 1>(enc, out, rows, cols, 0, pitch);
        run<1, 32, 1>(enc, out, rows, cols, 0, pitch);
    }
    return 0;
}
