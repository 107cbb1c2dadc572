// wbench.cu -- developer microbenchmark: how fast can 148 CTAs write a [32640 x 43350] fp32 matrix
// (the 1080p pyramid, 5.66 GB) depending on the ORDER in which 128-byte lines are written?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o wbench wbench.cu ; run on the GPU box.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ void st16(float* p, float v) {
    float4 x = make_float4(v, v, v, v);
    asm volatile("st.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(x.x), "f"(x.y), "f"(x.z), "f"(x.w) : "memory");
}

// pattern 0: linear.  Each CTA owns a contiguous span, warps write 512 B per instruction.
// pattern 1: tile order as in corr_build (linear mode): CTA = column tile of 256 floats (1 KB), loop over
//            128-row blocks; warp (q, sub) writes 32 rows x 128 B per "box", 4 boxes per tile.
// pattern 2: same tiles, but a warp writes each row's whole 512 B (its half of the 1 KB) back to back:
//            box = 8 rows x 512 B.
// pattern 3: same tiles, a warp writes 4 rows x 1 KB per step (both halves) -- whole 1 KB pages at once.
// pattern 4: like 1 but the CTA advances over COLUMN tiles for a fixed 128-row block (A-stationary order).
__global__ void __launch_bounds__(256) wkernel(float* out, long long rows, long long cols, int pattern,
                                              int ntiles, int mblks) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = warp & 3, sub = warp >> 2;
    const long long total_tiles = (long long)ntiles * mblks;
    if (pattern == 0) {
        const long long n16 = rows * cols / 4;
        const long long per = (n16 + gridDim.x - 1) / gridDim.x;
        const long long b0 = per * blockIdx.x, b1 = min(n16, b0 + per);
        for (long long i = b0 + threadIdx.x; i < b1; i += blockDim.x) st16(out + i * 4, 1.f);
        return;
    }
    for (long long t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        int nt, mb;
        if (pattern == 4) { mb = (int)(t / ntiles); nt = (int)(t % ntiles); mb = (mb * 1 + 0); }
        else { nt = (int)(t % ntiles); mb = (int)(t / ntiles); }
        // for patterns 1-3 emulate "CTA owns column tile nt, sweeps mb": t enumerates mb-major with nt fastest,
        // and CTAs (148) take consecutive nt -> all CTAs work on the same mb at the same time.
        const long long row0 = (long long)mb * 128 + q * 32;
        const long long col0 = (long long)nt * 256;
        if (pattern == 1 || pattern == 4) {
            for (int j = 0; j < 4; ++j)
                for (int it = 0; it < 8; ++it) {
                    const int r = it * 4 + (lane >> 3), c = (lane & 7) * 4;
                    const long long rr = row0 + r, cc = col0 + sub * 128 + j * 32 + c;
                    if (rr < rows && cc < cols) st16(out + rr * cols + cc, 1.f);
                }
        } else if (pattern == 2) {
            for (int it = 0; it < 32; ++it) {   // one row (512 B = 32 lanes x 16 B) per instruction
                const long long rr = row0 + it, cc = col0 + sub * 128 + lane * 4;
                if (rr < rows && cc < cols) st16(out + rr * cols + cc, 1.f);
            }
        } else if (pattern == 3) {
            // warp w handles rows row0.. but both halves: 16 rows each for sub 0/1, 1 KB per row = 2 instr
            for (int it = 0; it < 16; ++it) {
                const long long rr = row0 + sub * 16 + it;
                for (int hlf = 0; hlf < 2; ++hlf) {
                    const long long cc = col0 + hlf * 128 + lane * 4;
                    if (rr < rows && cc < cols) st16(out + rr * cols + cc, 1.f);
                }
            }
        }
    }
}

int main() {
    const long long rows = 32640, cols = 43360;   // cols padded to a multiple of 4 (43350 -> 43360)
    float* out;
    cudaMalloc(&out, rows * cols * 4);
    const int ntiles = (int)((cols + 255) / 256), mblks = (int)(rows / 128);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int pattern = 0; pattern <= 4; ++pattern) {
        for (int grid : {148, 296, 592}) {
            for (int rep = 0; rep < 2; ++rep) wkernel<<<grid, 256>>>(out, rows, cols, pattern, ntiles, mblks);
            cudaEventRecord(e0);
            for (int rep = 0; rep < 5; ++rep) wkernel<<<grid, 256>>>(out, rows, cols, pattern, ntiles, mblks);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
            printf("pattern %d grid %3d: %.3f ms  %.0f GB/s  (%s)\n", pattern, grid, ms, rows * cols * 4 / ms / 1e6,
                   cudaGetErrorString(cudaGetLastError()));
        }
    }
    return 0;
}
