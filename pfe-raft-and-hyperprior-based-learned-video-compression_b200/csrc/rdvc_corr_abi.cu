// rdvc_corr_abi.cu -- the C ABI of librdvc_corr.so (see include/rdvc_corr.h).
// Argument checking, pyramid layout arithmetic, TMA descriptor encoding and the
// kernel launches.  No torch, no C++ types across the boundary, no CPU fallback.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <vector>

#include "../../include/rdvc_corr.h"
#include "corr_build_sm100.cuh"
#ifdef RDVC_EXPERIMENTS
#include "corr_build2_sm100.cuh"
#endif
#include "corr_conv1x1_sm100.cuh"
#include "corr_encoder_tail_sm100.cuh"
#include "corr_lookup.cuh"
#include "corr_pack.cuh"
#include "entropy_coder.h"
#include "mcn_conv_sm100.cuh"
#include "mcn_convx_sm100.cuh"
#include "motion_warp.cuh"
#include "preprocess.cuh"

#ifndef RDVC_PAIR_DEFAULT
#define RDVC_PAIR_DEFAULT 0   // the CTA-pair build kernel is opt-in (option key 12 = 2) until it wins
#endif
// RDVC_EXPERIMENTS (off in the product build, `python -m ..._build --experiments` turns it on in a SEPARATE
// library, lib/librdvc_corr_exp.so): compiles in the knobs that skip work for timing experiments (level store
// mask, no-load / no-store lookups, L2 store policies) and the build variants that lost their measurements
// (fused pooling epilogue, CTA-pair kernel).  The product library accepts none of them.
#ifdef RDVC_EXPERIMENTS
#define RDVC_HAS_EXPERIMENTS 1
#else
#define RDVC_HAS_EXPERIMENTS 0
#endif
#ifndef RDVC_SRC_HASH
#define RDVC_SRC_HASH "unknown"
#endif
#define RDVC_STR2(x) #x
#define RDVC_STR(x) RDVC_STR2(x)

namespace {

thread_local char g_err[512] = "";
thread_local unsigned long long g_launches = 0;
thread_local cudaEvent_t g_prof_start = nullptr, g_prof_stop = nullptr;
std::atomic<int> g_opt_lookup{0};
std::atomic<int> g_opt_tile{0};
std::atomic<int> g_opt_msplit{0};
std::atomic<int> g_opt_store_mask{15};
std::atomic<int> g_opt_mode{0};
std::atomic<int> g_opt_tma_out{1};
std::atomic<int> g_opt_policy{0};
std::atomic<int> g_opt_twl{0};
std::atomic<int> g_opt_thl{0};
std::atomic<int> g_opt_epi_warps{0};
std::atomic<int> g_opt_pair{0};
std::atomic<int> g_opt_mcn_prefetch{0};
std::atomic<int> g_dbg_c1_flags{0};
std::atomic<unsigned long long*> g_dbg_timeline{nullptr};   // RDVC_EXPERIMENTS: device buffer for conv1x1 time stamps
std::atomic<int> g_opt_mcn_kernel{0};     // 0 = auto, 1 = three boxes per tile, 2 = one box per tile (x halo)   // measured: no effect at 1-4 tiles ahead, slower beyond

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int cuda_fail(cudaError_t e, const char* what) {
    snprintf(g_err, sizeof(g_err), "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
    return static_cast<int>(e);
}

size_t elem_size(int dt) {
    switch (dt) {
        case RDVC_DT_F32: return 4;
        case RDVC_DT_BF16: return 2;
        case RDVC_DT_F16: return 2;
        default: return 0;
    }
}

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// log2 tile shape of RDVC_LAYOUT_TILED: 16-byte rows x 4 rows (one 64-byte DRAM atom) unless
// overridden by the experiment options (keys 7 / 8)
void tile_log2(int vol_dtype, int* twl, int* thl) {
    const int ow = g_opt_twl.load(), oh = g_opt_thl.load();
    *twl = ow > 0 ? ow : (vol_dtype == RDVC_DT_F32 ? 2 : 3);
    *thl = oh > 0 ? oh : 2;
    if (vol_dtype != RDVC_DT_F32 && *twl < 3) *twl = 3;  // a 16-byte word must not straddle tiles
}

// elements of one level image in the given layout (TILED pads both sides to whole tiles)
size_t level_image_elems(int h, int w, int level, int vol_dtype, int layout) {
    if (vol_dtype != RDVC_DT_F32 && vol_dtype != RDVC_DT_BF16) return 0;   // size queries answer 0 for a bad enum
    if (h <= 0 || w <= 0 || level < 0 || level >= 31) return 0;
    const size_t hl = h >> level, wl = w >> level;
    if (layout != RDVC_LAYOUT_TILED) return hl * wl;
    int twl, thl;
    tile_log2(vol_dtype, &twl, &thl);
    const size_t tiles = (((hl + (1u << thl) - 1) >> thl) << thl) * (((wl + (1u << twl) - 1) >> twl) << twl);
    // whole 256-byte units per image row of the volume (a few all-zero tiles at the end): the build's
    // wide TMA boxes write 256 contiguous bytes per row visit, and a row pitch that keeps them
    // 256-byte aligned measures 6.1 instead of 5.7 TB/s (tools/micro/wbench3.cu, pitch sweep)
    const size_t per_line = 256 / elem_size(vol_dtype);
    return (tiles + per_line - 1) / per_line * per_line;
}

size_t level_bytes(int B, int h, int w, int level, int vol_dtype, int layout) {
    const size_t N = static_cast<size_t>(h) * w;
    return static_cast<size_t>(B) * N * level_image_elems(h, w, level, vol_dtype, layout) * elem_size(vol_dtype);
}

// upper bound of the K-major operand rows of fmap2 level l over every layout / tile option
size_t operand_rows_max(int h, int w, int level) {
    const size_t hl = h >> level, wl = w >> level;
    return ((hl + 15) / 16 * 16) * ((wl + 15) / 16 * 16);
}

int check_geometry(int B, int h, int w, int num_levels) {
    if (B <= 0 || h <= 0 || w <= 0) return fail(RDVC_E_SHAPE, "non-positive dimension B=%d h=%d w=%d", B, h, w);
    if (num_levels < 1 || num_levels > rdvc::BLD_MAX_LEVELS)
        return fail(RDVC_E_UNSUPPORTED, "num_levels=%d not in [1, %d]", num_levels, rdvc::BLD_MAX_LEVELS);
    const int min_size = 2 * (1 << (num_levels - 1));  // TV:raft.py:376
    if (h < min_size || w < min_size)
        return fail(RDVC_E_TOO_SMALL,
                    "Feature maps are too small to be down-sampled by the correlation pyramid. "
                    "H and W of feature maps should be at least %d; got: (%d, %d)", min_size, h, w);
    return RDVC_OK;
}

// ---- driver entry point for TMA descriptors (no -lcuda link dependency) ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
    static std::once_flag once;
    static EncodeTiledFn fn = nullptr;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) ==
                cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

// Tiled TMA descriptor, 128-byte swizzle, zero OOB fill (loads) / clipped (stores).
int make_tmap(CUtensorMap* m, CUtensorMapDataType dt, void* base, int rank, const cuuint64_t* dims,
              const cuuint64_t* strides_bytes, const cuuint32_t* box,
              CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return fail(RDVC_E_DRIVER, "cuTensorMapEncodeTiled not available from the driver");
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = fn(m, dt, rank, base, dims, strides_bytes, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(RDVC_E_DRIVER, "cuTensorMapEncodeTiled failed: CUresult %d", (int)r);
    return RDVC_OK;
}

// per-device facts, cached per device id (a process may drive several GPUs)
constexpr int kMaxDevices = 64;

int current_device() {
    int dev = 0;
    cudaGetDevice(&dev);
    return dev;
}

int sm_count() {
    static std::atomic<int> cache[kMaxDevices];
    const int dev = current_device();
    int n = (dev >= 0 && dev < kMaxDevices) ? cache[dev].load() : 0;
    if (n == 0) {
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
        if (dev >= 0 && dev < kMaxDevices) cache[dev].store(n);
    }
    return n;
}

// cudaFuncSetAttribute(max dynamic smem) is per device: do it once per (kernel instantiation, device)
template <typename K>
int ensure_dynamic_smem(K kern, int bytes, std::atomic<unsigned long long>& done_mask, const char* what) {
    const int dev = current_device();
    const unsigned long long bit = (dev >= 0 && dev < kMaxDevices) ? (1ull << dev) : 0ull;
    if (bit && (done_mask.load() & bit)) return RDVC_OK;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) return cuda_fail(e, what);
    if (bit) done_mask.fetch_or(bit);
    return RDVC_OK;
}

template <int MODE, int TY, int TX, typename OutT, int EW, int CL = 1>
int launch_build(const CUtensorMap& ta, const CUtensorMap* tb, const CUtensorMap* to, const rdvc::BuildParams& p,
                 cudaStream_t st) {
    auto kern = rdvc::corr_build_kernel<MODE, TY, TX, OutT, EW, CL>;
    using Cfg = rdvc::BuildCfg<EW>;
    static std::atomic<unsigned long long> attr_done{0};  // per instantiation, one bit per device
    if (int rc = ensure_dynamic_smem(kern, Cfg::SMEM_LAUNCH, attr_done, "cudaFuncSetAttribute(build, max dynamic smem)"))
        return rc;
    cudaError_t e;
    if constexpr (CL == 2) {
        // clusters of two CTAs that share the fmap1 stream (TMA multicast): items are PAIRS of fmap2 tiles
        long long grid = sm_count() & ~1;
        const long long n_items = static_cast<long long>(p.B) * ((p.ntiles + 1) / 2) * p.msplit;
        if (grid > 2 * n_items) grid = 2 * n_items;
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(static_cast<unsigned>(grid));
        cfg.blockDim = dim3(Cfg::THREADS);
        cfg.dynamicSmemBytes = Cfg::SMEM_LAUNCH;
        cfg.stream = st;
        cudaLaunchAttribute attr;
        attr.id = cudaLaunchAttributeClusterDimension;
        attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
        cfg.attrs = &attr;
        cfg.numAttrs = 1;
        if (g_prof_start) cudaEventRecord(g_prof_start, st);
        e = cudaLaunchKernelEx(&cfg, kern, ta, tb[0], tb[1], tb[2], tb[3], to[0], to[1], to[2], to[3], p);
        if (g_prof_stop) cudaEventRecord(g_prof_stop, st);
    } else {
        long long grid = sm_count();
        const long long n_items = static_cast<long long>(p.B) * p.ntiles * p.msplit;
        if (grid > n_items) grid = n_items;
        if (g_prof_start) cudaEventRecord(g_prof_start, st);
        kern<<<static_cast<unsigned>(grid), Cfg::THREADS, Cfg::SMEM_LAUNCH, st>>>(
            ta, tb[0], tb[1], tb[2], tb[3], to[0], to[1], to[2], to[3], p);
        if (g_prof_stop) cudaEventRecord(g_prof_stop, st);
        e = cudaGetLastError();
    }
    ++g_launches;
    if (e != cudaSuccess) return cuda_fail(e, "corr_build_kernel launch");
    return RDVC_OK;
}

#ifdef RDVC_EXPERIMENTS
template <typename OutT>
int launch_build_pair(const CUtensorMap& ta, const CUtensorMap* tb, const CUtensorMap* to, const rdvc::BuildParams& p,
                      cudaStream_t st) {
    auto kern = rdvc::corr_build2_kernel<OutT>;
    static std::atomic<unsigned long long> attr_done{0};
    if (int rc = ensure_dynamic_smem(kern, rdvc::B2_SMEM_LAUNCH, attr_done, "cudaFuncSetAttribute(build pair, max dynamic smem)"))
        return rc;
    long long grid = sm_count() & ~1;     // whole CTA pairs
    const long long n_items = static_cast<long long>(p.B) * p.ntiles * p.msplit;
    if (grid > 2 * n_items) grid = 2 * n_items;
    if (g_prof_start) cudaEventRecord(g_prof_start, st);
    kern<<<static_cast<unsigned>(grid), rdvc::B2_THREADS, rdvc::B2_SMEM_LAUNCH, st>>>(
        ta, tb[0], tb[1], tb[2], tb[3], to[0], to[1], to[2], to[3], p);
    if (g_prof_stop) cudaEventRecord(g_prof_stop, st);
    ++g_launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "corr_build2_kernel launch");
    return RDVC_OK;
}

#endif

// both feature maps -> K-major bf16 rows (fmap2 at `levels2` pyramid levels), one launch
template <typename T, bool QUICK, bool VEC>
int launch_pack_impl(const void* f1, const void* f2, void* a_km, void* const* b_km, int B, int D, int h, int w,
                int levels2, int layout, int twl, int thl, const size_t* nl_of, int f16, cudaStream_t st) {
    auto kern = rdvc::corr_pack_kernel<T, QUICK, VEC>;
    static std::atomic<unsigned long long> attr_done{0};
    if (int rc = ensure_dynamic_smem(kern, rdvc::PACK_SMEM_BYTES, attr_done, "cudaFuncSetAttribute(pack, max dynamic smem)"))
        return rc;
    rdvc::PackParams pp;
    memset(&pp, 0, sizeof(pp));
    pp.src[0] = f1; pp.src[1] = f2;
    pp.dst[0][0] = static_cast<__nv_bfloat16*>(a_km);
    for (int l = 0; l < 4; ++l) pp.dst[1][l] = static_cast<__nv_bfloat16*>(b_km[l]);
    pp.levels[0] = 1; pp.levels[1] = levels2;
    pp.B = B; pp.D = D; pp.h = h; pp.w = w;
    pp.tiled[0] = 0; pp.tiled[1] = (layout == RDVC_LAYOUT_TILED);
    pp.twl = twl; pp.thl = thl;
    pp.f16 = f16;
    pp.img[0][0] = h * w;
    for (int l = 0; l < 4; ++l) pp.img[1][l] = static_cast<int>(nl_of[l]);
    dim3 grid((D / rdvc::PACK_CG) * ((w + rdvc::PACK_TX - 1) / rdvc::PACK_TX), (h + rdvc::PACK_TY - 1) / rdvc::PACK_TY,
              2 * B);
    kern<<<grid, rdvc::PACK_THREADS, rdvc::PACK_SMEM_BYTES, st>>>(pp);
    ++g_launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "corr_pack_kernel launch");
    return RDVC_OK;
}

// KIND: 0 / 1 = row-major layout, scalar / 128-bit loads; 2.. = tiled layout:
//       2 + DBG (DBG in 0..2)
template <typename T>
int launch_pack(const void* f1, const void* f2, void* a_km, void* const* b_km, int B, int D, int h, int w,
                int levels2, int layout, int twl, int thl, const size_t* nl_of, int f16, cudaStream_t st) {
    const bool quick = (layout != RDVC_LAYOUT_TILED) || (thl == 2 && (twl == 2 || twl == 3));
    // 4-element vector loads need w % 4 == 0 AND 4-element-aligned bases (a contiguous torch view can start at
    // an odd storage offset); anything else takes the scalar instantiation
    const uintptr_t amask = 4 * sizeof(T) - 1;
    const bool vec = (w % 4) == 0 && !((reinterpret_cast<uintptr_t>(f1) | reinterpret_cast<uintptr_t>(f2)) & amask);
    if (quick && vec) return launch_pack_impl<T, true, true>(f1, f2, a_km, b_km, B, D, h, w, levels2, layout, twl, thl, nl_of, f16, st);
    if (quick) return launch_pack_impl<T, true, false>(f1, f2, a_km, b_km, B, D, h, w, levels2, layout, twl, thl, nl_of, f16, st);
    if (vec) return launch_pack_impl<T, false, true>(f1, f2, a_km, b_km, B, D, h, w, levels2, layout, twl, thl, nl_of, f16, st);
    return launch_pack_impl<T, false, false>(f1, f2, a_km, b_km, B, D, h, w, levels2, layout, twl, thl, nl_of, f16, st);
}

// row-major kernels: KIND 0 / 1 = scalar / 128-bit loads
template <int R, typename VolT, int KIND>
int launch_lookup(const rdvc::LookupParams& p, cudaStream_t st) {
    const unsigned grid = static_cast<unsigned>((p.total + 31) / 32);
    rdvc::corr_lookup_kernel<R, VolT, KIND == 1><<<grid, 32 * p.num_levels, 0, st>>>(p);
    ++g_launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "corr_lookup_kernel launch");
    return RDVC_OK;
}

template <typename VolT, int KIND>
int dispatch_lookup_radius(int radius, const rdvc::LookupParams& p, cudaStream_t st) {
    switch (radius) {
        case 1: return launch_lookup<1, VolT, KIND>(p, st);
        case 2: return launch_lookup<2, VolT, KIND>(p, st);
        case 3: return launch_lookup<3, VolT, KIND>(p, st);
        case 4: return launch_lookup<4, VolT, KIND>(p, st);
        default: return fail(RDVC_E_UNSUPPORTED, "radius=%d not in [1, 4]", radius);
    }
}

// tiled kernel: OUT = output form (rdvc::LKP_OUT_*), DBG = timing experiments (RDVC_EXPERIMENTS builds only)
template <int R, typename VolT, int OUT, int DBG>
int launch_lookup_tiled(const rdvc::LookupParams& p, cudaStream_t st) {
    const unsigned grid = static_cast<unsigned>((p.total + 31) / 32);
    rdvc::corr_lookup_tiled_kernel<R, VolT, OUT, DBG><<<grid, 32 * p.num_levels, 0, st>>>(p);
    ++g_launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "corr_lookup_tiled_kernel launch");
    return RDVC_OK;
}

template <typename VolT, int OUT, int DBG>
int dispatch_lookup_tiled_radius(int radius, const rdvc::LookupParams& p, cudaStream_t st) {
    switch (radius) {
        case 1: return launch_lookup_tiled<1, VolT, OUT, DBG>(p, st);
        case 2: return launch_lookup_tiled<2, VolT, OUT, DBG>(p, st);
        case 3: return launch_lookup_tiled<3, VolT, OUT, DBG>(p, st);
        case 4: return launch_lookup_tiled<4, VolT, OUT, DBG>(p, st);
        default: return fail(RDVC_E_UNSUPPORTED, "radius=%d not in [1, 4]", radius);
    }
}

template <typename VolT>
int dispatch_lookup_tiled(int out_mode, int dbg, int radius, const rdvc::LookupParams& p, cudaStream_t st) {
#ifdef RDVC_EXPERIMENTS
    if (dbg == 1 && out_mode == rdvc::LKP_OUT_NCHW_F32) return dispatch_lookup_tiled_radius<VolT, rdvc::LKP_OUT_NCHW_F32, 1>(radius, p, st);
    if (dbg == 2 && out_mode == rdvc::LKP_OUT_NCHW_F32) return dispatch_lookup_tiled_radius<VolT, rdvc::LKP_OUT_NCHW_F32, 2>(radius, p, st);
    if (dbg == 1 && out_mode == rdvc::LKP_OUT_KM_BF16) return dispatch_lookup_tiled_radius<VolT, rdvc::LKP_OUT_KM_BF16, 1>(radius, p, st);
    if (dbg == 2 && out_mode == rdvc::LKP_OUT_KM_BF16) return dispatch_lookup_tiled_radius<VolT, rdvc::LKP_OUT_KM_BF16, 2>(radius, p, st);
#else
    (void)dbg;
#endif
    switch (out_mode) {
        case rdvc::LKP_OUT_NCHW_F16: return dispatch_lookup_tiled_radius<VolT, rdvc::LKP_OUT_NCHW_F16, 0>(radius, p, st);
        case rdvc::LKP_OUT_KM_BF16: return dispatch_lookup_tiled_radius<VolT, rdvc::LKP_OUT_KM_BF16, 0>(radius, p, st);
        case rdvc::LKP_OUT_KM_F16: return dispatch_lookup_tiled_radius<VolT, rdvc::LKP_OUT_KM_F16, 0>(radius, p, st);
        default: return dispatch_lookup_tiled_radius<VolT, rdvc::LKP_OUT_NCHW_F32, 0>(radius, p, st);
    }
}

template <typename OutT>
cudaError_t launch_conv1x1(const CUtensorMap& tm_a, const CUtensorMap& tm_w, const rdvc::Conv1x1Params& p, unsigned grid,
                           cudaStream_t st) {
    auto kern = rdvc::corr_conv1x1_kernel<OutT>;
    static std::atomic<unsigned long long> attr_done{0};
    if (ensure_dynamic_smem(kern, rdvc::C1_SMEM_LAUNCH, attr_done, "cudaFuncSetAttribute(conv1x1, max dynamic smem)"))
        return cudaErrorInvalidValue;
    kern<<<grid, rdvc::C1_THREADS, rdvc::C1_SMEM_LAUNCH, st>>>(tm_a, tm_w, p);
    ++g_launches;
    return cudaGetLastError();
}

// ---- build plan: everything rdvc_corr_build derives from its arguments and the options, cached ----
struct BuildKey {            // compared with memcmp: no padding (2 pointers + 20 ints), zeroed before it is filled
    void* pyramid;
    void* workspace;
    int device, B, D, h, w, f16_ops, vol_dtype, layout, num_levels;
    int opt_mode, opt_tile, opt_msplit, opt_tma_out, opt_epi, opt_pair, opt_twl, opt_thl, opt_store_mask, opt_policy;
    int reserved;
};
static_assert(sizeof(BuildKey) == 2 * sizeof(void*) + 20 * sizeof(int), "BuildKey must have no padding");

struct BuildPlan {
    BuildKey key;
    void* a_km;                              // fmap1 as (B, N, D) 16-bit rows
    void* b_km[rdvc::BLD_MAX_LEVELS];        // fmap2 level l as (B, n_l, D)
    size_t nl_of[rdvc::BLD_MAX_LEVELS];      // pixels (operand rows / output columns) of level l in this layout
    int twl, thl;
    void* memset_ptr;
    size_t memset_bytes;
    bool linear, pair;
    bool cluster;                            // clusters of two CTAs sharing the fmap1 stream (TMA multicast)
    int tile, ew;
    CUtensorMap ta, tb[rdvc::BLD_MAX_LEVELS], to[rdvc::BLD_MAX_LEVELS], tb2[rdvc::BLD_MAX_LEVELS];
    rdvc::BuildParams p, p2;
};

constexpr size_t kPlanCacheSize = 8;
std::mutex g_plan_mutex;
std::vector<BuildPlan> g_plans;              // most recently used first
std::atomic<unsigned long long> g_plan_hits{0}, g_plan_misses{0};

bool plan_cache_get(const BuildKey& key, BuildPlan* out) {
    std::lock_guard<std::mutex> lock(g_plan_mutex);
    for (size_t i = 0; i < g_plans.size(); ++i)
        if (memcmp(&g_plans[i].key, &key, sizeof(key)) == 0) {
            if (i) std::swap(g_plans[0], g_plans[i]);
            *out = g_plans[0];
            ++g_plan_hits;
            return true;
        }
    ++g_plan_misses;
    return false;
}

void plan_cache_put(const BuildPlan& plan) {
    std::lock_guard<std::mutex> lock(g_plan_mutex);
    if (g_plans.size() >= kPlanCacheSize) g_plans.pop_back();
    g_plans.insert(g_plans.begin(), plan);
}

// m-range slices per fmap2 tile: enough (tile, slice) items to fill `G` CTAs (or pairs) in whole waves.
// cost = waves x (m-blocks per slice + ~0.5 for the tile reload); a mild bias toward few slices keeps
// concurrent CTAs on the same query rows (see the kernel).
int choose_msplit(long long units, int m_blks, int G, int forced) {
    int best = 1;
    double best_cost = 1e300;
    const int s_max = m_blks < 64 ? m_blks : 64;
    for (int S = 1; S <= s_max; ++S) {
        const long long waves = (units * S + G - 1) / G;
        const double cost = waves * ((m_blks + S - 1) / S + 0.5) * (1.0 + 0.005 * S);
        if (cost < best_cost) { best_cost = cost; best = S; }
    }
    return (forced > 0) ? (forced < m_blks ? forced : m_blks) : best;
}

// Where the K-major operand rows live inside a workspace of rdvc_corr_workspace_bytes(B, D, h, w): fmap1 rows first,
// then fmap2's rows per level (each block 256-byte aligned and sized for the largest layout), their row counts in this
// layout, and the byte range that must be zeroed before a pack (levels with layout padding).
struct WorkspaceLayout {
    void* a_km;
    void* b_km[rdvc::BLD_MAX_LEVELS];
    size_t nl_of[rdvc::BLD_MAX_LEVELS];
    int twl, thl;
    void* memset_ptr;
    size_t memset_bytes;
};

void workspace_layout(void* workspace, int B, int D, int h, int w, int vol_dtype, int layout, int num_levels,
                      WorkspaceLayout* out) {
    WorkspaceLayout& wl = *out;
    memset(&wl, 0, sizeof(wl));
    const int N = h * w;
    uint8_t* ws = static_cast<uint8_t*>(workspace);
    wl.a_km = ws;
    size_t off = align_up(static_cast<size_t>(B) * N * D * 2, 256);
    for (int l = 0; l < rdvc::BLD_MAX_LEVELS; ++l) {
        wl.b_km[l] = ws + off;
        off += align_up(static_cast<size_t>(B) * operand_rows_max(h, w, l) * D * 2, 256);
    }
    for (int l = 0; l < rdvc::BLD_MAX_LEVELS; ++l) wl.nl_of[l] = level_image_elems(h, w, l, vol_dtype, layout);
    if (layout == RDVC_LAYOUT_TILED) {
        tile_log2(vol_dtype, &wl.twl, &wl.thl);
        int first = -1;
        for (int l = 0; l < num_levels && first < 0; ++l)
            if (wl.nl_of[l] != static_cast<size_t>(h >> l) * (w >> l)) first = l;
        if (first >= 0) {
            uint8_t* lo = static_cast<uint8_t*>(wl.b_km[first]);
            uint8_t* hi = static_cast<uint8_t*>(wl.b_km[num_levels - 1]) + static_cast<size_t>(B) * wl.nl_of[num_levels - 1] * D * 2;
            wl.memset_ptr = lo;
            wl.memset_bytes = static_cast<size_t>(hi - lo);
        }
    }
}

int make_build_plan(const BuildKey& k, BuildPlan* plan) {
    BuildPlan& pl = *plan;
    memset(&pl, 0, sizeof(pl));
    pl.key = k;
    const int B = k.B, D = k.D, h = k.h, w = k.w, num_levels = k.num_levels, vol_dtype = k.vol_dtype, layout = k.layout;
    const int N = h * w;
    int mode = k.opt_mode;
    if (mode == 0) mode = 2;  // default: linear (pooled fmap2 rows), see corr_build_sm100.cuh
    pl.linear = (mode == 2);
    if (!pl.linear && !RDVC_HAS_EXPERIMENTS)
        return fail(RDVC_E_UNSUPPORTED, "the fused-epilogue build mode exists in RDVC_EXPERIMENTS builds only");
    if (!pl.linear && layout != RDVC_LAYOUT_ROWMAJOR)
        return fail(RDVC_E_UNSUPPORTED, "the fused-epilogue build mode writes RDVC_LAYOUT_ROWMAJOR only");
    {
        WorkspaceLayout wl;
        workspace_layout(k.workspace, B, D, h, w, vol_dtype, layout, num_levels, &wl);
        pl.a_km = wl.a_km;
        for (int l = 0; l < rdvc::BLD_MAX_LEVELS; ++l) { pl.b_km[l] = wl.b_km[l]; pl.nl_of[l] = wl.nl_of[l]; }
        pl.twl = wl.twl; pl.thl = wl.thl;
        pl.memset_ptr = wl.memset_ptr; pl.memset_bytes = wl.memset_bytes;
    }
    // operand format: fp16 inputs (the reference's default autocast, R:codec_processing.py:1436) stay fp16
    // -- same tensor-core rate, 11 instead of 8 mantissa bits; everything else is multiplied as bf16
    const CUtensorMapDataType op_dt = k.f16_ops ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;

    // fused mode tile shape: 16x16 fmap2 pixels unless 8x32 wastes less padding
    int tile = k.opt_tile;
    auto padded = [&](int ty, int tx) {
        return static_cast<long long>((h + ty - 1) / ty) * ty * ((w + tx - 1) / tx) * tx;
    };
    if (tile == 0) tile = (padded(8, 32) < padded(16, 16)) ? 2 : 1;
    pl.tile = tile;
    const int TY = (tile == 1) ? 16 : 8, TX = (tile == 1) ? 16 : 32;

    // fmap1 multicast across CTA pairs (experiments library, option key 12 = 3): bit-identical, measured no faster
    // (1080p bf16 0.654 -> 0.646 ms, fp32 unchanged, 1440p 9 % slower: DESIGN.md 3.2) -- the L2 -> SM operand feed is
    // not what bounds the build
    pl.cluster = RDVC_HAS_EXPERIMENTS && pl.linear && sm_count() >= 2 && k.opt_pair == 3;
    // TMA descriptors over the repacked maps
    int rc;
    {
        cuuint64_t dims[3] = {(cuuint64_t)D, (cuuint64_t)N, (cuuint64_t)B};
        cuuint64_t str[2] = {(cuuint64_t)D * 2, (cuuint64_t)N * D * 2};
        cuuint32_t box[3] = {rdvc::BLD_BLOCK_K, static_cast<cuuint32_t>(pl.cluster ? rdvc::BLD_BLOCK_M / 2 : rdvc::BLD_BLOCK_M), 1};
        rc = make_tmap(&pl.ta, op_dt, pl.a_km, 3, dims, str, box);
        if (rc) return rc;
    }
    if (pl.linear) {
        for (int l = 0; l < rdvc::BLD_MAX_LEVELS; ++l) {
            const int ll = l < num_levels ? l : 0;  // unused slots alias level 0
            const cuuint64_t nl = (cuuint64_t)pl.nl_of[ll];
            cuuint64_t dims[3] = {(cuuint64_t)D, nl, (cuuint64_t)B};
            cuuint64_t str[2] = {(cuuint64_t)D * 2, nl * D * 2};
            cuuint32_t box[3] = {rdvc::BLD_BLOCK_K, rdvc::BLD_BLOCK_N, 1};
            rc = make_tmap(&pl.tb[l], op_dt, pl.b_km[ll], 3, dims, str, box);
            if (rc) return rc;
        }
    } else {
        cuuint64_t dims[4] = {(cuuint64_t)D, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)B};
        cuuint64_t str[3] = {(cuuint64_t)D * 2, (cuuint64_t)w * D * 2, (cuuint64_t)N * D * 2};
        cuuint32_t box[4] = {rdvc::BLD_BLOCK_K, (cuuint32_t)TX, (cuuint32_t)TY, 1};
        rc = make_tmap(&pl.tb[0], op_dt, pl.b_km[0], 4, dims, str, box);
        if (rc) return rc;
        pl.tb[1] = pl.tb[2] = pl.tb[3] = pl.tb[0];
    }

    rdvc::BuildParams& p = pl.p;
    for (int l = 0; l < num_levels; ++l) {
        p.lvl[l] = static_cast<uint8_t*>(k.pyramid) + rdvc_corr_level_offset_bytes(B, h, w, l, vol_dtype, layout);
        p.hl[l] = h >> l;
        p.wl[l] = w >> l;
        p.nl[l] = static_cast<int>(pl.nl_of[l]);
    }
    p.B = B; p.h = h; p.w = w; p.N = N;
    p.num_levels = num_levels;
    p.kc = D / 64;
    p.m_blks = (N + rdvc::BLD_BLOCK_M - 1) / rdvc::BLD_BLOCK_M;
    p.nty = (h + TY - 1) / TY;
    p.ntx = (w + TX - 1) / TX;
    if (pl.linear) {
        int t = 0;
        for (int l = 0; l < rdvc::BLD_MAX_LEVELS; ++l) {
            p.tile_start[l] = t;
            if (l < num_levels) t += (p.nl[l] + rdvc::BLD_BLOCK_N - 1) / rdvc::BLD_BLOCK_N;
        }
        p.ntiles = t;
    } else {
        p.ntiles = p.nty * p.ntx;
    }
    p.scale = static_cast<float>(1.0 / std::sqrt(static_cast<double>(D)));
    p.ab_format = k.f16_ops ? 0 : 1;
    p.dbg_store_mask = RDVC_HAS_EXPERIMENTS ? k.opt_store_mask : 15;
    p.dbg_policy = RDVC_HAS_EXPERIMENTS ? k.opt_policy : 0;
    p.msplit = pl.cluster ? choose_msplit(static_cast<long long>(B) * ((p.ntiles + 1) / 2), p.m_blks, sm_count() / 2, k.opt_msplit)
                          : choose_msplit(static_cast<long long>(B) * p.ntiles, p.m_blks, sm_count(), k.opt_msplit);

    // output descriptors (linear mode).  Row pitch a multiple of 128 bytes (always, in the tiled
    // layout): level l as a {128 B, pitch/128, N, B} tensor written in boxes of 16 query rows x 256
    // contiguous bytes.  Row pitch only 16-byte aligned: a {n_l, N, B} tensor, boxes of 32 rows x 128
    // bytes.  Anything else takes the staged-store path.  (option key 5: 0 = staged only, 1 = auto,
    // 2 = never the wide boxes)
    for (int l = 0; l < rdvc::BLD_MAX_LEVELS; ++l) pl.to[l] = pl.tb[0];     // placeholders for levels without a store map
    const int tma_opt = k.opt_tma_out;
    if (pl.linear && tma_opt) {
        const bool f32 = (vol_dtype == RDVC_DT_F32);
        const cuuint64_t es = f32 ? 4 : 2;
        const CUtensorMapDataType dt = f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
        for (int l = 0; l < num_levels; ++l) {
            const cuuint64_t nl = (cuuint64_t)p.nl[l];
            if ((nl * es) % 128 == 0 && tma_opt == 1) {
                cuuint64_t dims[4] = {128 / es, nl * es / 128, (cuuint64_t)N, (cuuint64_t)B};
                cuuint64_t str[3] = {128, nl * es, (cuuint64_t)N * nl * es};
                cuuint32_t box[4] = {(cuuint32_t)(128 / es), 2, 16, 1};
                rc = make_tmap(&pl.to[l], dt, p.lvl[l], 4, dims, str, box);
                if (rc) return rc;
                p.tma_out |= 2 << (2 * l);
            } else if ((nl * es) % 16 == 0) {
                cuuint64_t dims[3] = {nl, (cuuint64_t)N, (cuuint64_t)B};
                cuuint64_t str[2] = {nl * es, (cuuint64_t)N * nl * es};
                cuuint32_t box[3] = {(cuuint32_t)(128 / es), 32, 1};
                rc = make_tmap(&pl.to[l], dt, p.lvl[l], 3, dims, str, box);
                if (rc) return rc;
                p.tma_out |= 1 << (2 * l);
            }
        }
    }
    // epilogue shape per storage type: see BuildCfg (option key 9: 0 = auto, 4 / 8 = force)
    pl.ew = k.opt_epi ? k.opt_epi : ((vol_dtype == RDVC_DT_F32) ? 8 : 4);

#ifdef RDVC_EXPERIMENTS
    // CTA-pair kernel (cta_group::2): linear mode with every level on the wide-box path (always true for
    // the tiled layout).  Option key 12: 0 = auto, 1 = single-CTA kernel, 2 = pair kernel where possible.
    bool pair_ok = pl.linear && (sm_count() >= 2);
    for (int l = 0; l < num_levels; ++l) pair_ok = pair_ok && (((p.tma_out >> (2 * l)) & 3) == 2);
    if (pair_ok && (k.opt_pair == 2 || (k.opt_pair == 0 && RDVC_PAIR_DEFAULT))) {
        for (int l = 0; l < rdvc::BLD_MAX_LEVELS; ++l) {
            const int ll = l < num_levels ? l : 0;
            const cuuint64_t nl = (cuuint64_t)pl.nl_of[ll];
            cuuint64_t dims[3] = {(cuuint64_t)D, nl, (cuuint64_t)B};
            cuuint64_t str[2] = {(cuuint64_t)D * 2, nl * D * 2};
            cuuint32_t box[3] = {rdvc::BLD_BLOCK_K, rdvc::BLD_BLOCK_N / 2, 1};   // each CTA loads half the tile
            rc = make_tmap(&pl.tb2[l], op_dt, pl.b_km[ll], 3, dims, str, box);
            if (rc) return rc;
        }
        pl.p2 = p;
        pl.p2.m_blks = (N + 2 * rdvc::BLD_BLOCK_M - 1) / (2 * rdvc::BLD_BLOCK_M);
        pl.p2.msplit = choose_msplit(static_cast<long long>(B) * pl.p2.ntiles, pl.p2.m_blks, sm_count() / 2, k.opt_msplit);
        pl.pair = true;
    }
#endif
    return RDVC_OK;
}

struct HostArena {  // scratch owned by rdvc_corr_pair_host*, one per (thread, slot), bound to one device
    int device = -1;
    void* dev = nullptr;
    size_t bytes = 0;
    cudaStream_t compute = nullptr, copy = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};   // one per device-side result buffer: its D2H copy finished
    cudaEvent_t done[4] = {nullptr, nullptr, nullptr, nullptr}; // one per device-side result buffer: its lookup finished
    bool pending = false;   // a submitted pair has not been waited for yet
};
constexpr int kHostSlots = 2;
constexpr int kOutBufs = 4;     // lookups run ahead of their device->host copies by up to this many results
thread_local HostArena g_arena[kHostSlots];

void release_arena(HostArena& a) {
    if (a.compute) cudaStreamSynchronize(a.compute);
    if (a.copy) cudaStreamSynchronize(a.copy);
    if (a.dev) cudaFree(a.dev);
    if (a.compute) cudaStreamDestroy(a.compute);
    if (a.copy) cudaStreamDestroy(a.copy);
    for (auto& e : a.ev) if (e) cudaEventDestroy(e);
    for (auto& e : a.done) if (e) cudaEventDestroy(e);
    a = HostArena();
}

}  // namespace

extern "C" {

int rdvc_corr_version(void) { return RDVC_CORR_VERSION; }
const char* rdvc_corr_last_error(void) { return g_err; }
unsigned long long rdvc_corr_launch_count(void) { return g_launches; }

void rdvc_corr_set_profile_events(void* start, void* stop) {
    g_prof_start = static_cast<cudaEvent_t>(start);
    g_prof_stop = static_cast<cudaEvent_t>(stop);
}

int rdvc_corr_set_option(int key, int value) {
    const bool exp = RDVC_HAS_EXPERIMENTS;
    if (key == 0 && value >= 0 && value <= (exp ? 4 : 2)) { g_opt_lookup = value; return RDVC_OK; }
    if (key == 1 && value >= 0 && value <= 2) { g_opt_tile = value; return RDVC_OK; }
    if (key == 2 && value >= 0) { g_opt_msplit = value; return RDVC_OK; }
    if (key == 3 && value >= 0 && value <= 63 && (exp || value == 15)) { g_opt_store_mask = value; return RDVC_OK; }
    if (key == 4 && value >= 0 && value <= 2 && (exp || value != 1)) { g_opt_mode = value; return RDVC_OK; }
    if (key == 5 && value >= 0 && value <= 2) { g_opt_tma_out = value; return RDVC_OK; }
    if (key == 6 && value >= 0 && value <= 2 && (exp || value == 0)) { g_opt_policy = value; return RDVC_OK; }
    if (key == 7 && value >= 0 && value <= 4) { g_opt_twl = value; return RDVC_OK; }
    if (key == 8 && value >= 0 && value <= 4) { g_opt_thl = value; return RDVC_OK; }
    if (key == 9 && (value == 0 || value == 4 || value == 8)) { g_opt_epi_warps = value; return RDVC_OK; }
    if (key == 12 && value >= 0 && value <= 3 && (exp || value <= 1)) { g_opt_pair = value; return RDVC_OK; }
    if (key == 13 && value >= 0 && value <= 16) { g_opt_mcn_prefetch = value; return RDVC_OK; }
    if (key == 14 && value >= 0 && value <= 2) { g_opt_mcn_kernel = value; return RDVC_OK; }
    return fail(RDVC_E_UNSUPPORTED, "unknown option key=%d value=%d%s", key, value,
                exp ? "" : " (work-skipping knobs and the fused / CTA-pair build variants exist in RDVC_EXPERIMENTS builds only)");
}

const char* rdvc_corr_build_info(void) {
    // the marker lets _build.py read the hash out of the file without loading it
    return "RDVC_SRC_HASH=" RDVC_SRC_HASH " version=" RDVC_STR(RDVC_CORR_VERSION) " experiments=" RDVC_STR(RDVC_HAS_EXPERIMENTS);
}

unsigned long long rdvc_corr_plan_cache_hits(void) { return g_plan_hits.load(); }

#ifdef RDVC_EXPERIMENTS
// experiments library only (not in the header): 16 globaltimer stamps per CTA of the next rdvc_conv1x1 launches
void rdvc_exp_conv1x1_timeline(void* device_buffer) { g_dbg_timeline = static_cast<unsigned long long*>(device_buffer); }
void rdvc_exp_conv1x1_flags(int flags) { g_dbg_c1_flags = flags; }     // 1 = no output stores, 2 = no TMEM reads
#endif

size_t rdvc_corr_level_image_elems(int h, int w, int level, int vol_dtype, int layout) {
    if (elem_size(vol_dtype) == 0 || h <= 0 || w <= 0 || level < 0) return 0;
    return level_image_elems(h, w, level, vol_dtype, layout);
}

int rdvc_corr_tile_shape(int vol_dtype, int* tile_w, int* tile_h) {
    if (vol_dtype != RDVC_DT_F32 && vol_dtype != RDVC_DT_BF16)
        return fail(RDVC_E_DTYPE, "unsupported vol_dtype=%d", vol_dtype);
    if (!tile_w || !tile_h) return fail(RDVC_E_NULL, "null pointer argument");
    int twl, thl;
    tile_log2(vol_dtype, &twl, &thl);
    *tile_w = 1 << twl;
    *tile_h = 1 << thl;
    return RDVC_OK;
}

size_t rdvc_corr_level_offset_bytes(int B, int h, int w, int level, int vol_dtype, int layout) {
    if ((vol_dtype != RDVC_DT_F32 && vol_dtype != RDVC_DT_BF16) || B <= 0 || h <= 0 || w <= 0 || level < 0) return 0;
    size_t off = 0;
    for (int l = 0; l < level; ++l) off += align_up(level_bytes(B, h, w, l, vol_dtype, layout), 256);
    return off;
}

size_t rdvc_corr_pyramid_bytes(int B, int h, int w, int num_levels, int vol_dtype, int layout) {
    // every level padded to 256 bytes, so 16-byte gathers at a level's tail stay inside
    return rdvc_corr_level_offset_bytes(B, h, w, num_levels, vol_dtype, layout);
}

size_t rdvc_corr_workspace_bytes(int B, int D, int h, int w) {
    // fmap1 K-major + fmap2 K-major at every pyramid level (levels 1..3 are used by the
    // linear build mode only), each 256-byte aligned
    size_t total = align_up(static_cast<size_t>(B) * h * w * D * 2, 256);
    for (int l = 0; l < rdvc::BLD_MAX_LEVELS; ++l)
        total += align_up(static_cast<size_t>(B) * operand_rows_max(h, w, l) * D * 2, 256);
    return total;
}

}  // extern "C"

namespace {
// fmap1 == nullptr: the workspace already holds the K-major operand rows (rdvc_corr_build_packed)
int build_impl(const void* fmap1, const void* fmap2, int B, int D, int h, int w, int in_dtype,
               void* pyramid, int vol_dtype, int layout, int num_levels, void* workspace,
               size_t workspace_bytes, void* stream) {
    const bool packed = (fmap1 == nullptr);
    if ((!packed && !fmap2) || !pyramid || !workspace) return fail(RDVC_E_NULL, "null pointer argument");
    int rc = check_geometry(B, h, w, num_levels);
    if (rc) return rc;
    if (D <= 0) return fail(RDVC_E_SHAPE, "non-positive channel count D=%d", D);
    if (D % 64 != 0 || D > 64 * rdvc::BLD_MAX_KC)
        return fail(RDVC_E_UNSUPPORTED, "D=%d must be a multiple of 64 and <= %d", D, 64 * rdvc::BLD_MAX_KC);
    if (in_dtype != RDVC_DT_F32 && in_dtype != RDVC_DT_BF16 && in_dtype != RDVC_DT_F16)
        return fail(RDVC_E_DTYPE, "unsupported in_dtype=%d", in_dtype);
    if (vol_dtype != RDVC_DT_F32 && vol_dtype != RDVC_DT_BF16)
        return fail(RDVC_E_DTYPE, "unsupported vol_dtype=%d", vol_dtype);
    if (layout != RDVC_LAYOUT_ROWMAJOR && layout != RDVC_LAYOUT_TILED)
        return fail(RDVC_E_UNSUPPORTED, "unknown pyramid layout %d", layout);
    if (workspace_bytes < rdvc_corr_workspace_bytes(B, D, h, w))
        return fail(RDVC_E_WORKSPACE, "workspace %zu < required %zu", workspace_bytes,
                    rdvc_corr_workspace_bytes(B, D, h, w));
    if ((reinterpret_cast<uintptr_t>(pyramid) & 255) || (reinterpret_cast<uintptr_t>(workspace) & 255))
        return fail(RDVC_E_ALIGN, "pyramid and workspace must be 256-byte aligned");
    if (static_cast<long long>(h) * w > (1LL << 30) / 4)
        return fail(RDVC_E_UNSUPPORTED, "h*w too large for 32-bit pixel indices");

    cudaStream_t st = static_cast<cudaStream_t>(stream);
    BuildKey key;
    memset(&key, 0, sizeof(key));
    key.device = current_device();
    key.pyramid = pyramid; key.workspace = workspace;
    key.B = B; key.D = D; key.h = h; key.w = w; key.f16_ops = (in_dtype == RDVC_DT_F16) ? 1 : 0;
    key.vol_dtype = vol_dtype; key.layout = layout; key.num_levels = num_levels;
    key.opt_mode = g_opt_mode.load(); key.opt_tile = g_opt_tile.load(); key.opt_msplit = g_opt_msplit.load();
    key.opt_tma_out = g_opt_tma_out.load(); key.opt_epi = g_opt_epi_warps.load(); key.opt_pair = g_opt_pair.load();
    key.opt_twl = g_opt_twl.load(); key.opt_thl = g_opt_thl.load();
    key.opt_store_mask = g_opt_store_mask.load(); key.opt_policy = g_opt_policy.load();

    // The plan (pointers into the workspace, 9 TMA descriptors, work split) depends only on the key: a per-process,
    // mutex-guarded cache of the last few plans makes a repeated call (RAFT calls build once per frame pair with
    // the same buffers) skip the descriptor encoding and the m-split search.
    BuildPlan plan;
    if (!plan_cache_get(key, &plan)) {
        rc = make_build_plan(key, &plan);
        if (rc) return rc;
        plan_cache_put(plan);
    }

    if (plan.memset_bytes && !packed) {
        // padding pixels of a level must come out of the GEMM as exact zeros: clear the operand rows of
        // the levels that have any (the pack kernel writes only real pixels) -- ONE memset from the first
        // padded level to the end of the last level (the per-level buffers are contiguous)
        cudaError_t e = cudaMemsetAsync(plan.memset_ptr, 0, plan.memset_bytes, st);
        if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(padded operand rows)");
    }
    // 1. repack to K-major 16-bit rows; the linear mode also needs the pooled fmap2 levels
    if (packed && !plan.linear) return fail(RDVC_E_UNSUPPORTED, "pre-packed operands need the linear build mode");
    if (!packed) {
        const int levels2 = plan.linear ? num_levels : 1;
        if (in_dtype == RDVC_DT_F32) rc = launch_pack<float>(fmap1, fmap2, plan.a_km, plan.b_km, B, D, h, w, levels2, layout, plan.twl, plan.thl, plan.nl_of, key.f16_ops, st);
        else if (in_dtype == RDVC_DT_BF16) rc = launch_pack<__nv_bfloat16>(fmap1, fmap2, plan.a_km, plan.b_km, B, D, h, w, levels2, layout, plan.twl, plan.thl, plan.nl_of, key.f16_ops, st);
        else rc = launch_pack<__half>(fmap1, fmap2, plan.a_km, plan.b_km, B, D, h, w, levels2, layout, plan.twl, plan.thl, plan.nl_of, key.f16_ops, st);
        if (rc) return rc;
    }
    // 2. the GEMM
    const CUtensorMap& ta = plan.ta;
    const CUtensorMap* tb = plan.tb;
    const CUtensorMap* to = plan.to;
    const rdvc::BuildParams& p = plan.p;
    using rdvc::MODE_LINEAR;
#ifdef RDVC_EXPERIMENTS
    using rdvc::MODE_FUSED;
    if (plan.pair)
        return (vol_dtype == RDVC_DT_F32) ? launch_build_pair<float>(ta, plan.tb2, to, plan.p2, st)
                                          : launch_build_pair<__nv_bfloat16>(ta, plan.tb2, to, plan.p2, st);
    if (!plan.linear) {
        if (vol_dtype == RDVC_DT_F32)
            return (plan.tile == 1) ? launch_build<MODE_FUSED, 16, 16, float, 8>(ta, tb, to, p, st)
                                    : launch_build<MODE_FUSED, 8, 32, float, 8>(ta, tb, to, p, st);
        return (plan.tile == 1) ? launch_build<MODE_FUSED, 16, 16, __nv_bfloat16, 8>(ta, tb, to, p, st)
                                : launch_build<MODE_FUSED, 8, 32, __nv_bfloat16, 8>(ta, tb, to, p, st);
    }
#endif
    // epilogue shape per storage type: see BuildCfg (option key 9: 0 = auto, 4 / 8 = force)
#ifdef RDVC_EXPERIMENTS
    if (plan.cluster) {
        if (vol_dtype == RDVC_DT_F32)
            return plan.ew == 8 ? launch_build<MODE_LINEAR, 16, 16, float, 8, 2>(ta, tb, to, p, st)
                                : launch_build<MODE_LINEAR, 16, 16, float, 4, 2>(ta, tb, to, p, st);
        return plan.ew == 8 ? launch_build<MODE_LINEAR, 16, 16, __nv_bfloat16, 8, 2>(ta, tb, to, p, st)
                            : launch_build<MODE_LINEAR, 16, 16, __nv_bfloat16, 4, 2>(ta, tb, to, p, st);
    }
#endif
    if (vol_dtype == RDVC_DT_F32)
        return plan.ew == 8 ? launch_build<MODE_LINEAR, 16, 16, float, 8>(ta, tb, to, p, st)
                            : launch_build<MODE_LINEAR, 16, 16, float, 4>(ta, tb, to, p, st);
    return plan.ew == 8 ? launch_build<MODE_LINEAR, 16, 16, __nv_bfloat16, 8>(ta, tb, to, p, st)
                        : launch_build<MODE_LINEAR, 16, 16, __nv_bfloat16, 4>(ta, tb, to, p, st);
}
}  // namespace

extern "C" {

int rdvc_corr_build(const void* fmap1, const void* fmap2, int B, int D, int h, int w, int in_dtype,
                    void* pyramid, int vol_dtype, int layout, int num_levels, void* workspace,
                    size_t workspace_bytes, void* stream) {
    if (!fmap1 || !fmap2) return fail(RDVC_E_NULL, "null pointer argument");
    return build_impl(fmap1, fmap2, B, D, h, w, in_dtype, pyramid, vol_dtype, layout, num_levels, workspace,
                      workspace_bytes, stream);
}

int rdvc_corr_build_packed(int B, int D, int h, int w, int op_dtype, void* pyramid, int vol_dtype, int layout,
                           int num_levels, void* workspace, size_t workspace_bytes, void* stream) {
    if (op_dtype != RDVC_DT_BF16 && op_dtype != RDVC_DT_F16)
        return fail(RDVC_E_DTYPE, "packed operands are BF16 or F16, got op_dtype=%d", op_dtype);
    return build_impl(nullptr, nullptr, B, D, h, w, op_dtype, pyramid, vol_dtype, layout, num_levels, workspace,
                      workspace_bytes, stream);
}

int rdvc_corr_pack(const void* x1, const void* x2, int B, int D, int h, int w, int in_dtype, int vol_dtype,
                   int layout, int num_levels, void* workspace, size_t workspace_bytes, void* stream) {
    if (!x1 || !x2 || !workspace) return fail(RDVC_E_NULL, "null pointer argument");
    int rc = check_geometry(B, h, w, num_levels);
    if (rc) return rc;
    if (D <= 0 || D % 64 != 0 || D > 64 * rdvc::BLD_MAX_KC)
        return fail(RDVC_E_UNSUPPORTED, "D=%d must be a multiple of 64 and <= %d", D, 64 * rdvc::BLD_MAX_KC);
    if (in_dtype != RDVC_DT_F32 && in_dtype != RDVC_DT_BF16 && in_dtype != RDVC_DT_F16)
        return fail(RDVC_E_DTYPE, "unsupported in_dtype=%d", in_dtype);
    if (vol_dtype != RDVC_DT_F32 && vol_dtype != RDVC_DT_BF16)
        return fail(RDVC_E_DTYPE, "unsupported vol_dtype=%d", vol_dtype);
    if (layout != RDVC_LAYOUT_ROWMAJOR && layout != RDVC_LAYOUT_TILED)
        return fail(RDVC_E_UNSUPPORTED, "unknown pyramid layout %d", layout);
    if (workspace_bytes < rdvc_corr_workspace_bytes(B, D, h, w))
        return fail(RDVC_E_WORKSPACE, "workspace %zu < required %zu", workspace_bytes, rdvc_corr_workspace_bytes(B, D, h, w));
    if (reinterpret_cast<uintptr_t>(workspace) & 255) return fail(RDVC_E_ALIGN, "workspace must be 256-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    WorkspaceLayout wl;
    workspace_layout(workspace, B, D, h, w, vol_dtype, layout, num_levels, &wl);
    if (wl.memset_bytes) {
        cudaError_t e = cudaMemsetAsync(wl.memset_ptr, 0, wl.memset_bytes, st);
        if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(padded operand rows)");
    }
    const int f16 = (in_dtype == RDVC_DT_F16) ? 1 : 0;
    if (in_dtype == RDVC_DT_F32) return launch_pack<float>(x1, x2, wl.a_km, wl.b_km, B, D, h, w, num_levels, layout, wl.twl, wl.thl, wl.nl_of, f16, st);
    if (in_dtype == RDVC_DT_BF16) return launch_pack<__nv_bfloat16>(x1, x2, wl.a_km, wl.b_km, B, D, h, w, num_levels, layout, wl.twl, wl.thl, wl.nl_of, f16, st);
    return launch_pack<__half>(x1, x2, wl.a_km, wl.b_km, B, D, h, w, num_levels, layout, wl.twl, wl.thl, wl.nl_of, f16, st);
}

size_t rdvc_linear_packed_weight_bytes(int cout, int cin) {
    if (cout <= 0 || cin <= 0) return 0;
    return static_cast<size_t>(cout) * cin * 2;
}

int rdvc_linear_pack_weights(const float* weight, int cout, int cin, int dtype, void* packed_host) {
    if (!weight || !packed_host) return fail(RDVC_E_NULL, "null pointer argument");
    if (dtype != RDVC_DT_BF16 && dtype != RDVC_DT_F16) return fail(RDVC_E_DTYPE, "dtype must be BF16 or F16, got %d", dtype);
    if (cout <= 0 || cin <= 0) return fail(RDVC_E_SHAPE, "non-positive dimension cout=%d cin=%d", cout, cin);
    uint16_t* out = static_cast<uint16_t*>(packed_host);
    for (size_t i = 0; i < static_cast<size_t>(cout) * cin; ++i) {
        if (dtype == RDVC_DT_F16) { const __half hv = __float2half_rn(weight[i]); memcpy(out + i, &hv, 2); }
        else { const __nv_bfloat16 bv = __float2bfloat16_rn(weight[i]); memcpy(out + i, &bv, 2); }
    }
    return RDVC_OK;
}

int rdvc_corr_encoder_tail(const void* ws_in, size_t ws_in_bytes, int D_in, const void* packed_w, const float* bias,
                           int D_out, int B, int h, int w, int op_dtype, int vol_dtype, int layout, int num_levels,
                           void* ws_out, size_t ws_out_bytes, void* stream) {
    if (!ws_in || !packed_w || !ws_out) return fail(RDVC_E_NULL, "null pointer argument");
    int rc = check_geometry(B, h, w, num_levels);
    if (rc) return rc;
    if (D_in != rdvc::ET_K || D_out != rdvc::ET_N)
        return fail(RDVC_E_UNSUPPORTED, "the encoder tail is a %d -> %d convolution, got %d -> %d", rdvc::ET_K, rdvc::ET_N, D_in, D_out);
    if (op_dtype != RDVC_DT_BF16 && op_dtype != RDVC_DT_F16)
        return fail(RDVC_E_DTYPE, "operands are BF16 or F16, got op_dtype=%d", op_dtype);
    if (vol_dtype != RDVC_DT_F32 && vol_dtype != RDVC_DT_BF16)
        return fail(RDVC_E_DTYPE, "unsupported vol_dtype=%d", vol_dtype);
    if (layout != RDVC_LAYOUT_ROWMAJOR && layout != RDVC_LAYOUT_TILED)
        return fail(RDVC_E_UNSUPPORTED, "unknown pyramid layout %d", layout);
    if (ws_in_bytes < rdvc_corr_workspace_bytes(B, D_in, h, w) || ws_out_bytes < rdvc_corr_workspace_bytes(B, D_out, h, w))
        return fail(RDVC_E_WORKSPACE, "workspaces must hold rdvc_corr_workspace_bytes for D = %d (in) and D = %d (out)", D_in, D_out);
    if ((reinterpret_cast<uintptr_t>(ws_in) & 255) || (reinterpret_cast<uintptr_t>(ws_out) & 255) ||
        (reinterpret_cast<uintptr_t>(packed_w) & 15))
        return fail(RDVC_E_ALIGN, "workspaces must be 256-byte aligned, packed weights 16-byte aligned");
    WorkspaceLayout li, lo;
    workspace_layout(const_cast<void*>(ws_in), B, D_in, h, w, vol_dtype, layout, num_levels, &li);
    workspace_layout(ws_out, B, D_out, h, w, vol_dtype, layout, num_levels, &lo);
    rdvc::EncoderTailParams p;
    memset(&p, 0, sizeof(p));
    const uint8_t* in0 = static_cast<const uint8_t*>(ws_in);
    const uint8_t* out0 = static_cast<const uint8_t*>(ws_out);
    const size_t in_row = static_cast<size_t>(D_in) * 2, out_row = static_cast<size_t>(D_out) * 2;
    int tiles = 0;
    auto add_seg = [&](const void* src, const void* dst, long long rows, int img, int hl, int wl_, int tiles_w) {
        const int s = p.n_segs++;
        p.in_row0[s] = static_cast<long long>((static_cast<const uint8_t*>(src) - in0) / in_row);
        p.out_row0[s] = static_cast<long long>((static_cast<const uint8_t*>(dst) - out0) / out_row);
        p.rows[s] = rows;
        p.tile0[s] = tiles;
        p.img[s] = img; p.hl[s] = hl; p.wl[s] = wl_; p.tiles_w[s] = tiles_w;
        tiles += static_cast<int>((rows + rdvc::ET_BLOCK_M - 1) / rdvc::ET_BLOCK_M);
    };
    add_seg(li.a_km, lo.a_km, static_cast<long long>(B) * h * w, h * w, h, w, 0);                 // image 1: every row real
    for (int l = 0; l < num_levels; ++l) {
        const int hl = h >> l, wl_ = w >> l;
        const bool padded = li.nl_of[l] != static_cast<size_t>(hl) * wl_;
        const int tiles_w = (layout == RDVC_LAYOUT_TILED && padded) ? ((wl_ + (1 << li.twl) - 1) >> li.twl) : 0;
        add_seg(li.b_km[l], lo.b_km[l], static_cast<long long>(B) * static_cast<long long>(li.nl_of[l]),
                static_cast<int>(li.nl_of[l]), hl, wl_, tiles_w);
    }
    p.tile0[p.n_segs] = tiles;
    p.n_tiles = tiles;
    p.twl = li.twl; p.thl = li.thl;
    p.out = ws_out; p.bias = bias; p.ab_format = (op_dtype == RDVC_DT_F16) ? 0 : 1;
    p.dbg_timeline = RDVC_HAS_EXPERIMENTS ? g_dbg_timeline.load() : nullptr;
    const CUtensorMapDataType op_dt = (op_dtype == RDVC_DT_F16) ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    CUtensorMap tm_a, tm_w;
    {
        const cuuint64_t total_rows = ws_in_bytes / in_row;
        if (total_rows >= (1ull << 31)) return fail(RDVC_E_UNSUPPORTED, "too many rows for 32-bit TMA coordinates");
        cuuint64_t dims[3] = {(cuuint64_t)D_in, total_rows, 1};
        cuuint64_t str[2] = {(cuuint64_t)in_row, total_rows * in_row};
        cuuint32_t box[3] = {64, rdvc::ET_BLOCK_M, 1};
        if ((rc = make_tmap(&tm_a, op_dt, const_cast<void*>(ws_in), 3, dims, str, box))) return rc;
    }
    {
        cuuint64_t dims[3] = {(cuuint64_t)D_in, (cuuint64_t)D_out, 1};
        cuuint64_t str[2] = {(cuuint64_t)in_row, (cuuint64_t)D_out * in_row};
        cuuint32_t box[3] = {64, (cuuint32_t)D_out, 1};
        if ((rc = make_tmap(&tm_w, op_dt, const_cast<void*>(packed_w), 3, dims, str, box))) return rc;
    }
    auto kern = rdvc::corr_encoder_tail_kernel;
    static std::atomic<unsigned long long> attr_done{0};
    if ((rc = ensure_dynamic_smem(kern, rdvc::ET_SMEM_LAUNCH, attr_done, "cudaFuncSetAttribute(encoder tail, max dynamic smem)"))) return rc;
    int grid = sm_count();
    if (grid > tiles) grid = tiles;
    kern<<<grid, rdvc::ET_THREADS, rdvc::ET_SMEM_LAUNCH, static_cast<cudaStream_t>(stream)>>>(tm_a, tm_w, p);
    ++g_launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "corr_encoder_tail_kernel launch");
    return RDVC_OK;
}

size_t rdvc_corr_feat_pitch(int num_levels, int radius) {
    if (num_levels < 1 || num_levels > rdvc::LKP_MAX_LEVELS || radius < 1 || radius > 4) return 0;
    return static_cast<size_t>(rdvc::lkp_feat_pitch(num_levels, radius));
}

size_t rdvc_corr_feat_rows(int B, int h, int w) {
    if (B <= 0 || h <= 0 || w <= 0) return 0;
    return align_up(static_cast<size_t>(B) * h * w, 8);
}

size_t rdvc_corr_feat_bytes(int B, int h, int w, int num_levels, int radius) {
    return rdvc_corr_feat_pitch(num_levels, radius) * rdvc_corr_feat_rows(B, h, w) * 2;
}

int rdvc_corr_lookup_ex(const void* pyramid, int vol_dtype, int layout, const float* coords, int B, int h,
                        int w, int num_levels, int radius, void* out, int out_dtype, int out_form, void* stream) {
    if (!pyramid || !coords || !out) return fail(RDVC_E_NULL, "null pointer argument");
    int rc = check_geometry(B, h, w, num_levels);
    if (rc) return rc;
    if (vol_dtype != RDVC_DT_F32 && vol_dtype != RDVC_DT_BF16)
        return fail(RDVC_E_DTYPE, "unsupported vol_dtype=%d", vol_dtype);
    if (layout != RDVC_LAYOUT_ROWMAJOR && layout != RDVC_LAYOUT_TILED)
        return fail(RDVC_E_UNSUPPORTED, "unknown pyramid layout %d", layout);
    if (reinterpret_cast<uintptr_t>(pyramid) & 15)
        return fail(RDVC_E_ALIGN, "pyramid must be 16-byte aligned");
    if (radius < 1 || radius > 4) return fail(RDVC_E_UNSUPPORTED, "radius=%d not in [1, 4]", radius);
    int out_mode;
    if (out_form == RDVC_OUT_NCHW) {
        if (out_dtype == RDVC_DT_F32) out_mode = rdvc::LKP_OUT_NCHW_F32;
        else if (out_dtype == RDVC_DT_F16) out_mode = rdvc::LKP_OUT_NCHW_F16;
        else return fail(RDVC_E_DTYPE, "NCHW lookup output must be F32 or F16, got out_dtype=%d", out_dtype);
    } else if (out_form == RDVC_OUT_KMAJOR) {
        if (out_dtype == RDVC_DT_BF16) out_mode = rdvc::LKP_OUT_KM_BF16;
        else if (out_dtype == RDVC_DT_F16) out_mode = rdvc::LKP_OUT_KM_F16;
        else return fail(RDVC_E_DTYPE, "K-major feature rows must be BF16 or F16, got out_dtype=%d", out_dtype);
        if (reinterpret_cast<uintptr_t>(out) & 15) return fail(RDVC_E_ALIGN, "feature rows must be 16-byte aligned");
    } else {
        return fail(RDVC_E_UNSUPPORTED, "unknown lookup output form %d", out_form);
    }
    if (out_mode != rdvc::LKP_OUT_NCHW_F32 && layout != RDVC_LAYOUT_TILED)
        return fail(RDVC_E_UNSUPPORTED, "fp16 / K-major lookup outputs need RDVC_LAYOUT_TILED");
    rdvc::LookupParams p;
    memset(&p, 0, sizeof(p));
    for (int l = 0; l < num_levels; ++l) {
        p.lvl[l] = static_cast<const uint8_t*>(pyramid) + rdvc_corr_level_offset_bytes(B, h, w, l, vol_dtype, layout);
        p.hl[l] = h >> l;
        p.wl[l] = w >> l;
        p.img[l] = static_cast<long long>(level_image_elems(h, w, l, vol_dtype, layout));
    }
    if (layout == RDVC_LAYOUT_TILED) {
        tile_log2(vol_dtype, &p.twl, &p.thl);
        for (int l = 0; l < num_levels; ++l) p.tiles_w[l] = ((w >> l) + (1 << p.twl) - 1) >> p.twl;
    }
    p.coords = coords;
    p.out = out;
    p.feat_pitch = rdvc::lkp_feat_pitch(num_levels, radius);
    p.feat_rows = static_cast<long long>(rdvc_corr_feat_rows(B, h, w));
    p.B = B;
    p.N = h * w;
    p.num_levels = num_levels;
    p.total = static_cast<long long>(B) * p.N;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int variant = g_opt_lookup.load();
    if (layout == RDVC_LAYOUT_TILED) {
        const int dbg = variant >= 3 ? variant - 2 : 0;
        return (vol_dtype == RDVC_DT_F32) ? dispatch_lookup_tiled<float>(out_mode, dbg, radius, p, st)
                                          : dispatch_lookup_tiled<__nv_bfloat16>(out_mode, dbg, radius, p, st);
    }
    if (vol_dtype == RDVC_DT_F32) {
        if (variant == 1) return dispatch_lookup_radius<float, 0>(radius, p, st);
        return dispatch_lookup_radius<float, 1>(radius, p, st);
    }
    return dispatch_lookup_radius<__nv_bfloat16, 0>(radius, p, st);
}

int rdvc_corr_lookup(const void* pyramid, int vol_dtype, int layout, const float* coords, int B, int h,
                     int w, int num_levels, int radius, float* out, void* stream) {
    return rdvc_corr_lookup_ex(pyramid, vol_dtype, layout, coords, B, h, w, num_levels, radius, out, RDVC_DT_F32,
                               RDVC_OUT_NCHW, stream);
}

// ---- next row f-1: lookup + MotionEncoder.convcorr1 ------------------------------------------
size_t rdvc_conv1x1_packed_weight_bytes(int cout, int num_levels, int radius) {
    const size_t kp = rdvc_corr_feat_pitch(num_levels, radius);
    if (kp == 0 || cout <= 0 || cout > rdvc::C1_MAX_COUT || cout % 32 != 0) return 0;
    return static_cast<size_t>(cout) * align_up(kp, 64) * 2;
}

int rdvc_conv1x1_pack_weights(const float* weight, int cout, int num_levels, int radius, int feat_dtype,
                              void* packed_host) {
    if (!weight || !packed_host) return fail(RDVC_E_NULL, "null pointer argument");
    if (feat_dtype != RDVC_DT_BF16 && feat_dtype != RDVC_DT_F16)
        return fail(RDVC_E_DTYPE, "feat_dtype must be BF16 or F16, got %d", feat_dtype);
    const size_t kp = rdvc_corr_feat_pitch(num_levels, radius);
    if (kp == 0) return fail(RDVC_E_UNSUPPORTED, "num_levels=%d / radius=%d not supported", num_levels, radius);
    if (cout <= 0 || cout > rdvc::C1_MAX_COUT || cout % 32 != 0)
        return fail(RDVC_E_UNSUPPORTED, "cout=%d must be a multiple of 32 and <= %d", cout, rdvc::C1_MAX_COUT);
    const int S = 2 * radius + 1, PL = rdvc::lkp_level_pitch(radius);
    const size_t kpw = align_up(kp, 64);
    const int cin = num_levels * S * S;
    uint16_t* out = static_cast<uint16_t*>(packed_host);
    memset(out, 0, static_cast<size_t>(cout) * kpw * 2);
    for (int n = 0; n < cout; ++n)
        for (int l = 0; l < num_levels; ++l)
            for (int i = 0; i < S; ++i)          // torchvision channel = l*S*S + i*S + j  (i moves x, TV:raft.py:404-412)
                for (int j = 0; j < S; ++j) {
                    const float v = weight[static_cast<size_t>(n) * cin + l * S * S + i * S + j];
                    uint16_t bits;
                    if (feat_dtype == RDVC_DT_F16) { const __half hv = __float2half_rn(v); memcpy(&bits, &hv, 2); }
                    else { const __nv_bfloat16 bv = __float2bfloat16_rn(v); memcpy(&bits, &bv, 2); }
                    out[static_cast<size_t>(n) * kpw + l * PL + j * S + i] = bits;   // the lookup's column order
                }
    return RDVC_OK;
}

int rdvc_conv1x1(const void* feat, int feat_dtype, const void* packed_w, const float* bias, int B, int h, int w,
                 int num_levels, int radius, int cout, int act, void* out, int out_dtype, void* stream) {
    if (!feat || !packed_w || !out) return fail(RDVC_E_NULL, "null pointer argument");
    if (B <= 0 || h <= 0 || w <= 0) return fail(RDVC_E_SHAPE, "non-positive dimension B=%d h=%d w=%d", B, h, w);
    if (feat_dtype != RDVC_DT_BF16 && feat_dtype != RDVC_DT_F16)
        return fail(RDVC_E_DTYPE, "feat_dtype must be BF16 or F16, got %d", feat_dtype);
    if (out_dtype != RDVC_DT_F32 && out_dtype != RDVC_DT_F16 && out_dtype != RDVC_DT_BF16)
        return fail(RDVC_E_DTYPE, "unsupported out_dtype=%d", out_dtype);
    if (act != RDVC_ACT_NONE && act != RDVC_ACT_RELU) return fail(RDVC_E_UNSUPPORTED, "unknown activation %d", act);
    const size_t kp = rdvc_corr_feat_pitch(num_levels, radius);
    if (kp == 0) return fail(RDVC_E_UNSUPPORTED, "num_levels=%d / radius=%d not supported", num_levels, radius);
    if (cout <= 0 || cout > rdvc::C1_MAX_COUT || cout % 32 != 0)
        return fail(RDVC_E_UNSUPPORTED, "cout=%d must be a multiple of 32 and <= %d", cout, rdvc::C1_MAX_COUT);
    if ((reinterpret_cast<uintptr_t>(feat) & 15) || (reinterpret_cast<uintptr_t>(packed_w) & 15))
        return fail(RDVC_E_ALIGN, "feature rows and packed weights must be 16-byte aligned");
    const long long m_total = static_cast<long long>(B) * h * w;
    if (m_total >= (1LL << 31) - 256) return fail(RDVC_E_UNSUPPORTED, "too many pixels for 32-bit TMA coordinates");
    const size_t kpw = align_up(kp, 64);
    const CUtensorMapDataType op_dt = (feat_dtype == RDVC_DT_F16) ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    CUtensorMap tm_a, tm_w;
    {
        // chunk-major features [kp / 8][rows][8]: 8 consecutive rows of one chunk are 128 contiguous bytes, so the map
        // is {64 elements, rows / 8, chunks}; a box of {64, 16, 8} = 8 chunks x 128 rows lands in shared memory as
        // [chunk][row][16 B], the un-swizzled K-major core-matrix layout (corr_conv1x1_sm100.cuh)
        const cuuint64_t rows = (cuuint64_t)rdvc_corr_feat_rows(B, h, w);
        cuuint64_t dims[3] = {64, rows / 8, (cuuint64_t)kp / 8};
        cuuint64_t str[2] = {128, rows * 16};
        cuuint32_t box[3] = {64, rdvc::C1_BLOCK_M / 8, rdvc::C1_BLOCK_K / 8};
        if (int rc = make_tmap(&tm_a, op_dt, const_cast<void*>(feat), 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_NONE)) return rc;
    }
    {
        cuuint64_t dims[3] = {(cuuint64_t)kpw, (cuuint64_t)cout, 1};
        cuuint64_t str[2] = {(cuuint64_t)kpw * 2, (cuuint64_t)cout * kpw * 2};
        cuuint32_t box[3] = {rdvc::C1_BLOCK_K, (cuuint32_t)(cout / 2), 1};
        if (int rc = make_tmap(&tm_w, op_dt, const_cast<void*>(packed_w), 3, dims, str, box)) return rc;
    }
    rdvc::Conv1x1Params p;
    memset(&p, 0, sizeof(p));
    p.out = out; p.bias = bias; p.m_total = m_total; p.n_pix = h * w; p.cout = cout; p.kp = static_cast<int>(kp);
    p.relu = (act == RDVC_ACT_RELU); p.ab_format = (feat_dtype == RDVC_DT_F16) ? 0 : 1;
    p.dbg_timeline = RDVC_HAS_EXPERIMENTS ? g_dbg_timeline.load() : nullptr;
    p.dbg_flags = RDVC_HAS_EXPERIMENTS ? g_dbg_c1_flags.load() : 0;
    p.vec4 = ((h * w) % 4 == 0) && !(reinterpret_cast<uintptr_t>(out) & 15);
    // every CTA owns a contiguous range of pixel rows (a multiple of 32, one warp's rows), as equal as possible
    const int G = sm_count() & ~1;
    if (G < 2) return fail(RDVC_E_UNSUPPORTED, "the 1x1 convolution runs on CTA pairs: needs at least 2 SMs");
    long long rows = (m_total + G - 1) / G;
    rows = (rows + 31) / 32 * 32;
    p.rows_per_cta = static_cast<int>(rows);
    p.tiles_per_cta = static_cast<int>((rows + rdvc::C1_BLOCK_M - 1) / rdvc::C1_BLOCK_M);
    long long grid = (m_total + rows - 1) / rows;
    grid = (grid + 1) & ~1LL;                                  // whole CTA pairs (a trailing CTA may own no rows)
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaError_t e;
    if (out_dtype == RDVC_DT_F32) e = launch_conv1x1<float>(tm_a, tm_w, p, static_cast<unsigned>(grid), st);
    else if (out_dtype == RDVC_DT_F16) e = launch_conv1x1<__half>(tm_a, tm_w, p, static_cast<unsigned>(grid), st);
    else e = launch_conv1x1<__nv_bfloat16>(tm_a, tm_w, p, static_cast<unsigned>(grid), st);
    if (e != cudaSuccess) return cuda_fail(e, "corr_conv1x1_kernel launch");
    return RDVC_OK;
}

int rdvc_corr_lookup_conv1x1(const void* pyramid, int vol_dtype, int layout, const float* coords, int B, int h,
                             int w, int num_levels, int radius, const void* packed_w, const float* bias, int cout,
                             int act, int feat_dtype, void* feat_ws, size_t feat_ws_bytes, void* out, int out_dtype,
                             void* stream) {
    if (!feat_ws) return fail(RDVC_E_NULL, "null pointer argument");
    const size_t kp = rdvc_corr_feat_pitch(num_levels, radius);
    if (kp == 0) return fail(RDVC_E_UNSUPPORTED, "num_levels=%d / radius=%d not supported", num_levels, radius);
    if (B <= 0 || h <= 0 || w <= 0) return fail(RDVC_E_SHAPE, "non-positive dimension B=%d h=%d w=%d", B, h, w);
    const size_t need = rdvc_corr_feat_bytes(B, h, w, num_levels, radius);
    if (feat_ws_bytes < need) return fail(RDVC_E_WORKSPACE, "feature workspace %zu < required %zu", feat_ws_bytes, need);
    int rc = rdvc_corr_lookup_ex(pyramid, vol_dtype, layout, coords, B, h, w, num_levels, radius, feat_ws, feat_dtype,
                                 RDVC_OUT_KMAJOR, stream);
    if (rc) return rc;
    return rdvc_conv1x1(feat_ws, feat_dtype, packed_w, bias, B, h, w, num_levels, radius, cout, act, out, out_dtype, stream);
}

int rdvc_motion_warp(const float* prev, const float* flow, int B, int C, int H, int W, int h_in, int w_in,
                     float* warped, float* flow_out, void* stream) {
    if (!flow) return fail(RDVC_E_NULL, "null flow pointer");
    if ((prev == nullptr) != (warped == nullptr))
        return fail(RDVC_E_NULL, "prev and warped must both be given (warp) or both be NULL (resize only)");
    if (!warped && !flow_out) return fail(RDVC_E_NULL, "nothing to compute: warped and flow_out are both NULL");
    if (B <= 0 || H <= 0 || W <= 0 || h_in <= 0 || w_in <= 0 || (prev && C <= 0))
        return fail(RDVC_E_SHAPE, "non-positive dimension B=%d C=%d H=%d W=%d h_in=%d w_in=%d", B, C, H, W, h_in, w_in);
    if (H > 65535 || B > 65535) return fail(RDVC_E_UNSUPPORTED, "H and B must be <= 65535 (grid limits)");
    if (static_cast<long long>(H) * W >= (1LL << 31))
        return fail(RDVC_E_UNSUPPORTED, "H * W must fit 31 bits (in-plane offsets are 32-bit)");
    rdvc::WarpParams p;
    memset(&p, 0, sizeof(p));
    p.prev = prev; p.flow = flow; p.warped = warped; p.flow_out = flow_out;
    p.B = B; p.C = C; p.H = H; p.W = W; p.h_in = h_in; p.w_in = w_in;
    p.ry = static_cast<float>(h_in) / static_cast<float>(H);
    p.rx = static_cast<float>(w_in) / static_cast<float>(W);
    p.sh = static_cast<float>(static_cast<double>(H) / h_in);
    p.sw = static_cast<float>(static_cast<double>(W) / w_in);
    p.same_size = (h_in == H && w_in == W);
    const int per_block = rdvc::WARP_THREADS * rdvc::WARP_PIX;
    dim3 grid((W + per_block - 1) / per_block, H, B);
    rdvc::motion_warp_kernel<<<grid, rdvc::WARP_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(p);
    ++g_launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "motion_warp_kernel launch");
    return RDVC_OK;
}

int rdvc_preprocess_frame(const unsigned char* frame_hwc, int H, int W, int C, float* out, int h_out, int w_out,
                          void* stream) {
    if (!frame_hwc || !out) return fail(RDVC_E_NULL, "null pointer argument");
    if (H <= 0 || W <= 0 || h_out <= 0 || w_out <= 0) return fail(RDVC_E_SHAPE, "non-positive dimension H=%d W=%d h_out=%d w_out=%d", H, W, h_out, w_out);
    if (C < 1 || C > 4) return fail(RDVC_E_UNSUPPORTED, "C=%d not in [1, 4]", C);
    if (h_out > 65535) return fail(RDVC_E_UNSUPPORTED, "h_out must be <= 65535 (grid limit)");
    const double sy = static_cast<double>(H) / h_out, sx = static_cast<double>(W) / w_out;
    if (2.0 * (sy > 1 ? sy : 1) + 2 > rdvc::PREP_MAX_TAPS || 2.0 * (sx > 1 ? sx : 1) + 2 > rdvc::PREP_MAX_TAPS)
        return fail(RDVC_E_UNSUPPORTED, "down-scaling by more than %dx is not supported", (rdvc::PREP_MAX_TAPS - 2) / 2);
    rdvc::PrepParams p;
    memset(&p, 0, sizeof(p));
    p.src = frame_hwc; p.dst = out; p.H = H; p.W = W; p.C = C; p.h_out = h_out; p.w_out = w_out;
    p.sy = static_cast<float>(H) / static_cast<float>(h_out);    // aten: area_pixel_compute_scale, fp32
    p.sx = static_cast<float>(W) / static_cast<float>(w_out);
    dim3 grid((w_out + 127) / 128, h_out);
    rdvc::preprocess_kernel<<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(p);
    ++g_launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "preprocess_kernel launch");
    return RDVC_OK;
}

void rdvc_corr_release(void) {
    for (auto& a : g_arena) release_arena(a);
}

int rdvc_corr_pair_host_submit_ex(const float* fmap1_host, const float* fmap2_host, const float* coords_host,
                                  void* out_host, int B, int D, int h, int w, int num_levels, int radius,
                                  int iters, int vol_dtype, int out_dtype, int slot) {
    if (!fmap1_host || !fmap2_host || !coords_host || !out_host) return fail(RDVC_E_NULL, "null pointer argument");
    if (slot < 0 || slot >= kHostSlots) return fail(RDVC_E_UNSUPPORTED, "slot=%d not in [0, %d)", slot, kHostSlots);
    // validated before any size arithmetic or allocation (a bad enum is an argument error, not a crash)
    if (vol_dtype != RDVC_DT_F32 && vol_dtype != RDVC_DT_BF16)
        return fail(RDVC_E_DTYPE, "unsupported vol_dtype=%d", vol_dtype);
    if (out_dtype != RDVC_DT_F32 && out_dtype != RDVC_DT_F16)
        return fail(RDVC_E_DTYPE, "out_dtype must be F32 or F16, got %d", out_dtype);
    int rc = check_geometry(B, h, w, num_levels);
    if (rc) return rc;
    if (iters <= 0 || D <= 0) return fail(RDVC_E_SHAPE, "iters=%d D=%d must be positive", iters, D);
    if (radius < 1 || radius > 4) return fail(RDVC_E_UNSUPPORTED, "radius=%d not in [1, 4]", radius);
    const size_t N = static_cast<size_t>(h) * w;
    const size_t S = 2 * radius + 1;
    const size_t oes = (out_dtype == RDVC_DT_F32) ? 4 : 2;
    const size_t fmap_bytes = align_up(static_cast<size_t>(B) * D * N * 4, 256);
    const size_t coords_bytes = align_up(static_cast<size_t>(B) * 2 * N * 4, 256);
    const size_t out_elems = static_cast<size_t>(B) * num_levels * S * S * N;
    const size_t out_bytes = align_up(out_elems * oes, 256);
    const int layout = (g_opt_mode.load() == 1) ? RDVC_LAYOUT_ROWMAJOR : RDVC_LAYOUT_TILED;
    const size_t pyr_bytes = rdvc_corr_pyramid_bytes(B, h, w, num_levels, vol_dtype, layout);
    const size_t ws_bytes = rdvc_corr_workspace_bytes(B, D, h, w);
    const size_t need = 2 * fmap_bytes + iters * coords_bytes + kOutBufs * out_bytes + pyr_bytes + ws_bytes;

    HostArena& a = g_arena[slot];
    cudaError_t e;
    if (a.pending) return fail(RDVC_E_UNSUPPORTED, "slot %d still has a submitted pair: call rdvc_corr_pair_host_wait first", slot);
    if (a.device != current_device()) {   // the caller switched GPUs: streams and scratch belong to the old one
        release_arena(a);
        a.device = current_device();
    }
    if (!a.compute) {
        if ((e = cudaStreamCreateWithFlags(&a.compute, cudaStreamNonBlocking)) != cudaSuccess) return cuda_fail(e, "stream create");
        if ((e = cudaStreamCreateWithFlags(&a.copy, cudaStreamNonBlocking)) != cudaSuccess) return cuda_fail(e, "stream create");
        for (auto& ev : a.ev)
            if ((e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)) != cudaSuccess) return cuda_fail(e, "event create");
        for (auto& ev : a.done)
            if ((e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)) != cudaSuccess) return cuda_fail(e, "event create");
    }
    if (a.bytes < need) {
        if (a.dev) cudaFree(a.dev);
        a.dev = nullptr; a.bytes = 0;
        if ((e = cudaMalloc(&a.dev, need)) != cudaSuccess) return cuda_fail(e, "cudaMalloc(scratch arena)");
        a.bytes = need;
    }
    uint8_t* base = static_cast<uint8_t*>(a.dev);
    uint8_t* d_f1 = base;
    uint8_t* d_f2 = d_f1 + fmap_bytes;
    uint8_t* d_co = d_f2 + fmap_bytes;
    uint8_t* d_out[kOutBufs];
    for (int k = 0; k < kOutBufs; ++k) d_out[k] = d_co + iters * coords_bytes + k * out_bytes;
    uint8_t* d_pyr = d_out[kOutBufs - 1] + out_bytes;
    uint8_t* d_ws = d_pyr + pyr_bytes;

    // inputs in (one copy each), then build
    const size_t fmap_raw = static_cast<size_t>(B) * D * N * 4;
    const size_t coords_raw = static_cast<size_t>(B) * 2 * N * 4;
    if ((e = cudaMemcpyAsync(d_f1, fmap1_host, fmap_raw, cudaMemcpyHostToDevice, a.compute)) != cudaSuccess) return cuda_fail(e, "H2D fmap1");
    if ((e = cudaMemcpyAsync(d_f2, fmap2_host, fmap_raw, cudaMemcpyHostToDevice, a.compute)) != cudaSuccess) return cuda_fail(e, "H2D fmap2");
    for (int it = 0; it < iters; ++it)
        if ((e = cudaMemcpyAsync(d_co + it * coords_bytes, coords_host + it * (coords_raw / 4), coords_raw,
                                 cudaMemcpyHostToDevice, a.compute)) != cudaSuccess) return cuda_fail(e, "H2D coords");
    a.pending = true;
    rc = rdvc_corr_build(d_f1, d_f2, B, D, h, w, RDVC_DT_F32, d_pyr, vol_dtype, layout, num_levels, d_ws, ws_bytes, a.compute);
    if (rc) return rc;
    // lookups rotate over kOutBufs device buffers; the copy stream drains them
    bool used[kOutBufs] = {};
    for (int it = 0; it < iters; ++it) {
        const int s = it % kOutBufs;
        if (used[s]) {  // the D2H that last read this buffer must have finished
            if ((e = cudaStreamWaitEvent(a.compute, a.ev[s], 0)) != cudaSuccess) return cuda_fail(e, "wait event");
        }
        rc = rdvc_corr_lookup_ex(d_pyr, vol_dtype, layout, reinterpret_cast<const float*>(d_co + it * coords_bytes), B, h, w,
                                 num_levels, radius, d_out[s], out_dtype, RDVC_OUT_NCHW, a.compute);
        if (rc) return rc;
        // lookup finished -> copy stream may read (the slot's own events, created once with the arena)
        if ((e = cudaEventRecord(a.done[s], a.compute)) != cudaSuccess) return cuda_fail(e, "event record");
        if ((e = cudaStreamWaitEvent(a.copy, a.done[s], 0)) != cudaSuccess) return cuda_fail(e, "wait event");
        if ((e = cudaMemcpyAsync(static_cast<uint8_t*>(out_host) + it * out_elems * oes, d_out[s], out_elems * oes,
                                 cudaMemcpyDeviceToHost, a.copy)) != cudaSuccess) return cuda_fail(e, "D2H out");
        cudaEventRecord(a.ev[s], a.copy);
        used[s] = true;
    }
    return RDVC_OK;
}

int rdvc_corr_pair_host_submit(const float* fmap1_host, const float* fmap2_host, const float* coords_host,
                               float* out_host, int B, int D, int h, int w, int num_levels, int radius,
                               int iters, int vol_dtype, int slot) {
    return rdvc_corr_pair_host_submit_ex(fmap1_host, fmap2_host, coords_host, out_host, B, D, h, w, num_levels, radius,
                                         iters, vol_dtype, RDVC_DT_F32, slot);
}

int rdvc_corr_pair_host_wait(int slot) {
    if (slot < 0 || slot >= kHostSlots) return fail(RDVC_E_UNSUPPORTED, "slot=%d not in [0, %d)", slot, kHostSlots);
    HostArena& a = g_arena[slot];
    if (!a.pending) return RDVC_OK;
    a.pending = false;
    cudaError_t e;
    if ((e = cudaStreamSynchronize(a.compute)) != cudaSuccess) return cuda_fail(e, "sync compute");
    if ((e = cudaStreamSynchronize(a.copy)) != cudaSuccess) return cuda_fail(e, "sync copy");
    return RDVC_OK;
}

int rdvc_corr_pair_host(const float* fmap1_host, const float* fmap2_host, const float* coords_host,
                        float* out_host, int B, int D, int h, int w, int num_levels, int radius,
                        int iters, int vol_dtype) {
    int rc = rdvc_corr_pair_host_wait(0);
    if (rc) return rc;
    rc = rdvc_corr_pair_host_submit(fmap1_host, fmap2_host, coords_host, out_host, B, D, h, w, num_levels, radius,
                                    iters, vol_dtype, 0);
    const int rw = rdvc_corr_pair_host_wait(0);
    return rc ? rc : rw;
}

}  // extern "C"

// ---- next row f-4 (cont.): motion compensation network ----------------------------------------
namespace {

int mcn_check_geometry(int B, int H, int W) {
    if (B <= 0 || H <= 0 || W <= 0) return fail(RDVC_E_SHAPE, "non-positive dimension B=%d H=%d W=%d", B, H, W);
    if (H > 65535 || B > 65535) return fail(RDVC_E_UNSUPPORTED, "H and B must be <= 65535 (grid limits)");
    if (static_cast<long long>(B) * ((H + 7) / 8) * (((W + 1) / 2 + rdvc::MCNX_TXO - 1) / rdvc::MCNX_TXO) >= (1LL << 31))
        return fail(RDVC_E_UNSUPPORTED, "too many tiles for 32-bit tile indices");
    return RDVC_OK;
}

// XH = false: mcn_conv_kernel (three boxes per tile); XH = true: mcn_convx_kernel (one box per tile, 14 output columns)
template <int R, int NOUT, int KPAT, bool XH>
int launch_mcn_conv_impl(const void* act_in, const void* packed_w, void* act_out, rdvc::McnConvParams& p, cudaStream_t st) {
    using Cfg = rdvc::McnCfg<R, NOUT>;
    constexpr int SMEM = XH ? rdvc::McnXCfg<R, NOUT>::SMEM_LAUNCH : Cfg::SMEM_LAUNCH;
    auto kern = XH ? rdvc::mcn_convx_kernel<R, NOUT, KPAT> : rdvc::mcn_conv_kernel<R, NOUT, KPAT>;
    static std::atomic<unsigned long long> attr_done{0};
    if (int rc = ensure_dynamic_smem(kern, SMEM, attr_done, "cudaFuncSetAttribute(mcn_conv, max dynamic smem)"))
        return rc;
    const int out_cols = XH ? rdvc::MCNX_TXO : rdvc::MCN_TX;
    p.ntx = (p.Wsp + out_cols - 1) / out_cols;
    const cuuint64_t Wsp = static_cast<cuuint64_t>(p.Wsp), H = static_cast<cuuint64_t>(p.H), B = static_cast<cuuint64_t>(p.B);
    CUtensorMap tm_in, tm_w, tm_out, tm_res;
    {
        const cuuint64_t dims[4] = {64, Wsp, H, B};
        const cuuint64_t strides[3] = {128, Wsp * 128, H * Wsp * 128};
        const cuuint32_t box_in[4] = {64, rdvc::MCN_TX, static_cast<cuuint32_t>(Cfg::BOX_ROWS), 1};
        if (int rc = make_tmap(&tm_in, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, const_cast<void*>(act_in), 4, dims, strides, box_in))
            return rc;
        // the last layer writes NCHW fp32 itself; its store map is a placeholder on the input tensor
        const cuuint32_t box_out[4] = {64, static_cast<cuuint32_t>(out_cols), 2, 1};
        if (int rc = make_tmap(&tm_out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, act_out ? act_out : const_cast<void*>(act_in), 4,
                               dims, strides, box_out))
            return rc;
        tm_res = tm_out;
        if (p.residual)
            if (int rc = make_tmap(&tm_res, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, const_cast<__half*>(p.residual), 4, dims,
                                   strides, box_out))
                return rc;
    }
    {
        const cuuint64_t dims[3] = {64, static_cast<cuuint64_t>(NOUT), static_cast<cuuint64_t>(Cfg::NTAPS)};
        const cuuint64_t strides[2] = {128, static_cast<cuuint64_t>(NOUT) * 128};
        const cuuint32_t box[3] = {64, static_cast<cuuint32_t>(NOUT), 1};
        if (int rc = make_tmap(&tm_w, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, const_cast<void*>(packed_w), 3, dims, strides, box))
            return rc;
    }
    long long grid = sm_count();
    const long long n_tiles = static_cast<long long>(p.B) * p.ntx * p.nty;
    if (grid > n_tiles) grid = n_tiles;
    kern<<<static_cast<unsigned>(grid), rdvc::MCN_THREADS, SMEM, st>>>(tm_in, tm_w, tm_out, tm_res, p);
    ++g_launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "mcn_conv_kernel launch");
    return RDVC_OK;
}

template <int R, int NOUT, int KPAT>
int launch_mcn_conv(const void* act_in, const void* packed_w, void* act_out, rdvc::McnConvParams& p, cudaStream_t st) {
    // auto (measured at 1080p, us per layer, three boxes / one box): 3x3 68 / 58, first layer 103 / 89 -- but with a
    // residual 77 / 81 and for the output layer 88 / 98 (the one-box epilogue does more work per tile)
    const int k = g_opt_mcn_kernel.load();
    const bool xh = (k == 2) || (k == 0 && NOUT == 64 && p.residual == nullptr);
    return xh ? launch_mcn_conv_impl<R, NOUT, KPAT, true>(act_in, packed_w, act_out, p, st)
              : launch_mcn_conv_impl<R, NOUT, KPAT, false>(act_in, packed_w, act_out, p, st);
}

int mcn_fill_params(rdvc::McnConvParams& p, int B, int H, int W, const float* bias, int nbias) {
    memset(&p, 0, sizeof(p));
    p.B = B; p.H = H; p.W = W; p.Wsp = (W + 1) / 2;
    p.ntx = (p.Wsp + rdvc::MCN_TX - 1) / rdvc::MCN_TX;
    p.nty = (H + rdvc::MCN_TY - 1) / rdvc::MCN_TY;
    if (bias) for (int i = 0; i < nbias; ++i) p.bias[i] = bias[i];
    p.prefetch_dist = g_opt_mcn_prefetch.load();
    return RDVC_OK;
}

}  // namespace

extern "C" {

size_t rdvc_mcn_plane_bytes(int B, int H, int W) {
    if (B <= 0 || H <= 0 || W <= 0) return 0;
    return align_up(static_cast<size_t>(B) * H * ((W + 1) / 2) * 128, 1024);
}

size_t rdvc_mcn_workspace_bytes(int B, int H, int W) { return 3 * rdvc_mcn_plane_bytes(B, H, W); }

size_t rdvc_mcn_packed_weight_bytes(int ksize, int cout) {
    if ((ksize != 3 && ksize != 5) || cout <= 0 || cout > rdvc::MCN_C) return 0;
    const int nout = (cout > 8) ? 64 : 16;
    return static_cast<size_t>(ksize) * 3 * nout * 64 * 2;
}

int rdvc_mcn_pack_weights(const float* weight, int cout, int cin, int ksize, void* packed_host,
                          unsigned long long* kmask) {
    if (!weight || !packed_host || !kmask) return fail(RDVC_E_NULL, "null pointer argument");
    if (ksize != 3 && ksize != 5) return fail(RDVC_E_UNSUPPORTED, "kernel size %d not in {3, 5}", ksize);
    if (cin <= 0 || cin > rdvc::MCN_C || cout <= 0 || cout > rdvc::MCN_C)
        return fail(RDVC_E_UNSUPPORTED, "channel counts must be in [1, %d], got cin=%d cout=%d", rdvc::MCN_C, cin, cout);
    const int R = ksize / 2, nout = (cout > 8) ? 64 : 16, co_n = nout / 2;
    __half* out = static_cast<__half*>(packed_host);
    unsigned long long mask = 0;
    for (int dy = -R; dy <= R; ++dy)
        for (int dsx = -1; dsx <= 1; ++dsx) {
            const int t = (dy + R) * 3 + (dsx + 1);
            for (int n = 0; n < nout; ++n) {
                const int q = n / co_n, co = n % co_n;
                for (int k = 0; k < 64; ++k) {
                    const int pp = k / rdvc::MCN_C, ci = k % rdvc::MCN_C;
                    const int dx = 2 * dsx + pp - q;
                    float v = 0.f;
                    if (co < cout && ci < cin && dx >= -R && dx <= R)
                        v = weight[((static_cast<size_t>(co) * cin + ci) * ksize + (dy + R)) * ksize + (dx + R)];
                    const __half hv = __float2half_rn(v);
                    out[(static_cast<size_t>(t) * nout + n) * 64 + k] = hv;
                    if (__half2float(hv) != 0.f) mask |= 1ull << (4 * t + k / 16);
                }
            }
        }
    *kmask = mask;
    return RDVC_OK;
}

int rdvc_mcn_pack_input(const float* warped, const float* flow, const float* ref, int B, int c_warped, int c_flow,
                        int c_ref, int H, int W, void* act, void* stream) {
    if (!warped || !flow || !ref || !act) return fail(RDVC_E_NULL, "null pointer argument");
    if (int rc = mcn_check_geometry(B, H, W)) return rc;
    if (c_warped < 0 || c_flow < 0 || c_ref < 0 || c_warped + c_flow + c_ref > 8 || c_warped + c_flow + c_ref <= 0)
        return fail(RDVC_E_UNSUPPORTED, "input channels %d + %d + %d must sum to 1..8", c_warped, c_flow, c_ref);
    if (reinterpret_cast<uintptr_t>(act) % 16) return fail(RDVC_E_ALIGN, "activation buffer must be 16-byte aligned");
    rdvc::McnInputParams p;
    memset(&p, 0, sizeof(p));
    p.src[0] = warped; p.src[1] = flow; p.src[2] = ref;
    p.ch[0] = c_warped; p.ch[1] = c_flow; p.ch[2] = c_ref;
    p.dst = static_cast<__half*>(act);
    p.B = B; p.H = H; p.W = W; p.Wsp = (W + 1) / 2;
    dim3 grid((2 * p.Wsp + 127) / 128, H, B);
    rdvc::mcn_pack_input_kernel<<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(p);
    ++g_launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "mcn_pack_input_kernel launch");
    return RDVC_OK;
}

int rdvc_mcn_conv(const void* act_in, const void* packed_weights, unsigned long long kmask, const float* bias,
                  int ksize, int act, const void* residual, void* act_out, int B, int H, int W, void* stream) {
    if (!act_in || !packed_weights || !act_out) return fail(RDVC_E_NULL, "null pointer argument");
    if (ksize != 3 && ksize != 5) return fail(RDVC_E_UNSUPPORTED, "kernel size %d not in {3, 5}", ksize);
    const int reverse = (act & RDVC_MCN_REVERSE_ORDER) ? 1 : 0;
    act &= ~RDVC_MCN_REVERSE_ORDER;
    if (act != RDVC_MCN_ACT_NONE && act != RDVC_MCN_ACT_LEAKY) return fail(RDVC_E_UNSUPPORTED, "unknown activation %d", act);
    if (int rc = mcn_check_geometry(B, H, W)) return rc;
    if (act_in == act_out) return fail(RDVC_E_UNSUPPORTED, "a layer cannot run in place (neighbouring tiles read its input)");
    if (reinterpret_cast<uintptr_t>(act_in) % 16 || reinterpret_cast<uintptr_t>(act_out) % 16 ||
        reinterpret_cast<uintptr_t>(packed_weights) % 16 || reinterpret_cast<uintptr_t>(residual) % 16)
        return fail(RDVC_E_ALIGN, "activation / weight buffers must be 16-byte aligned");
    rdvc::McnConvParams p;
    mcn_fill_params(p, B, H, W, bias, rdvc::MCN_C);
    p.act = act;
    p.reverse = reverse;
    p.residual = static_cast<const __half*>(residual);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // the cheapest compile-time MMA schedule that covers the non-zero k-steps of these weights
    auto covered = [&](int kpat) { return (kmask & ~rdvc::mcn_pattern_mask(kpat, ksize)) == 0; };
    if (ksize == 3) {
        if (covered(rdvc::MCN_K_3X3)) return launch_mcn_conv<1, 64, rdvc::MCN_K_3X3>(act_in, packed_weights, act_out, p, st);
        return launch_mcn_conv<1, 64, rdvc::MCN_K_FULL>(act_in, packed_weights, act_out, p, st);
    }
    if (covered(rdvc::MCN_K_C16)) return launch_mcn_conv<2, 64, rdvc::MCN_K_C16>(act_in, packed_weights, act_out, p, st);
    return launch_mcn_conv<2, 64, rdvc::MCN_K_FULL>(act_in, packed_weights, act_out, p, st);
}

int rdvc_mcn_conv_out(const void* act_in, const void* packed_weights, unsigned long long kmask, const float* bias,
                      int ksize, int cout, const float* warped, float* out, int B, int H, int W, void* stream) {
    if (!act_in || !packed_weights || !warped || !out) return fail(RDVC_E_NULL, "null pointer argument");
    if (ksize != 5) return fail(RDVC_E_UNSUPPORTED, "the output layer is 5x5 (got %d)", ksize);
    if (cout <= 0 || cout > 8) return fail(RDVC_E_UNSUPPORTED, "output channels %d not in [1, 8]", cout);
    if (int rc = mcn_check_geometry(B, H, W)) return rc;
    if (reinterpret_cast<uintptr_t>(act_in) % 16 || reinterpret_cast<uintptr_t>(packed_weights) % 16)
        return fail(RDVC_E_ALIGN, "activation / weight buffers must be 16-byte aligned");
    if ((W % 2 == 0) && (reinterpret_cast<uintptr_t>(warped) % 8 || reinterpret_cast<uintptr_t>(out) % 8))
        return fail(RDVC_E_ALIGN, "warped / out must be 8-byte aligned");
    rdvc::McnConvParams p;
    mcn_fill_params(p, B, H, W, bias, cout);
    p.cout = cout;
    p.warped = warped;
    p.out = out;
    (void)kmask;   // the output layer always runs the full schedule
    return launch_mcn_conv<2, 16, rdvc::MCN_K_FULL>(act_in, packed_weights, nullptr, p, static_cast<cudaStream_t>(stream));
}

int rdvc_mcn_forward(const float* warped, const float* flow, const float* ref, int B, int H, int W,
                     int num_res_blocks, const void* const* packed_weights, const unsigned long long* kmasks,
                     const float* biases, void* workspace, size_t workspace_bytes, float* out, void* stream) {
    if (!warped || !flow || !ref || !packed_weights || !kmasks || !biases || !workspace || !out)
        return fail(RDVC_E_NULL, "null pointer argument");
    if (num_res_blocks < 0 || num_res_blocks > 64) return fail(RDVC_E_UNSUPPORTED, "num_res_blocks=%d", num_res_blocks);
    if (int rc = mcn_check_geometry(B, H, W)) return rc;
    const size_t plane = rdvc_mcn_plane_bytes(B, H, W);
    if (workspace_bytes < 3 * plane)
        return fail(RDVC_E_WORKSPACE, "workspace too small: %zu < %zu bytes", workspace_bytes, 3 * plane);
    if (reinterpret_cast<uintptr_t>(workspace) % 256) return fail(RDVC_E_ALIGN, "workspace must be 256-byte aligned");
    char* ws = static_cast<char*>(workspace);
    void* t = ws;                 // network input, then every block's middle activation
    void* x = ws + plane;         // block input (kept for the residual)
    void* y = ws + 2 * plane;     // block output
    const int n_layers = 2 + 2 * num_res_blocks;
    if (int rc = rdvc_mcn_pack_input(warped, flow, ref, B, 3, 2, 3, H, W, t, stream)) return rc;
    // Successive layers walk the tiles in opposite directions: a plane (133 MB at 1080p) is about the size of
    // the L2, so the part of it the previous launch wrote LAST is still cached when the next launch starts --
    // reading it first turns those loads into L2 hits.  The input pack writes first-to-last, so layer 0 runs
    // last-to-first; the output layer (odd index) runs first-to-last.
    const int REV = RDVC_MCN_REVERSE_ORDER;
    if (int rc = rdvc_mcn_conv(t, packed_weights[0], kmasks[0], biases, 5, RDVC_MCN_ACT_LEAKY | REV, nullptr, x, B, H, W, stream))
        return rc;
    for (int r = 0; r < num_res_blocks; ++r) {
        const int l = 1 + 2 * r;
        if (int rc = rdvc_mcn_conv(x, packed_weights[l], kmasks[l], biases + l * 32, 3, RDVC_MCN_ACT_LEAKY, nullptr, t,
                                   B, H, W, stream))                                   // odd layer: first-to-last
            return rc;
        if (int rc = rdvc_mcn_conv(t, packed_weights[l + 1], kmasks[l + 1], biases + (l + 1) * 32, 3, RDVC_MCN_ACT_LEAKY | REV,
                                   x, y, B, H, W, stream))                             // even layer: last-to-first
            return rc;
        void* tmp = x; x = y; y = tmp;
    }
    return rdvc_mcn_conv_out(x, packed_weights[n_layers - 1], kmasks[n_layers - 1], biases + (n_layers - 1) * 32, 5, 3,
                             warped, out, B, H, W, stream);
}

}  // extern "C"

// ---- next row f-4 (last piece): entropy-coder stand-in (host only) ----------------------------
extern "C" {

size_t rdvc_ec_max_encoded_bytes(size_t n) { return rdvc::ec::max_encoded_bytes(n); }

size_t rdvc_ec_encode_with_indexes(const int* symbols, const int* indexes, size_t n, const unsigned int* cdfs,
                                   const int* cdf_lengths, const int* offsets, int n_tables, int max_len,
                                   unsigned char* out, size_t out_capacity) {
    if ((n && (!symbols || !indexes)) || !cdfs || !cdf_lengths || !offsets || !out) {
        fail(RDVC_E_NULL, "null pointer argument");
        return 0;
    }
    const size_t r = rdvc::ec::encode_with_indexes(symbols, indexes, n, cdfs, cdf_lengths, offsets, n_tables, max_len,
                                                   out, out_capacity);
    if (r == 0) fail(RDVC_E_UNSUPPORTED, "entropy encode failed: malformed CDF table / index, or output buffer too small");
    return r;
}

int rdvc_ec_decode_with_indexes(const unsigned char* in, size_t nbytes, const int* indexes, size_t n,
                                const unsigned int* cdfs, const int* cdf_lengths, const int* offsets, int n_tables,
                                int max_len, int* symbols_out) {
    if ((n && (!in || !indexes || !symbols_out)) || !cdfs || !cdf_lengths || !offsets)
        return fail(RDVC_E_NULL, "null pointer argument");
    if (rdvc::ec::decode_with_indexes(in, nbytes, indexes, n, cdfs, cdf_lengths, offsets, n_tables, max_len, symbols_out))
        return fail(RDVC_E_UNSUPPORTED, "entropy decode failed: truncated or malformed stream / table");
    return RDVC_OK;
}

}  // extern "C"
