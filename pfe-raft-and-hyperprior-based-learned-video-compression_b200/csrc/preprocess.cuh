// preprocess.cuh -- the frame preparation in front of RAFT and of the codec, on the GPU ("next" row: the
// caller side of the path):
//
//   preprocess_frame_raft   R:codec_processing.py:751-761  TF.to_tensor (uint8 HWC -> float CHW / 255) followed
//                           by TF.resize(..., antialias=True): bilinear with aten's anti-aliasing filter
//                           (called twice per P-frame at :1430-1431, on the CPU, then .to(device))
//   preprocess_frame_codec  R:codec_processing.py:763-769  TF.to_tensor only (:1450)
//
// One thread per output pixel, all channels: it forms the separable triangle-filter weights of aten's
// _upsample_bilinear2d_aa for its row and column
//     scale = in / out, support = max(scale, 1), centre = scale * (i + 0.5),
//     first = max(int(centre - support + 0.5), 0), count = min(int(centre + support + 0.5), in) - first,
//     w_j = max(0, 1 - |(j + first - centre + 0.5) / max(scale, 1)|), normalised to sum 1
// and accumulates the footprint straight from the uint8 frame.  The host uploads 1 byte per sample instead
// of 4, and the CPU no longer spends 15-25 ms per 1080p frame in to_tensor + resize.
#pragma once
#include <cstdint>

namespace rdvc {

struct PrepParams {
    const uint8_t* src;   // (H, W, C) uint8, HWC as decoded frames are
    float* dst;           // (C, h_out, w_out) fp32, values in [0, 1]
    int H, W, C, h_out, w_out;
    float sy, sx;         // H / h_out, W / w_out
};

constexpr int PREP_MAX_TAPS = 16;   // count <= 2 * support + 2: scale factors up to 7x down

__device__ __forceinline__ void prep_span(int i, float scale, int n_in, int& first, int& count, float& centre,
                                          float& invscale) {
    const float support = scale >= 1.f ? scale : 1.f;
    centre = scale * (static_cast<float>(i) + 0.5f);
    invscale = scale >= 1.f ? 1.f / scale : 1.f;
    first = max(static_cast<int>(centre - support + 0.5f), 0);
    count = min(static_cast<int>(centre + support + 0.5f), n_in) - first;
    count = min(count, PREP_MAX_TAPS);
}

// grid: (ceil(w_out / 128), h_out); block: 128 threads
__global__ void __launch_bounds__(128)
preprocess_kernel(const __grid_constant__ PrepParams p) {
    const int j = blockIdx.x * 128 + threadIdx.x, i = blockIdx.y;
    if (j >= p.w_out) return;
    int y0, ny, x0, nx;
    float cy, cx, iy, ix;
    prep_span(i, p.sy, p.H, y0, ny, cy, iy);
    prep_span(j, p.sx, p.W, x0, nx, cx, ix);
    auto tri = [](int k, int first, float centre, float inv) {
        return fmaxf(0.f, 1.f - fabsf((static_cast<float>(k + first) - centre + 0.5f) * inv));
    };
    float wxs = 0.f, wys = 0.f;
    for (int k = 0; k < nx; ++k) wxs += tri(k, x0, cx, ix);
    for (int k = 0; k < ny; ++k) wys += tri(k, y0, cy, iy);
    const float nxs = wxs != 0.f ? 1.f / wxs : 0.f, nys = wys != 0.f ? 1.f / wys : 0.f;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int ky = 0; ky < ny; ++ky) {
        const float wy = tri(ky, y0, cy, iy) * nys;
        const uint8_t* row = p.src + (static_cast<size_t>(y0 + ky) * p.W + x0) * p.C;
        float r[4] = {0.f, 0.f, 0.f, 0.f};
        for (int kx = 0; kx < nx; ++kx) {
            const float w = tri(kx, x0, cx, ix) * nxs;
#pragma unroll
            for (int c = 0; c < 4; ++c)
                if (c < p.C) r[c] += w * static_cast<float>(__ldg(row + kx * p.C + c));
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[c] += wy * r[c];
    }
    const size_t plane = static_cast<size_t>(p.h_out) * p.w_out;
#pragma unroll
    for (int c = 0; c < 4; ++c)
        if (c < p.C) p.dst[static_cast<size_t>(c) * plane + static_cast<size_t>(i) * p.w_out + j] = acc[c] * (1.0f / 255.0f);
}

}  // namespace rdvc
