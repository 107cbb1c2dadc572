"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the two steps between RAFT and the motion
compensation network in RDVC's P-frame path ("next" row f-4 of SURVEY.md section 8):

* ``resize_flow``    R:codec_processing.py:772-818  (bilinear resize, antialias=False, then
                     dx *= W_out / W_in, dy *= H_out / H_in; called at :1446 and :1471)
* ``WarpingLayer``   R:codec_processing.py:322-367  (grid_sample bilinear / border / align_corners=True
                     of the previous frame at pixel + flow; called at :1456)

Plain numpy, absolute pixel coordinates (no [-1, 1] round trip), float64 arithmetic.  PINNED: the
fixtures in tests/golden/motion_warp.npz are outputs of the REFERENCE'S OWN definitions, cut out of
R:codec_processing.py by tests/golden/make_golden_warp.py and executed in the build container;
tests/test_oracle.py checks this restatement against them.
"""
from __future__ import annotations

import numpy as np


def _src_index(n_out: int, n_in: int):
    """aten upsample_bilinear2d, align_corners=False, no explicit scale (what TF.resize passes):
    src = (dst + 0.5) * n_in / n_out - 0.5, clamped below at 0; i1 = min(i0 + 1, n_in - 1)."""
    scale = np.float32(n_in) / np.float32(n_out)                # aten computes the ratio in fp32
    src = (np.arange(n_out, dtype=np.float32) + np.float32(0.5)) * scale - np.float32(0.5)
    src = np.maximum(src, np.float32(0.0))
    i0 = np.minimum(np.floor(src).astype(np.int64), n_in - 1)
    i1 = np.minimum(i0 + 1, n_in - 1)
    lam = (src - i0.astype(np.float32)).astype(np.float64)
    return i0, i1, lam


def resize_flow(flow: np.ndarray, target_hw) -> np.ndarray:
    """R:codec_processing.py:772-818."""
    B, C, h_in, w_in = flow.shape
    if C != 2:
        raise ValueError(f"Flow tensor must have 2 channels, got {C}")          # :784
    H, W = target_hw
    if (h_in, w_in) == (H, W):
        return flow                                                             # :788-789
    if h_in == 0 or w_in == 0 or H == 0 or W == 0:
        return np.zeros((B, C, H, W), flow.dtype)                               # :792-797
    y0, y1, ly = _src_index(H, h_in)
    x0, x1, lx = _src_index(W, w_in)
    f = flow.astype(np.float64)
    top = f[:, :, y0][:, :, :, x0] * (1 - lx) + f[:, :, y0][:, :, :, x1] * lx
    bot = f[:, :, y1][:, :, :, x0] * (1 - lx) + f[:, :, y1][:, :, :, x1] * lx
    out = top * (1 - ly)[None, None, :, None] + bot * ly[None, None, :, None]
    out[:, 0] *= float(W) / w_in                                                # :808-813
    out[:, 1] *= float(H) / h_in
    return out.astype(np.float32)


def warp(x: np.ndarray, flow: np.ndarray) -> np.ndarray:
    """R:codec_processing.py:326-367: out[b,c,i,j] = bilinear(x[b,c]; i + dy, j + dx), sample
    coordinates clamped to the image (padding_mode='border'), pixel centres at integers
    (align_corners=True)."""
    B, C, H, W = x.shape
    if flow.shape[-2:] != (H, W) or flow.shape[1] != 2:
        raise ValueError(f"Input image ({B},{C},{H},{W}) and flow ({flow.shape}) shape/channel mismatch.")
    ii, jj = np.meshgrid(np.arange(H, dtype=np.float64), np.arange(W, dtype=np.float64), indexing="ij")
    sx = np.clip(jj[None] + flow[:, 0].astype(np.float64), 0, W - 1)
    sy = np.clip(ii[None] + flow[:, 1].astype(np.float64), 0, H - 1)
    x0 = np.floor(sx).astype(np.int64); y0 = np.floor(sy).astype(np.int64)
    x1 = np.minimum(x0 + 1, W - 1); y1 = np.minimum(y0 + 1, H - 1)
    fx = (sx - x0)[:, None]; fy = (sy - y0)[:, None]
    xd = x.astype(np.float64)
    b = np.arange(B)[:, None, None, None]; c = np.arange(C)[None, :, None, None]
    g = lambda yy, xx: xd[b, c, yy[:, None], xx[:, None]]
    out = (g(y0, x0) * (1 - fx) + g(y0, x1) * fx) * (1 - fy) + (g(y1, x0) * (1 - fx) + g(y1, x1) * fx) * fy
    return out.astype(np.float32)


def motion_warp(prev: np.ndarray, flow_raft: np.ndarray, frame_hw):
    """Steps 3 + 5a of the P-frame path (R:codec_processing.py:1446,1456): (warped_prev, flow at frame
    resolution)."""
    f = resize_flow(flow_raft, frame_hw)
    return warp(prev, f), f


def synth_case(B, C, H, W, h_in, w_in, sigma, seed):
    """Deterministic inputs: a smooth-ish image in [0,1] and a flow field of scale `sigma` px (RAFT res)."""
    rng = np.random.default_rng(seed)
    img = rng.random((B, C, H, W), dtype=np.float32)
    yy, xx = np.meshgrid(np.linspace(0, 3.1, h_in), np.linspace(0, 4.7, w_in), indexing="ij")
    base = np.stack([np.sin(yy + 0.3 * xx), np.cos(0.7 * yy - xx)], 0)[None].astype(np.float32)
    flow = (sigma * (base + 0.25 * rng.standard_normal((B, 2, h_in, w_in)))).astype(np.float32)
    return img, flow
