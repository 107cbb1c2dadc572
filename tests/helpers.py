"""Shared helpers for the parity tests (test infrastructure)."""
from __future__ import annotations

import numpy as np
import torch


def rel_max(a, b) -> float:
    """max|a-b| / max|b|  -- the max-norm relative error SURVEY.md 8(c) states tolerances in."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    d = np.abs(a - b).max() if a.size else 0.0
    m = np.abs(b).max() if b.size else 1.0
    return float(d / (m if m > 0 else 1.0))


def pyramid_from_levels(rc, levels, B, h, w, volume_dtype=torch.float32, device="cuda", layout=None):
    """Pack oracle level tensors [(B*N, h_l, w_l)] into a CorrPyramid in one of the library's layouts."""
    lib = rc._cabi.load()
    vd = {torch.float32: rc.RDVC_DT_F32, torch.bfloat16: rc.RDVC_DT_BF16}[volume_dtype]
    layout = rc.RDVC_LAYOUT_TILED if layout is None else layout
    L = len(levels)
    nbytes = lib.rdvc_corr_pyramid_bytes(B, h, w, L, vd, layout)
    buf = torch.zeros(nbytes, dtype=torch.uint8, device=device)
    pyr = rc.CorrPyramid(B, h, w, L, volume_dtype, buf, layout)
    for l, lv in enumerate(levels):
        pyr.set_level(l, torch.as_tensor(np.asarray(lv), dtype=torch.float32).to(device))
    return pyr


def bf16_round(x: np.ndarray) -> np.ndarray:
    return torch.from_numpy(np.ascontiguousarray(x)).to(torch.bfloat16).to(torch.float32).numpy()


def pool_fmap(f: np.ndarray, level: int) -> np.ndarray:
    """(B,C,h,w) -> mean over 2^level x 2^level blocks, floor-cropped (fp64 math)."""
    B, C, h, w = f.shape
    k = 1 << level
    hl, wl = h >> level, w >> level
    x = f[:, :, : hl * k, : wl * k].astype(np.float64).reshape(B, C, hl, k, wl, k)
    return x.mean(axis=(3, 5))


def fp16_round(x: np.ndarray) -> np.ndarray:
    return np.asarray(x, np.float32).astype(np.float16).astype(np.float32)


def ref_pyramid_linear(f1: np.ndarray, f2: np.ndarray, num_levels: int, round_fn=None):
    """What the library's linear build mode computes, in fp64: level l =
    r(fmap1)^T . r(avgpool_l(fmap2)) / sqrt(C)  (pooled in full precision, rounded once); r = bf16
    rounding, or fp16 rounding (`round_fn=fp16_round`) for fp16 inputs."""
    round_fn = round_fn or bf16_round
    B, C, h, w = f1.shape
    a = round_fn(f1).reshape(B, C, h * w).astype(np.float64)
    out = []
    for l in range(num_levels):
        pl = pool_fmap(f2, l)
        b = round_fn(pl.astype(np.float32)).astype(np.float64)
        hl, wl = pl.shape[-2:]
        v = np.einsum("bci,bcj->bij", a, b.reshape(B, C, hl * wl)) / np.sqrt(float(C))
        out.append(v.reshape(B * h * w, hl, wl))
    return out
