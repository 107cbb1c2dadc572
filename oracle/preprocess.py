"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's frame preparation:

* ``preprocess_frame_raft``   R:codec_processing.py:751-761  TF.to_tensor + TF.resize(antialias=True)
* ``preprocess_frame_codec``  R:codec_processing.py:763-769  TF.to_tensor

numpy, fp64.  The resize is aten's ``_upsample_bilinear2d_aa`` (align_corners=False, no explicit scale), written
out: per output index i along an axis of n_in -> n_out samples,
    scale = n_in / n_out, support = max(scale, 1), centre = scale * (i + 0.5),
    first = max(int(centre - support + 0.5), 0), count = min(int(centre + support + 0.5), n_in) - first,
    w_j = max(0, 1 - |(j + first - centre + 0.5) / max(scale, 1)|) / sum_j(...)
applied separably (columns, then rows).  PINNED: tests/golden/preprocess.npz holds outputs of the reference's
OWN two functions, cut out of R:codec_processing.py and executed by tests/golden/make_golden_preprocess.py.
"""
from __future__ import annotations

import numpy as np


def _aa_matrix(n_in: int, n_out: int) -> np.ndarray:
    """(n_out, n_in) weights of one separable pass."""
    scale = np.float32(n_in) / np.float32(n_out)              # aten computes the ratio in fp32
    support = max(float(scale), 1.0)
    inv = 1.0 / float(scale) if scale >= 1.0 else 1.0
    m = np.zeros((n_out, n_in), np.float64)
    for i in range(n_out):
        centre = float(scale) * (i + 0.5)
        first = max(int(centre - support + 0.5), 0)
        count = min(int(centre + support + 0.5), n_in) - first
        j = np.arange(count)
        w = np.maximum(0.0, 1.0 - np.abs((j + first - centre + 0.5) * inv))
        m[i, first:first + count] = w / w.sum()
    return m


def to_tensor(frame_u8: np.ndarray) -> np.ndarray:
    """TF.to_tensor: (H, W, C) uint8 -> (C, H, W) float in [0, 1]."""
    f = frame_u8 if frame_u8.ndim == 3 else frame_u8[:, :, None]
    return np.transpose(f, (2, 0, 1)).astype(np.float64) / 255.0


def preprocess_frame_raft(frame_u8: np.ndarray, resize_shape_hw) -> np.ndarray:
    """(1, C, h, w) float32 (R:codec_processing.py:751-759)."""
    t = to_tensor(frame_u8)
    C, H, W = t.shape
    h, w = resize_shape_hw
    out = t
    if (h, w) != (H, W):
        out = np.einsum("chw,jw->chj", out, _aa_matrix(W, w))
        out = np.einsum("chj,ih->cij", out, _aa_matrix(H, h))
    return out[None].astype(np.float32)


def preprocess_frame_codec(frame_u8: np.ndarray) -> np.ndarray:
    return to_tensor(frame_u8)[None].astype(np.float32)


def synth_frame(H: int, W: int, C: int, seed: int) -> np.ndarray:
    """A deterministic uint8 frame with both smooth and noisy content."""
    rng = np.random.default_rng(seed)
    yy, xx = np.meshgrid(np.linspace(0, 6.0, H), np.linspace(0, 9.0, W), indexing="ij")
    base = 127.5 + 100.0 * np.sin(yy[..., None] + np.arange(C)) * np.cos(xx[..., None])
    return np.clip(base + rng.integers(-20, 21, (H, W, C)), 0, 255).astype(np.uint8)
