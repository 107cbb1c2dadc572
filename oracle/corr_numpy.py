"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the RAFT correlation path.

Follows torchvision 0.26.0 ``models/optical_flow`` (prefix TV:), which is the
code RDVC's encoder actually runs (R:codec_processing.py:48-53, :1442):

* :func:`corr_volume`     TV:raft.py:424-431
* :func:`build_pyramid`   TV:raft.py:360-392
* :func:`index_pyramid`   TV:raft.py:394-422 + TV:_utils.py:8-19 + aten
  ``grid_sampler_2d`` (bilinear, zeros padding, align_corners=True)
* :func:`make_coords_grid` TV:_utils.py:22-26

It samples at ABSOLUTE coordinates (no normalise/un-normalise round trip), so
it differs from aten by ~1e-6 relative -- that is deliberate: it is the
second, independent statement the C oracle and torchvision are checked
against (tests/test_oracle.py).
"""
from __future__ import annotations

import numpy as np


def level_dims(h: int, w: int, num_levels: int):
    """Spatial size of every pyramid level; floor halving (TV:raft.py:390-392)."""
    dims = []
    for _ in range(num_levels):
        dims.append((h, w))
        h, w = h // 2, w // 2
    return dims


def check_shapes(shape1, shape2, num_levels: int):
    """The two ValueErrors of TV:raft.py:368-383."""
    if tuple(shape1) != tuple(shape2):
        raise ValueError(
            f"Input feature maps should have the same shape, instead got {tuple(shape1)} "
            f"(fmap1.shape) != {tuple(shape2)} (fmap2.shape)"
        )
    min_fmap_size = 2 * (2 ** (num_levels - 1))
    if any(s < min_fmap_size for s in shape1[-2:]):
        raise ValueError(
            "Feature maps are too small to be down-sampled by the correlation pyramid. "
            f"H and W of feature maps should be at least {min_fmap_size}; got: {tuple(shape1[-2:])}."
        )


def corr_volume(fmap1: np.ndarray, fmap2: np.ndarray, dtype=np.float64) -> np.ndarray:
    """(B,C,h,w) x2 -> (B*h*w, h, w) = fmap1^T . fmap2 / sqrt(C)."""
    B, C, h, w = fmap1.shape
    a = fmap1.reshape(B, C, h * w).astype(dtype)
    b = fmap2.reshape(B, C, h * w).astype(dtype)
    vol = np.einsum("bci,bcj->bij", a, b) / np.sqrt(dtype(C))
    return vol.reshape(B * h * w, h, w)


def avg_pool2(x: np.ndarray) -> np.ndarray:
    """2x2 stride-2 mean over the last two dims, odd trailing row/col dropped."""
    q, h, w = x.shape
    ho, wo = h // 2, w // 2
    x = x[:, : 2 * ho, : 2 * wo].reshape(q, ho, 2, wo, 2)
    return x.sum(axis=(2, 4)) / x.dtype.type(4)


def build_pyramid(fmap1, fmap2, num_levels: int = 4, dtype=np.float64):
    check_shapes(fmap1.shape, fmap2.shape, num_levels)
    levels = [corr_volume(fmap1, fmap2, dtype)]
    for _ in range(num_levels - 1):
        levels.append(avg_pool2(levels[-1]))
    return levels


def make_coords_grid(B: int, h: int, w: int) -> np.ndarray:
    ys, xs = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
    g = np.stack([xs, ys], axis=0).astype(np.float32)
    return np.broadcast_to(g[None], (B, 2, h, w)).copy()


def index_pyramid(levels, coords: np.ndarray, radius: int = 4, dtype=np.float64) -> np.ndarray:
    """levels[l]: (B*N, h_l, w_l); coords: (B,2,h,w) -> (B, L*(2r+1)^2, h, w).

    Output channel = l*S*S + i*S + j with i offsetting x and j offsetting y.
    """
    B, two, h, w = coords.shape
    assert two == 2
    N = h * w
    S = 2 * radius + 1
    L = len(levels)
    out = np.zeros((B, L * S * S, N), dtype=dtype)
    cx = coords[:, 0].reshape(B * N).astype(dtype)
    cy = coords[:, 1].reshape(B * N).astype(dtype)
    d = np.arange(-radius, radius + 1, dtype=dtype)
    rows = np.arange(B * N)
    for l, lvl in enumerate(levels):
        hl, wl = lvl.shape[-2:]
        lv = lvl.astype(dtype)
        xs = cx[:, None] / (2.0 ** l) + d[None, :]  # (BN, S) over i
        ys = cy[:, None] / (2.0 ** l) + d[None, :]  # (BN, S) over j
        x0 = np.floor(xs); y0 = np.floor(ys)
        fx = xs - x0; fy = ys - y0
        x0 = x0.astype(np.int64); y0 = y0.astype(np.int64)

        def tap(yy, xx):  # yy: (BN,S_j) xx: (BN,S_i) -> (BN, S_i, S_j)
            ok = ((xx >= 0) & (xx < wl))[:, :, None] & ((yy >= 0) & (yy < hl))[:, None, :]
            xc = np.clip(xx, 0, wl - 1)[:, :, None]
            yc = np.clip(yy, 0, hl - 1)[:, None, :]
            v = lv[rows[:, None, None], yc, xc]
            return np.where(ok, v, 0.0)

        wx0 = (1.0 - fx)[:, :, None]; wx1 = fx[:, :, None]
        wy0 = (1.0 - fy)[:, None, :]; wy1 = fy[:, None, :]
        val = (tap(y0, x0) * wx0 * wy0 + tap(y0, x0 + 1) * wx1 * wy0 +
               tap(y0 + 1, x0) * wx0 * wy1 + tap(y0 + 1, x0 + 1) * wx1 * wy1)
        # val: (BN, i, j) -> channels l*S*S + i*S + j
        out[:, l * S * S:(l + 1) * S * S, :] = (
            val.reshape(B, N, S * S).transpose(0, 2, 1)
        )
    return out.reshape(B, L * S * S, h, w)


# ---------------------------------------------------------------------------
# Deterministic, RNG-free synthetic inputs (identical on every platform, so
# fixtures need to store outputs only).
# ---------------------------------------------------------------------------
def hash_uniform(n: int, seed: int) -> np.ndarray:
    """n floats in [-0.5, 0.5): a Knuth multiplicative hash of the index."""
    i = np.arange(n, dtype=np.uint64)
    x = (i * np.uint64(2654435761) + np.uint64(seed) * np.uint64(0x9E3779B9)) & np.uint64(0xFFFFFFFF)
    x ^= x >> np.uint64(15)
    x = (x * np.uint64(2246822519)) & np.uint64(0xFFFFFFFF)
    x ^= x >> np.uint64(13)
    return ((x >> np.uint64(8)).astype(np.float64) / float(1 << 24) - 0.5).astype(np.float32)


def synth_fmaps(B: int, C: int, h: int, w: int, seed: int = 0, scale: float = 2.0):
    n = B * C * h * w
    f1 = (hash_uniform(n, 2 * seed + 1) * scale).reshape(B, C, h, w)
    f2 = (hash_uniform(n, 2 * seed + 2) * scale).reshape(B, C, h, w)
    return f1, f2


def synth_coords(B: int, h: int, w: int, sigma: float, seed: int = 0) -> np.ndarray:
    """Identity grid plus a deterministic perturbation of amplitude ~sigma px."""
    g = make_coords_grid(B, h, w)
    if sigma == 0:
        return g
    pert = hash_uniform(g.size, 1000 + seed).reshape(g.shape) * np.float32(2.0 * sigma)
    return (g + pert).astype(np.float32)
