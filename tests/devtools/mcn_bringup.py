#!/usr/bin/env python
"""Developer tool: structured single-layer probes of the MCN convolution kernel (identity, one-pixel shifts,
random weights) with an error breakdown by channel / pixel parity / row, so a layout mistake shows its shape."""
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import rdvc_corr_b200 as rc  # noqa: F401
from oracle import mcn as om
from rdvc_corr_b200 import mcn as hm


def probe(name, w, B=1, H=16, W=64, k=3, seed=0, bias=None, res=False, act=False):
    rng = np.random.default_rng(seed)
    cin = w.shape[1]
    x = np.zeros((B, 32, H, W), np.float32)
    x[:, :cin] = rng.standard_normal((B, cin, H, W)).astype(np.float16)
    r = rng.standard_normal((B, 32, H, W)).astype(np.float16).astype(np.float32) if res else None
    packed, mask = hm.pack_conv_weights(torch.from_numpy(w))
    plane = hm.plane_from_nchw(torch.from_numpy(x).cuda())
    rp = None if r is None else hm.plane_from_nchw(torch.from_numpy(r).cuda())
    out = hm.conv_layer(plane, packed.cuda(), mask, None if bias is None else torch.from_numpy(bias), k,
                        hm.ACT_LEAKY if act else hm.ACT_NONE, B, H, W, residual=rp)
    torch.cuda.synchronize()
    got = hm.plane_to_nchw(out, B, H, W).cpu().numpy()
    want = om.conv_layer(x[:, :cin], w, bias, act, residual=r, emulate_fp16=True)
    err = np.abs(got - want)
    print(f"{name}: max err {err.max():.4g} (max |ref| {np.abs(want).max():.3g}) mask={mask:#x}", flush=True)
    if err.max() > 4e-3 * max(1.0, np.abs(want).max()):
        print("   by channel   :", np.round(err.max(axis=(0, 2, 3))[:8], 3), "...")
        print("   by x parity  :", err[..., 0::2].max(), err[..., 1::2].max())
        print("   by row       :", np.round(err.max(axis=(0, 1, 3)), 3))
        print("   by x (first 16):", np.round(err.max(axis=(0, 1, 2))[:16], 3))
        print("   got[0,0,:2,:6]\n", got[0, 0, :2, :6], "\n   want\n", want[0, 0, :2, :6])
    return err.max()


def main():
    eye = np.zeros((32, 32, 3, 3), np.float32)
    eye[np.arange(32), np.arange(32), 1, 1] = 1
    probe("identity 3x3", eye)
    for (dy, dx) in [(0, 2), (0, 0), (2, 1), (1, 0)]:
        w = np.zeros((32, 32, 3, 3), np.float32)
        w[np.arange(32), np.arange(32), dy, dx] = 1
        probe(f"shift ky={dy} kx={dx}", w)
    rng = np.random.default_rng(1)
    w = (rng.standard_normal((32, 32, 3, 3)) / 17).astype(np.float32)
    probe("random 3x3", w)
    probe("random 3x3 + bias + res + act", w, bias=rng.standard_normal(32).astype(np.float32), res=True, act=True)
    probe("random 3x3 ragged", w, B=2, H=21, W=45)
    w5 = (rng.standard_normal((32, 32, 5, 5)) / 28).astype(np.float32)
    probe("random 5x5", w5, k=5, H=19, W=70)
    probe("random 3x3 many tiles", w, H=70, W=330, res=True, act=True)


if __name__ == "__main__":
    main()
