#!/usr/bin/env python
"""Golden fixtures for the post-RAFT motion-branch steps from the REFERENCE'S OWN CODE.

    python tests/golden/make_golden_warp.py          # build container only: reads /root/reference

`codec_processing.py` cannot be imported (it exits at import time without compressai / skimage,
R:codec_processing.py:26-33), so the two definitions this path needs -- `class WarpingLayer`
(R:codec_processing.py:322-367) and `def resize_flow` (:772-818) -- are cut out of the file with `ast`
and executed unmodified against the names they use (torch, F, TF_tv, transforms, traceback).  Nothing
from the reference is copied into the repo: only the outputs are stored (tests/golden/motion_warp.npz).
"""
from __future__ import annotations

import ast
import os
import sys
import traceback

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
import torchvision.transforms as transforms
import torchvision.transforms.functional as TF_tv

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
from oracle import motion_warp as mw  # noqa: E402

REF = "/root/reference/codec_processing.py"

# (name, B, C, H, W, h_in, w_in, sigma): frame size (H, W), RAFT-resolution flow (h_in, w_in)
CASES = [
    ("same_size", 1, 3, 24, 40, 24, 40, 1.5),        # resize_flow returns its input (:788)
    ("shrink_1088_to_1080_like", 1, 3, 45, 80, 48, 80, 2.0),   # the 1080p case in miniature: H only
    ("grow", 2, 3, 37, 61, 16, 24, 1.0),             # both axes, odd sizes, B > 1
    ("border", 1, 1, 20, 28, 24, 32, 25.0),          # most samples clamp to the border
    ("zero_flow", 1, 3, 18, 22, 24, 32, 0.0),        # identity warp
]


def reference_definitions():
    src = open(REF).read()
    tree = ast.parse(src)
    want = {"WarpingLayer", "resize_flow"}
    ns = {"torch": torch, "nn": nn, "F": F, "TF_tv": TF_tv, "transforms": transforms, "traceback": traceback}
    for node in tree.body:
        if isinstance(node, (ast.ClassDef, ast.FunctionDef)) and node.name in want:
            exec(compile(ast.Module([node], []), REF, "exec"), ns)
            want.discard(node.name)
    assert not want, f"not found in the reference: {want}"
    return ns["WarpingLayer"](), ns["resize_flow"]


def main():
    warp_layer, resize_flow = reference_definitions()
    out = {}
    for (name, B, C, H, W, h_in, w_in, sigma) in CASES:
        img, flow = mw.synth_case(B, C, H, W, h_in, w_in, sigma, seed=len(name))
        with torch.no_grad():
            f = resize_flow(torch.from_numpy(flow), (H, W))
            wv = warp_layer(torch.from_numpy(img), f)
        out[f"{name}_flow_frame"] = f.numpy()
        out[f"{name}_warped"] = wv.numpy()
        out[f"{name}_shape"] = np.array([B, C, H, W, h_in, w_in])
        out[f"{name}_sigma"] = np.array(sigma, np.float32)
    np.savez_compressed(os.path.join(HERE, "motion_warp.npz"), **out)
    print("wrote motion_warp.npz:", {k: v.shape for k, v in out.items() if k.endswith("_warped")})


if __name__ == "__main__":
    main()
