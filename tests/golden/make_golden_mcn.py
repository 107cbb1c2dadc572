#!/usr/bin/env python
"""Golden fixture for the motion-compensation network from the REFERENCE'S OWN CODE.

    python tests/golden/make_golden_mcn.py          # build container only: reads /root/reference

`codec_processing.py` cannot be imported (R:codec_processing.py:26-33 exits without compressai / skimage), so the
four definitions this row needs -- `get_activation` (:101-115), `ConvNormAct` (:117-156), `ResidualBlock`
(:190-217) and `MotionCompensationNetwork` (:369-406) -- are cut out of the file with `ast` and executed unmodified.
Nothing from the reference is copied into the repo: the fixture (tests/golden/mcn.npz) holds a seeded state_dict
(random weights AND non-trivial BatchNorm running statistics, so the folding is exercised), seeded inputs, and the
reference's eval-mode outputs.
"""
from __future__ import annotations

import ast
import os

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/codec_processing.py"

# (name, B, H, W, flow sigma): even / odd widths, sizes that are not multiples of the 8 x 32 pixel tile, B > 1
CASES = [
    ("tile_exact", 1, 16, 64, 2.0),
    ("ragged", 2, 21, 45, 6.0),
    ("tiny", 1, 5, 7, 1.0),
]


def reference_mcn():
    tree = ast.parse(open(REF).read())
    want = ["get_activation", "ConvNormAct", "ResidualBlock", "MotionCompensationNetwork"]
    ns = {"torch": torch, "nn": nn, "F": F}
    for node in tree.body:
        if isinstance(node, (ast.ClassDef, ast.FunctionDef)) and node.name in want:
            exec(compile(ast.Module([node], []), REF, "exec"), ns)
            want.remove(node.name)
    assert not want, f"not found in the reference: {want}"
    return ns["MotionCompensationNetwork"]


def seeded_network(cls, seed: int = 0):
    torch.manual_seed(seed)
    net = cls()                                   # the reference's defaults: 8 -> 32, 3 residual blocks, -> 3
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for m in net.modules():
            if isinstance(m, nn.BatchNorm2d):     # a trained network's statistics are not (0, 1)
                m.running_mean.copy_(torch.randn(m.num_features, generator=g) * 0.3)
                m.running_var.copy_(torch.rand(m.num_features, generator=g) * 1.5 + 0.25)
                m.weight.copy_(torch.rand(m.num_features, generator=g) + 0.5)
                m.bias.copy_(torch.randn(m.num_features, generator=g) * 0.2)
    return net.eval()


def synth_inputs(B: int, H: int, W: int, sigma: float, seed: int):
    rng = np.random.default_rng(seed)
    warped = rng.random((B, 3, H, W), dtype=np.float32)
    ref = rng.random((B, 3, H, W), dtype=np.float32)
    flow = (rng.standard_normal((B, 2, H, W)) * sigma).astype(np.float32)
    return warped, flow, ref


def main():
    net = seeded_network(reference_mcn())
    out = {"state:" + k: v.numpy() for k, v in net.state_dict().items()}
    for i, (name, B, H, W, sigma) in enumerate(CASES):
        warped, flow, ref = synth_inputs(B, H, W, sigma, seed=100 + i)
        with torch.no_grad():
            y = net(torch.from_numpy(warped), torch.from_numpy(flow), torch.from_numpy(ref))
        out[f"{name}:warped"], out[f"{name}:flow"], out[f"{name}:ref"] = warped, flow, ref
        out[f"{name}:out"] = y.numpy()
    np.savez_compressed(os.path.join(HERE, "mcn.npz"), **out)
    print("wrote mcn.npz:", {k: v.shape for k, v in out.items() if k.endswith(":out")})


if __name__ == "__main__":
    main()
