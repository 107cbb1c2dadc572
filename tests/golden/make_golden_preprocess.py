#!/usr/bin/env python
"""Golden fixtures for the frame preparation from the REFERENCE'S OWN CODE.

    python tests/golden/make_golden_preprocess.py       # build container only: reads /root/reference

`def preprocess_frame_raft` (R:codec_processing.py:751-761) and `def preprocess_frame_codec` (:763-769) are cut
out of the reference file with `ast` and executed unmodified on CPU against the names they use (TF_tv).  Only the
outputs are stored (tests/golden/preprocess.npz).
"""
from __future__ import annotations

import ast
import os
import sys

import numpy as np
import torch
import torchvision.transforms.functional as TF_tv

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
from oracle import preprocess as pp  # noqa: E402

REF = "/root/reference/codec_processing.py"

# (name, H, W, C, out_h, out_w): frame size -> RAFT input size
CASES = [
    ("up_1080_to_1088_like", 135, 240, 3, 136, 240),    # the 1080p case in miniature: slight up-scaling in H
    ("down_3x", 138, 240, 3, 46, 80),                   # 1080p -> the reference's default 368x640, in miniature
    ("down_odd", 101, 67, 3, 40, 31),                   # non-integer factors, odd sizes
    ("same", 48, 64, 3, 48, 64),                        # resize to the same size
    ("gray_up", 20, 30, 1, 33, 47),                     # single channel, up-scaling both ways
]


def reference_functions():
    tree = ast.parse(open(REF).read())
    want = {"preprocess_frame_raft", "preprocess_frame_codec"}
    ns = {"torch": torch, "TF_tv": TF_tv}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in want:
            exec(compile(ast.Module([node], []), REF, "exec"), ns)
            want.discard(node.name)
    assert not want, want
    return ns["preprocess_frame_raft"], ns["preprocess_frame_codec"]


def main():
    raft_fn, codec_fn = reference_functions()
    out = {}
    for (name, H, W, C, h, w) in CASES:
        frame = pp.synth_frame(H, W, C, seed=len(name))
        fr = frame if C > 1 else frame[:, :, 0]
        out[f"{name}_shape"] = np.array([H, W, C, h, w])
        out[f"{name}_raft"] = raft_fn(fr, (h, w), torch.device("cpu")).numpy()
        out[f"{name}_codec"] = codec_fn(fr, torch.device("cpu")).numpy()
    np.savez_compressed(os.path.join(HERE, "preprocess.npz"), **out)
    print({k: v.shape for k, v in out.items() if k.endswith("_raft")})


if __name__ == "__main__":
    main()
