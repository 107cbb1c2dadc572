#!/usr/bin/env python
"""ncu target for row f-2's encoder tail (developer tool): pack128 -> encoder-tail GEMM -> packed build at 1080p, three times.
    ncu --set full --clock-control none --import-source on -k regex:"corr_encoder_tail" -c 2 -o out python tools/prof_tail.py"""
import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import torch
import rdvc_corr_b200 as rc

dev = torch.device("cuda", 0)
B, Din, D, h, w = 1, 128, 256, 136, 240
g = torch.Generator(device=dev).manual_seed(0)
x1 = torch.randn(B, Din, h, w, device=dev, generator=g)
x2 = torch.randn(B, Din, h, w, device=dev, generator=g)
weight = torch.randn(D, Din, 1, 1, device=dev, generator=g) / Din ** 0.5
bias = torch.randn(D, device=dev, generator=g) * 0.1
blk = rc.TVCorrBlock()
with torch.no_grad():
    for _ in range(3):
        blk.build_pyramid_from_encoder(x1, x2, weight, bias)
torch.cuda.synchronize()
print("ok", blk._pyr.B)
