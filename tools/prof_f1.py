#!/usr/bin/env python
"""ncu target for row f-1 (developer tool): one 1080p pyramid, then lookup (K-major bf16 rows) + 1x1 GEMM, three times.
    ncu --set full --clock-control none --import-source on -k regex:"corr_conv1x1|corr_lookup_tiled" -c 6 -o out python tools/prof_f1.py"""
import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import torch
import rdvc_corr_b200 as rc

dev = torch.device("cuda", 0)
B, D, h, w = 1, 256, 136, 240
vol = torch.bfloat16 if "bf16" in sys.argv else torch.float32
g = torch.Generator(device=dev).manual_seed(0)
f1 = torch.randn(B, D, h, w, device=dev, generator=g)
f2 = torch.randn(B, D, h, w, device=dev, generator=g)
ys, xs = torch.meshgrid(torch.arange(h, device=dev), torch.arange(w, device=dev), indexing="ij")
weight = torch.randn(256, 324, 1, 1, device=dev, generator=g) * 0.05
bias = torch.randn(256, device=dev, generator=g)
blk = rc.TVCorrBlock(volume_dtype=vol)
blk.build_pyramid(f1, f2)
for k in range(3):
    co = torch.stack([xs, ys], 0).float()[None] + 0.5 * (k + 1) * torch.randn(1, 2, h, w, device=dev, generator=g)
    out = blk.index_pyramid_convcorr1(co, weight, bias)
torch.cuda.synchronize()
print("ok", float(out.abs().mean()))
