#!/usr/bin/env python
"""Profiling target: a few builds + lookups at 1080p (developer tool; run under ncu)."""
import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import torch
import rdvc_corr_b200 as rc

tile = int(sys.argv[1]) if len(sys.argv) > 1 else 0
vol = torch.bfloat16 if (len(sys.argv) > 2 and sys.argv[2] == "bf16") else torch.float32
nb = int(sys.argv[3]) if len(sys.argv) > 3 else 2
B, D, h, w = 1, 256, 136, 240
g = torch.Generator(device="cuda").manual_seed(0)
f1 = torch.randn(B, D, h, w, device="cuda", generator=g)
f2 = torch.randn(B, D, h, w, device="cuda", generator=g)
rc._cabi.load().rdvc_corr_set_option(1, tile)
for kv in os.environ.get("RDVC_OPTS", "").split(","):
    if kv:
        k, v = kv.split("=")
        assert rc._cabi.load().rdvc_corr_set_option(int(k), int(v)) == 0, kv
blk = rc.TVCorrBlock(volume_dtype=vol)
for _ in range(nb):
    blk.build_pyramid(f1, f2)
for i in range(3):
    ys, xs = torch.meshgrid(torch.arange(h, device="cuda"), torch.arange(w, device="cuda"), indexing="ij")
    co = torch.stack([xs, ys]).float()[None] + (1.0 + i) * torch.randn(B, 2, h, w, device="cuda", generator=g)
    blk.index_pyramid(co)
torch.cuda.synchronize()
print("done")
