#!/usr/bin/env python
"""Lookup timing at 1080p for the library RDVC_CORR_LIB points at (developer tool): 12 drifting lookups back to back,
CUDA events, median of 20 pairs -- to compare builds with different RDVC_LKP_LD load hints."""
import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import torch
import rdvc_corr_b200 as rc
dev = torch.device("cuda", 0)
B, D, h, w = 1, 256, 136, 240
g = torch.Generator(device=dev).manual_seed(0)
f1 = torch.randn(B, D, h, w, device=dev, generator=g); f2 = torch.randn(B, D, h, w, device=dev, generator=g)
ys, xs = torch.meshgrid(torch.arange(h, device=dev), torch.arange(w, device=dev), indexing="ij")
c = torch.stack([xs, ys], 0).float()[None]
coords = []
for _ in range(12):
    c = c + 0.5 * torch.randn(c.shape, device=dev, generator=g); coords.append(c.clone())
for vol in (torch.float32, torch.bfloat16):
    blk = rc.TVCorrBlock(volume_dtype=vol); blk.build_pyramid(f1, f2)
    out = torch.empty(B, 324, h, w, device=dev)
    ts = []
    for rep in range(22):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(12): rc.index_pyramid(blk._pyr, coords[k], 4, out=out)
        e1.record(); torch.cuda.synchronize()
        if rep >= 2: ts.append(e0.elapsed_time(e1) / 12 * 1e3)
    ts.sort()
    print(os.path.basename(rc._cabi.lib_path()), str(vol).split(".")[1], "lookup us/launch median %.2f min %.2f" % (ts[len(ts) // 2], ts[0]), flush=True)
    blk.release()
