#!/usr/bin/env python
"""Build-kernel experiments at 1080p (developer tool): option sweeps, timing only."""
import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import torch
import rdvc_corr_b200 as rc

lib = rc._cabi.load()
B, D, h, w = 1, 256, 136, 240
g = torch.Generator(device="cuda").manual_seed(0)
f1 = torch.randn(B, D, h, w, device="cuda", generator=g)
f2 = torch.randn(B, D, h, w, device="cuda", generator=g)

def timeit(vol, n=5):
    blk = rc.TVCorrBlock(volume_dtype=vol)
    for _ in range(2):
        blk.build_pyramid(f1, f2)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        blk.build_pyramid(f1, f2)
    e1.record(); torch.cuda.synchronize()
    blk.release()
    return e0.elapsed_time(e1) / n

cfgs = [tuple(int(x) for x in a.split(",")) for a in sys.argv[1:] if "=" not in a] or [(2, 0, 0, 15)]
for a in sys.argv[1:]:
    if "=" in a:
        k, v = a.split("=")
        assert lib.rdvc_corr_set_option(int(k), int(v)) == 0, a
for vol in (torch.float32,):
    for (mode, tile, msplit, mask) in cfgs:
        lib.rdvc_corr_set_option(4, mode); lib.rdvc_corr_set_option(1, tile)
        lib.rdvc_corr_set_option(2, msplit); lib.rdvc_corr_set_option(3, mask)
        print(f"{str(vol):15s} mode={mode} tile={tile} msplit={msplit} mask={mask:2d}: {timeit(vol):.3f} ms", flush=True)
