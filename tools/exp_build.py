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
    """(ms per build call, ms of the main build kernel alone) -- the difference is the pack pre-pass."""
    blk = rc.TVCorrBlock(volume_dtype=vol)
    for _ in range(2):
        blk.build_pyramid(f1, f2)
    torch.cuda.synchronize()
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0.record(); k1.record()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        if i == n - 1:
            lib.rdvc_corr_set_profile_events(k0.cuda_event, k1.cuda_event)
        blk.build_pyramid(f1, f2)
    e1.record(); torch.cuda.synchronize()
    lib.rdvc_corr_set_profile_events(None, None)
    blk.release()
    return e0.elapsed_time(e1) / n, k0.elapsed_time(k1)

cfgs = [tuple(int(x) for x in a.split(",")) for a in sys.argv[1:] if "=" not in a] or [(2, 0, 0, 15)]
for a in sys.argv[1:]:
    if "=" in a:
        k, v = a.split("=")
        assert lib.rdvc_corr_set_option(int(k), int(v)) == 0, a
vols = (torch.float32, torch.bfloat16) if os.environ.get("BOTH") else (torch.float32,)
for vol in vols:
    for (mode, tile, msplit, mask) in cfgs:
        lib.rdvc_corr_set_option(4, mode); lib.rdvc_corr_set_option(1, tile)
        lib.rdvc_corr_set_option(2, msplit); lib.rdvc_corr_set_option(3, mask)
        t, k = timeit(vol)
        print(f"{str(vol):15s} mode={mode} tile={tile} msplit={msplit} mask={mask:2d}: {t:.3f} ms/build call, "
              f"build kernel {k:.3f} ms, pre-pass {t - k:.3f} ms", flush=True)
