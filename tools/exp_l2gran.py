#!/usr/bin/env python
"""Experiment (developer tool): cudaLimitMaxL2FetchGranularity = 32 / 64 / 128 against pyramid tile shapes.
Outcome on B200: the limit changes nothing (28.6 us fp32 4x4 tiles, 33.8 us 4x2 tiles at every setting)."""
import os, sys, ctypes
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch
import rdvc_corr_b200 as rc
lib = rc._cabi.load()
rt = ctypes.CDLL("libcudart.so.12")
def get():
    v = ctypes.c_size_t(0); rt.cudaDeviceGetLimit(ctypes.byref(v), 5); return v.value
B, D, h, w = 1, 256, 136, 240
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
f1 = torch.randn(B, D, h, w, device=dev, generator=g); f2 = torch.randn(B, D, h, w, device=dev, generator=g)
ys, xs = torch.meshgrid(torch.arange(h, device=dev), torch.arange(w, device=dev), indexing="ij")
base = torch.stack([xs, ys], 0).float()[None]
coords = []; c = base.clone()
for _ in range(12):
    c = c + 0.5 * torch.randn(c.shape, device=dev, generator=g); coords.append(c.clone())
out = torch.empty(B, 324, h, w, device=dev)
def time_lookups(blk):
    for k in range(12): rc.index_pyramid(blk._pyr, coords[k], 4, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for rep in range(5):
        for k in range(12): rc.index_pyramid(blk._pyr, coords[k], 4, out=out)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 60 * 1000
print("default L2 fetch granularity:", get())
for gran in (64, 32, 128):
    r = rt.cudaDeviceSetLimit(5, ctypes.c_size_t(gran)); print("set", gran, "rc", r, "now", get())
    for vol in (torch.float32, torch.bfloat16):
        for (twl, thl) in ((0, 0), (2, 1), (3, 1), (3, 2), (2, 2)):
            if vol == torch.bfloat16 and twl == 2: continue
            lib.rdvc_corr_set_option(7, twl); lib.rdvc_corr_set_option(8, thl)
            blk = rc.TVCorrBlock(volume_dtype=vol); blk.build_pyramid(f1, f2); torch.cuda.synchronize()
            tw, th = rc.corr_block.tile_shape(vol)
            print(f"  gran {gran} {str(vol).split('.')[1]:8s} tile {tw}x{th}: {time_lookups(blk):6.1f} us/lookup", flush=True)
            blk.release()
lib.rdvc_corr_set_option(7, 0); lib.rdvc_corr_set_option(8, 0)
