#!/usr/bin/env python
"""Lookup-kernel timing at 1080p (developer tool): drifting coords, layouts x tile shapes x volume dtypes."""
import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import torch
import rdvc_corr_b200 as rc
lib = rc._cabi.load()
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


cases = [("tiled", 0, 0, 1)]
for vol in (torch.float32, torch.bfloat16):
    ref = None
    for (name, twl, thl, split) in cases:
        if vol == torch.bfloat16 and twl in (2,):
            continue
        lib.rdvc_corr_set_option(7, twl); lib.rdvc_corr_set_option(8, thl)
        layout = rc.RDVC_LAYOUT_ROWMAJOR if name == "row" else rc.RDVC_LAYOUT_TILED
        blk = rc.TVCorrBlock(volume_dtype=vol, layout=layout); blk.build_pyramid(f1, f2)
        torch.cuda.synchronize()
        res = []
        for variant in (0, 3, 4):
            lib.rdvc_corr_set_option(0, variant)
            res.append(time_lookups(blk))
        lib.rdvc_corr_set_option(0, 0)
        rc.index_pyramid(blk._pyr, coords[11], 4, out=out)
        if ref is None: ref = out.clone()
        same = torch.equal(out, ref)
        tw, th = rc.corr_block.tile_shape(vol)
        print(f"{str(vol):15s} {name:6s} tile {tw}x{th} split {split}: {res[0]:6.1f} us/lookup  (no loads {res[1]:5.1f}, no stores {res[2]:5.1f})  same_as_row={same}", flush=True)
        blk.release()
lib.rdvc_corr_set_option(7, 0); lib.rdvc_corr_set_option(8, 0)
