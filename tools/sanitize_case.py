#!/usr/bin/env python
"""Small end-to-end case for compute-sanitizer (developer tool): odd shapes, both build modes, both
volume dtypes, all lookup variants."""
import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import numpy as np, torch
import rdvc_corr_b200 as rc
lib = rc._cabi.load()
for (B, D, h, w) in [(2, 64, 18, 22), (1, 128, 33, 47), (1, 64, 16, 16)]:
    rng = np.random.default_rng(1)
    f1, f2 = (rng.uniform(-2, 2, (B, D, h, w)).astype(np.float32) for _ in range(2))
    ys, xs = np.meshgrid(np.arange(h, dtype=np.float32), np.arange(w, dtype=np.float32), indexing="ij")
    grid = np.broadcast_to(np.stack([xs, ys])[None], (B, 2, h, w))
    for mode in (1, 2):
        for vol in (torch.float32, torch.bfloat16):
            for tma in (0, 1):
                lib.rdvc_corr_set_option(4, mode); lib.rdvc_corr_set_option(5, tma)
                blk = rc.TVCorrBlock(volume_dtype=vol)
                blk.build_pyramid(torch.from_numpy(f1).cuda(), torch.from_numpy(f2).cuda())
                for sigma in (0.0, 3.0, 60.0):
                    for variant in (1, 2):
                        lib.rdvc_corr_set_option(0, variant)
                        co = torch.from_numpy((grid + rng.uniform(-sigma, sigma, grid.shape)).astype(np.float32)).cuda()
                        out = blk.index_pyramid(co)
                torch.cuda.synchronize()
print("sanitize case done", float(out.abs().sum()))
