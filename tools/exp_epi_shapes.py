#!/usr/bin/env python
"""Build-kernel time at 1080p for the library RDVC_CORR_LIB points at, per volume type and epilogue shape (option key 9:
4 or 8 epilogue warps) -- to compare builds with other RDVC_EW{4,8}_{A_STAGES,STG_BUFS} settings (developer tool)."""
import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import torch
import rdvc_corr_b200 as rc
lib = rc._cabi.load()
dev = torch.device("cuda", 0)
B, D, h, w = 1, 256, 136, 240
g = torch.Generator(device=dev).manual_seed(0)
f1 = torch.randn(B, D, h, w, device=dev, generator=g); f2 = torch.randn(B, D, h, w, device=dev, generator=g)
for vol in (torch.float32, torch.bfloat16):
    for ew in (4, 8):
        lib.rdvc_corr_set_option(9, ew)
        blk = rc.TVCorrBlock(volume_dtype=vol)
        for _ in range(2): blk.build_pyramid(f1, f2)
        torch.cuda.synchronize()
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record(); k1.record(); torch.cuda.synchronize()
        ts = []
        for _ in range(7):
            lib.rdvc_corr_set_profile_events(k0.cuda_event, k1.cuda_event)
            blk.build_pyramid(f1, f2); torch.cuda.synchronize()
            ts.append(k0.elapsed_time(k1))
        lib.rdvc_corr_set_profile_events(None, None)
        print(f"{os.path.basename(rc._cabi.lib_path()):24s} {str(vol).split('.')[1]:8s} EW={ew}: build kernel {sorted(ts)[3]:.4f} ms", flush=True)
        blk.release()
lib.rdvc_corr_set_option(9, 0)
