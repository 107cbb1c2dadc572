#!/usr/bin/env python
"""CTA-pair build kernel (option key 12 = 2) vs the single-CTA kernel: equality on small shapes, timing at 1080p."""
import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import torch
import rdvc_corr_b200 as rc
lib = rc._cabi.load()
dev = torch.device("cuda", 0)

def build(shape, vol, pair):
    B, D, h, w = shape
    g = torch.Generator(device=dev).manual_seed(1)
    f1 = torch.randn(B, D, h, w, device=dev, generator=g); f2 = torch.randn(B, D, h, w, device=dev, generator=g)
    assert lib.rdvc_corr_set_option(12, 2 if pair else 1) == 0
    pyr = rc.build_pyramid(f1, f2, 4, vol)
    torch.cuda.synchronize()
    return pyr

n0 = lib.rdvc_corr_launch_count()
for shape in [(1, 64, 16, 16), (2, 64, 18, 22), (1, 128, 33, 47), (1, 256, 46, 80), (3, 64, 24, 40), (1, 256, 90, 160)]:
    for vol in (torch.float32, torch.bfloat16):
        a = build(shape, vol, False); b = build(shape, vol, True)
        same = torch.equal(a.buffer, b.buffer)
        diff = max((a.level(l).float() - b.level(l).float()).abs().max().item() for l in range(4))
        print(shape, str(vol).split(".")[1], "identical" if same else f"DIFFERENT max abs {diff}", flush=True)

B, D, h, w = 1, 256, 136, 240
g = torch.Generator(device=dev).manual_seed(0)
f1 = torch.randn(B, D, h, w, device=dev, generator=g); f2 = torch.randn(B, D, h, w, device=dev, generator=g)
for vol in (torch.float32, torch.bfloat16):
    for pair in (1, 2):
        lib.rdvc_corr_set_option(12, pair)
        blk = rc.TVCorrBlock(volume_dtype=vol)
        for _ in range(2): blk.build_pyramid(f1, f2)
        torch.cuda.synchronize()
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record(); k1.record(); torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            lib.rdvc_corr_set_profile_events(k0.cuda_event, k1.cuda_event)
            blk.build_pyramid(f1, f2); torch.cuda.synchronize()
            ts.append(k0.elapsed_time(k1))
        lib.rdvc_corr_set_profile_events(None, None)
        print(f"1080p {str(vol).split('.')[1]:8s} {'pair  ' if pair == 2 else 'single'}: build kernel {sorted(ts)[2]:.3f} ms", flush=True)
        blk.release()
lib.rdvc_corr_set_option(12, 0)
