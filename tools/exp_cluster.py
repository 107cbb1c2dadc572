#!/usr/bin/env python
"""fmap1 multicast across CTA pairs (option key 12 = 3) vs one CTA per tile (= 1): equality on small shapes, timing of the
build kernel alone (the library's profile events) at 720p / 1080p / 1440p for both volume types.  Experiments library (RDVC_CORR_LIB=.../librdvc_corr_exp.so)."""
import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import torch
import rdvc_corr_b200 as rc
lib = rc._cabi.load()
dev = torch.device("cuda", 0)


def build(shape, vol, mode):
    B, D, h, w = shape
    g = torch.Generator(device=dev).manual_seed(1)
    f1 = torch.randn(B, D, h, w, device=dev, generator=g); f2 = torch.randn(B, D, h, w, device=dev, generator=g)
    assert lib.rdvc_corr_set_option(12, mode) == 0
    pyr = rc.build_pyramid(f1, f2, 4, vol)
    torch.cuda.synchronize()
    return pyr


for shape in [(1, 64, 16, 16), (2, 64, 18, 22), (1, 128, 33, 47), (1, 256, 46, 80), (3, 64, 24, 40), (1, 256, 90, 160)]:
    for vol in (torch.float32, torch.bfloat16):
        a = build(shape, vol, 1); b = build(shape, vol, 3)
        print(shape, str(vol).split(".")[1], "identical" if torch.equal(a.buffer, b.buffer) else "DIFFERENT", flush=True)

for (h, w) in [(90, 160), (136, 240), (180, 320)]:
    B, D = 1, 256
    g = torch.Generator(device=dev).manual_seed(0)
    f1 = torch.randn(B, D, h, w, device=dev, generator=g); f2 = torch.randn(B, D, h, w, device=dev, generator=g)
    for vol in (torch.float32, torch.bfloat16):
        for mode in (1, 3):
            lib.rdvc_corr_set_option(12, mode)
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
            print(f"{h}x{w} {str(vol).split('.')[1]:8s} {'cluster' if mode == 3 else 'single '}: build kernel {sorted(ts)[3]:.4f} ms", flush=True)
            blk.release()
lib.rdvc_corr_set_option(12, 0)
