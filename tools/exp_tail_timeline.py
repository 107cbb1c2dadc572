#!/usr/bin/env python
"""Where the encoder-tail kernel's time goes (developer tool, EXPERIMENTS library only): globaltimer stamps per CTA.
    RDVC_CORR_LIB=<pkg>/lib/librdvc_corr_exp.so python tools/exp_tail_timeline.py"""
import ctypes, os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import torch
import rdvc_corr_b200 as rc
assert rc._cabi.has_experiments(), "point RDVC_CORR_LIB at librdvc_corr_exp.so"
lib = rc._cabi.load()
lib.rdvc_exp_conv1x1_timeline.argtypes = [ctypes.c_void_p]
dev = torch.device("cuda", 0)
B, h, w = 1, 136, 240
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(2 * B, 128, h, w, device=dev, generator=g).relu()
conv = torch.nn.Conv2d(128, 256, 1).to(dev)
blk = rc.TVCorrBlock()
x1, x2 = torch.chunk(x, 2, 0)
stamps = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
names = ["entry", "setup done", "W landed", "A t0", "A t1", "A t2", "A t3", "A t4", "acc t0", "stores t0", "acc t1", "stores t1",
         "acc t2", "stores t2", "acc t3", "stores t3"]
with torch.no_grad():
    for rep in range(3):
        stamps.zero_()
        lib.rdvc_exp_conv1x1_timeline(stamps.data_ptr())
        blk.build_pyramid_from_encoder(x1, x2, conv.weight, conv.bias)
        torch.cuda.synchronize()
        lib.rdvc_exp_conv1x1_timeline(None)
        t = stamps.view(148, 16).cpu()
        t0 = t[:, 0][t[:, 0] > 0].min().item()
        print(f"--- rep {rep}: ns after the first CTA's entry: CTA 0 | CTA 77 | median | max")
        for i, n in enumerate(names):
            col = t[:, i]; ok = col > 0
            if ok.any():
                v = (col[ok] - t0).float()
                print(f"{n:12s} {(t[0, i].item() - t0) if t[0, i] > 0 else -1:8d} {(t[77, i].item() - t0) if t[77, i] > 0 else -1:8d} {int(v.median()):8d} {int(v.max()):8d}")
