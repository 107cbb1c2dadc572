#!/usr/bin/env python
"""Where the 1x1 GEMM kernel's time goes (developer tool, EXPERIMENTS library only): globaltimer stamps per CTA.
    python -c "import rdvc_corr_b200 as rc; rc._build.build(experiments=True)"
    RDVC_CORR_LIB=<pkg>/lib/librdvc_corr_exp.so python tools/exp_conv1x1_timeline.py"""
import ctypes, os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import torch
import rdvc_corr_b200 as rc

assert rc._cabi.has_experiments(), "point RDVC_CORR_LIB at librdvc_corr_exp.so"
lib = rc._cabi.load()
dev = torch.device("cuda", 0)
B, D, h, w = 1, 256, 136, 240
g = torch.Generator(device=dev).manual_seed(0)
f1 = torch.randn(B, D, h, w, device=dev, generator=g); f2 = torch.randn(B, D, h, w, device=dev, generator=g)
ys, xs = torch.meshgrid(torch.arange(h, device=dev), torch.arange(w, device=dev), indexing="ij")
weight = torch.randn(256, 324, 1, 1, device=dev, generator=g) * 0.05
bias = torch.randn(256, device=dev, generator=g)
blk = rc.TVCorrBlock(); blk.build_pyramid(f1, f2)
co = torch.stack([xs, ys], 0).float()[None] + torch.randn(1, 2, h, w, device=dev, generator=g)
stamps = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
names = ["entry", "setup done", "W landed", "A first t0", "A first t1", "A last t0", "A last t1", "acc t0", "acc t1",
         "stores t0", "stores t1", "cta done", "cluster done", "tmem 1st t0", "tmem 1st t1"]
lib.rdvc_exp_conv1x1_flags.argtypes = [ctypes.c_int]
for rep, flags in enumerate([0, 0, 1, 2, 3]):
    lib.rdvc_exp_conv1x1_flags(flags)
    stamps.zero_()
    lib.rdvc_exp_conv1x1_timeline.argtypes = [ctypes.c_void_p]
    lib.rdvc_exp_conv1x1_timeline(stamps.data_ptr())
    blk.index_pyramid_convcorr1(co, weight, bias)
    torch.cuda.synchronize()
    lib.rdvc_exp_conv1x1_timeline(None)
    t = stamps.view(148, 16).cpu()
    t0 = t[:146, 0].min().item()
    print(f"--- rep {rep} flags {flags} (1 = no stores, 2 = no TMEM reads): ns after the first CTA's entry: CTA 0 (leader) | CTA 1 (peer) | median over CTAs | max")
    for i, n in enumerate(names):
        col = t[:146, i]
        ok = col > 0
        if ok.any():
            v = (col[ok] - t0).float()
            c0 = (t[0, i].item() - t0) if t[0, i] > 0 else -1
            c1 = (t[1, i].item() - t0) if t[1, i] > 0 else -1
            print(f"{n:14s} {c0:8d} {c1:8d} {int(v.median()):8d} {int(v.max()):8d}")
