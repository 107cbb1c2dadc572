#!/usr/bin/env python
"""bench_gop.py -- BASELINE.json config 3 ("next" row): P-frames/s of the motion branch's RAFT call
on a synthetic 1080p GOP (I-frame every 10 -> 9 P-frames), 1 GPU.

Not the driver's bench (that is bench.py).  Compares, on the same B200 and the same seeded
random-init raft_large (no weights offline):
  stock   torchvision RAFT.forward with its own CorrBlock (what RDVC runs today)
  ours    rc.raft_flow with TVCorrBlock (B200 correlation block, final-only upsampling)
and reports the end-point error between the two.  Frames are a smooth random texture translated
by a known per-frame motion, generated at RAFT's input size 1088x1920, values in [0,1] like
R:codec_processing.py:751-759.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch
import torch.nn.functional as F


def make_gop(n, h, w, device, seed=0):
    g = torch.Generator(device=device).manual_seed(seed)
    base = torch.rand(1, 3, h // 16 + 8, w // 16 + 8, device=device, generator=g)
    big = F.interpolate(base, size=(h + 64, w + 64), mode="bicubic", align_corners=False).clamp(0, 1)
    frames = []
    for t in range(n):
        dx, dy = 3 * t, 2 * t          # known global motion: (3, 2) px per frame
        frames.append(big[:, :, 32 - dy:32 - dy + h, 32 - dx:32 - dx + w].contiguous())
    return frames


def timed(fn, n_warm=1):
    for _ in range(n_warm):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    return time.perf_counter() - t0, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--height", type=int, default=1088)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--gop", type=int, default=10)
    ap.add_argument("--stock-pframes", type=int, default=2, help="stock RAFT is slow at 1080p: time only this many")
    ap.add_argument("--amp", action="store_true", help="fp16 autocast like the reference's GPU default")
    args = ap.parse_args()
    import rdvc_corr_b200 as rc
    from torchvision.models.optical_flow import raft_large
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    stock = raft_large(weights=None).eval().to(dev)
    torch.manual_seed(0)
    ours = raft_large(weights=None, corr_block=rc.TVCorrBlock()).eval().to(dev)
    frames = make_gop(args.gop, args.height, args.width, dev)
    pairs = [(frames[t - 1], frames[t]) for t in range(1, args.gop)]     # open loop: originals

    ctx = lambda: torch.autocast("cuda", dtype=torch.float16, enabled=args.amp)

    def run_ours():
        with torch.no_grad(), ctx():
            return [rc.raft_flow(ours, a, b, 12) for a, b in pairs]

    def run_stock():
        with torch.no_grad(), ctx():
            return [stock(a, b, num_flow_updates=12)[-1] for a, b in pairs[:args.stock_pframes]]

    t_ours, f_ours = timed(run_ours)
    t_stock, f_stock = timed(run_stock, n_warm=1)
    epe = torch.stack([(x.float() - y.float()).pow(2).sum(1).sqrt().mean() for x, y in zip(f_ours, f_stock)])
    line = {
        "metric": "raft_motion_branch_p_frames_per_s_1080p", "unit": "P-frames/s", "n_gpus": 1,
        "config": {"workload": f"synthetic GOP of {args.gop} frames {args.width}x{args.height}, 12 RAFT updates, "
                               "seed-0 random-init raft_large", "amp_fp16": args.amp},
        "ours": {"value": len(pairs) / t_ours, "ms_per_pframe": 1e3 * t_ours / len(pairs), "pframes": len(pairs)},
        "stock_torchvision_same_gpu": {"value": args.stock_pframes / t_stock,
                                       "ms_per_pframe": 1e3 * t_stock / args.stock_pframes,
                                       "pframes": args.stock_pframes},
        "speedup": (len(pairs) / t_ours) / (args.stock_pframes / t_stock),
        "epe_vs_stock_px": {"mean": epe.mean().item(), "max_over_frames": epe.max().item()},
    }
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
