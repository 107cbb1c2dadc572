#!/usr/bin/env python
"""BASELINE.json config 5: resolution sweep 448x256 -> 3840x2160, fp32 and bf16 volume.
Timing (CUDA events) + parity on sampled query rows against a torch fp32 matmul of the same
bf16-rounded operands (the CPU oracle cannot hold the larger volumes).  Developer/bench tool."""
import json, os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import torch
import rdvc_corr_b200 as rc

SHAPES = [(448, 256), (640, 368), (1280, 720), (1920, 1088), (2560, 1440), (3840, 2160)]
D, L, R, ITERS = 256, 4, 4, 12
HBM = 6537.6e9
dev = torch.device("cuda", 0)
for (W, H) in SHAPES:
    h, w = H // 8, W // 8
    N = h * w
    g = torch.Generator(device=dev).manual_seed(0)
    f1 = torch.randn(1, D, h, w, device=dev, generator=g)
    f2 = torch.randn(1, D, h, w, device=dev, generator=g)
    ys, xs = torch.meshgrid(torch.arange(h, device=dev), torch.arange(w, device=dev), indexing="ij")
    co = torch.stack([xs, ys], 0).float()[None] + 1.5 * torch.randn(1, 2, h, w, device=dev, generator=g)
    for vol in (torch.float32, torch.bfloat16):
        es = 4 if vol == torch.float32 else 2
        blk = rc.TVCorrBlock(volume_dtype=vol)
        out = torch.empty(1, 324, h, w, device=dev)
        def pair():
            blk.build_pyramid(f1, f2)
            for _ in range(ITERS):
                rc.index_pyramid(blk._pyr, co, R, out=out)
        for _ in range(2):
            pair()
        torch.cuda.synchronize()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        n = 5 if N < 60000 else 2
        tb = tl = 0.0
        for _ in range(n):
            e0.record(); blk.build_pyramid(f1, f2); e1.record()
            for _ in range(ITERS):
                rc.index_pyramid(blk._pyr, co, R, out=out)
            e2.record(); torch.cuda.synchronize()
            tb += e0.elapsed_time(e1); tl += e1.elapsed_time(e2)
        tb /= n; tl /= n
        # parity on sampled query rows (all levels) vs torch fp32 on the same rounded operands
        rows = torch.tensor([0, 1, N // 3, N // 2 + 7, N - 2, N - 1], device=dev)
        a = f1.to(torch.bfloat16).float().view(D, N)[:, rows].t().double()
        worst = 0.0
        for l in range(L):
            b = torch.nn.functional.avg_pool2d(f2, 2 ** l) if l else f2
            ref = (a @ b.to(torch.bfloat16).double().view(D, -1) / 16.0).float()
            got = blk._pyr.level(l, rows)[:, 0].reshape(len(rows), -1).float()
            worst = max(worst, ((got - ref).abs().max() / ref.abs().max()).item())
        pyr_elems = sum((h >> l) * (w >> l) for l in range(L)) * N
        bytes_build = 2 * N * D * 2 + es * pyr_elems
        bytes_lookup = N * (L * 100 * es + 8 + 324 * 4)
        roof = (bytes_build + ITERS * bytes_lookup) / HBM * 1e3
        print(json.dumps({"frame": f"{W}x{H}", "fmap": [h, w], "volume": str(vol).split(".")[1],
                          "pyramid_GB": round(es * pyr_elems / 1e9, 3), "build_ms": round(tb, 3),
                          "lookup12_ms": round(tl, 3), "pairs_per_s": round(1e3 / (tb + tl), 1),
                          "roofline_ms": round(roof, 3), "frac_of_roofline": round(roof / (tb + tl), 3),
                          "max_rel_err_sampled_rows": float(f"{worst:.2e}")}), flush=True)
        blk.release(); del out
        torch.cuda.empty_cache()
