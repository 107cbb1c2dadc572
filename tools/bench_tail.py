#!/usr/bin/env python
"""Row f-2 (last sub-item) measured (developer / bench tool), 1920x1088, CUDA events, 268 MB zeroed between iterations:
  stock   FeatureEncoder.conv as PyTorch runs it (cuDNN 1x1, 128 -> 256, both images) + rdvc_corr_build (pack + GEMM)
  fused   TVCorrBlock.build_pyramid_from_encoder: pack of the 128-channel activations + encoder-tail GEMM + build GEMM
`python tools/bench_tail.py prof` runs the fused path three times and exits (ncu target).  Prints one JSON line."""
import json, os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import torch
import rdvc_corr_b200 as rc

dev = torch.device("cuda", 0)
B, h, w = 1, 136, 240
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(2 * B, 128, h, w, device=dev, generator=g).relu()
conv = torch.nn.Conv2d(128, 256, 1).to(dev)
flush = torch.empty(268 * 1024 * 1024, dtype=torch.uint8, device=dev)
lib = rc._cabi.load()


def timed(fn, reps=10):
    ms = []
    for k in range(reps + 2):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        if k >= 2:
            ms.append(e0.elapsed_time(e1))
    ms.sort()
    return 1e3 * ms[len(ms) // 2]


out = {"shape": [B, 128, h, w], "unit": "us (median of 10, L2 flushed)"}
for vol in (torch.float32, torch.bfloat16):
    blk = rc.TVCorrBlock(volume_dtype=vol)
    x1, x2 = torch.chunk(x, 2, 0)
    with torch.no_grad():
        if "prof" in sys.argv:
            for _ in range(3):
                blk.build_pyramid_from_encoder(x1, x2, conv.weight, conv.bias)
            torch.cuda.synchronize()
            continue
        r = {}
        def stock():
            f = conv(x)
            blk.build_pyramid(*torch.chunk(f, 2, 0))
        r["stock_conv_cudnn_tf32_plus_build"] = timed(stock)
        r["stock_conv_cudnn_tf32_alone"] = timed(lambda: conv(x))
        f = conv(x)
        r["build_alone_from_fp32_fmaps"] = timed(lambda: blk.build_pyramid(*torch.chunk(f, 2, 0)))
        r["fused_pack128_tail_build"] = timed(lambda: blk.build_pyramid_from_encoder(x1, x2, conv.weight, conv.bias))
        with torch.autocast("cuda", dtype=torch.float16):
            r["amp_stock_conv_plus_build"] = timed(stock)
        out["fp32_volume" if vol == torch.float32 else "bf16_volume"] = {k: round(v, 1) for k, v in r.items()}
    blk.release()
if "prof" not in sys.argv:
    print(json.dumps(out), flush=True)
