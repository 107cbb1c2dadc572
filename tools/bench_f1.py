#!/usr/bin/env python
"""Row f-1 measured (developer / bench tool): per GRU iteration at 1920x1088, CUDA events, 268 MB zeroed between
timed iterations (L2 flush):
  stock      rdvc lookup (B, 324, h, w) fp32  +  MotionEncoder.convcorr1 as PyTorch runs it (cuDNN 1x1 + ReLU)
  fused      rdvc lookup -> K-major bf16 rows  +  tcgen05 1x1 GEMM with bias + ReLU epilogue (rdvc_corr_lookup_conv1x1)
and each piece alone.  Prints one JSON line."""
import json, os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import torch
import rdvc_corr_b200 as rc
from torchvision.models.optical_flow import raft_large

dev = torch.device("cuda", 0)
B, D, h, w = 1, 256, 136, 240
torch.manual_seed(0)
model = raft_large(weights=None, corr_block=rc.TVCorrBlock()).eval().to(dev)
conv = model.update_block.motion_encoder.convcorr1
g = torch.Generator(device=dev).manual_seed(0)
f1 = torch.randn(B, D, h, w, device=dev, generator=g)
f2 = torch.randn(B, D, h, w, device=dev, generator=g)
ys, xs = torch.meshgrid(torch.arange(h, device=dev), torch.arange(w, device=dev), indexing="ij")
coords = [torch.stack([xs, ys], 0).float()[None] + 0.5 * (k + 1) * torch.randn(1, 2, h, w, device=dev, generator=g) for k in range(12)]
flush = torch.empty(268 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timed(fn, reps=12):
    ms = []
    for k in range(reps + 2):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(coords[k % 12]); e1.record()
        torch.cuda.synchronize()
        if k >= 2:
            ms.append(e0.elapsed_time(e1))
    ms.sort()
    return 1e3 * ms[len(ms) // 2]


out = {"shape": [B, D, h, w], "unit": "us per GRU iteration (median of 12, L2 flushed)"}
for vol in (torch.float32, torch.bfloat16):
    blk = rc.TVCorrBlock(volume_dtype=vol)
    blk.build_pyramid(f1, f2)
    name = "fp32_volume" if vol == torch.float32 else "bf16_volume"
    r = {}
    buf = torch.empty(B, 324, h, w, device=dev)
    feat = torch.zeros(rc.corr_block.feat_shape(B, h, w), dtype=torch.bfloat16, device=dev)
    packed = rc.corr_block.PackedConv1x1(conv[0].weight, conv[0].bias, 4, 4, torch.bfloat16, dev)
    with torch.no_grad():
        r["lookup_nchw_fp32"] = timed(lambda c: rc.index_pyramid(blk._pyr, c, 4, out=buf))
        r["lookup_kmajor_bf16"] = timed(lambda c: rc.corr_block.index_pyramid_kmajor(blk._pyr, c, 4, torch.bfloat16, out=feat))
        r["conv1x1_gemm_alone"] = timed(lambda c: rc.corr_block.conv1x1(feat, packed, B, h, w))
        r["fused_lookup_convcorr1"] = timed(lambda c: blk.index_pyramid_convcorr1(c, conv[0].weight, conv[0].bias))
        for tf32 in (True, False):
            torch.backends.cudnn.allow_tf32 = tf32
            r["stock_convcorr1_cudnn_" + ("tf32" if tf32 else "fp32")] = timed(lambda c: conv(buf))
            r["stock_lookup_plus_convcorr1_" + ("tf32" if tf32 else "fp32")] = timed(
                lambda c: conv(rc.index_pyramid(blk._pyr, c, 4, out=buf)))
        torch.backends.cudnn.allow_tf32 = True
        with torch.autocast("cuda", dtype=torch.float16):
            r["amp_stock_lookup_plus_convcorr1"] = timed(lambda c: conv(rc.index_pyramid(blk._pyr, c, 4, out=buf)))
            r["amp_fused_lookup_convcorr1"] = timed(lambda c: blk.index_pyramid_convcorr1(c, conv[0].weight, conv[0].bias))
    out[name] = {k: round(v, 2) for k, v in r.items()}
    blk.release()
print(json.dumps(out), flush=True)
