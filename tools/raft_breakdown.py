#!/usr/bin/env python
"""Where a 1080p P-frame's RAFT time goes (developer tool): CUDA-event breakdown of rc.raft_flow's pieces,
for fp32 / fp16-autocast, NCHW / channels_last."""
import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import torch, torch.nn.functional as F
import rdvc_corr_b200 as rc
from torchvision.models.optical_flow import raft_large
from torchvision.models.optical_flow._utils import upsample_flow

from torchvision.models.optical_flow._utils import make_coords_grid as _grid
dev = torch.device("cuda", 0)
h, w = 1088, 1920
g = torch.Generator(device=dev).manual_seed(0)
a = torch.rand(1, 3, h, w, device=dev, generator=g); b = torch.rand(1, 3, h, w, device=dev, generator=g)


def run(amp, cl, tf32=False):
    torch.backends.cudnn.allow_tf32 = tf32; torch.backends.cuda.matmul.allow_tf32 = tf32
    torch.manual_seed(0)
    m = raft_large(weights=None, corr_block=rc.TVCorrBlock()).eval().to(dev)
    if cl: m = m.to(memory_format=torch.channels_last)
    x1, x2 = (a.contiguous(memory_format=torch.channels_last), b.contiguous(memory_format=torch.channels_last)) if cl else (a, b)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    def once(timed):
        t = {}
        def mark(name, e0, e1): t[name] = t.get(name, 0.0) + (e0.elapsed_time(e1) if timed else 0.0)
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16, enabled=amp):
            e = [ev() for _ in range(8)]
            e[0].record()
            fm = m.feature_encoder(torch.cat([x1, x2], 0)); f1, f2 = torch.chunk(fm, 2, 0)
            e[1].record()
            m.corr_block.build_pyramid(f1, f2)
            e[2].record()
            ctx = m.context_encoder(x1)
            hs = m.update_block.hidden_state_size
            hid, cx = torch.split(ctx, [hs, ctx.shape[1] - hs], 1); hid = torch.tanh(hid); cx = F.relu(cx)
            e[3].record()
            c0 = _grid(1, h // 8, w // 8, dev); c1 = c0.clone()
            lk = [(ev(), ev(), ev()) for _ in range(12)]
            for it in range(12):
                lk[it][0].record()
                cf = m.corr_block.index_pyramid(centroids_coords=c1)
                lk[it][1].record()
                hid, d = m.update_block(hid, cx, cf, c1 - c0)
                c1 = c1 + d
                lk[it][2].record()
            e[4].record()
            up = upsample_flow(flow=(c1 - c0), up_mask=m.mask_predictor(hid))
            e[5].record()
        torch.cuda.synchronize()
        if timed:
            return {"feature_enc": e[0].elapsed_time(e[1]), "corr_build": e[1].elapsed_time(e[2]), "context_enc": e[2].elapsed_time(e[3]),
                    "12x lookup": sum(x[0].elapsed_time(x[1]) for x in lk), "12x update_block": sum(x[1].elapsed_time(x[2]) for x in lk),
                    "mask+upsample": e[4].elapsed_time(e[5]), "total": e[0].elapsed_time(e[5])}
    once(False); once(False)
    r = once(True)
    print(f"amp={amp} channels_last={cl} tf32={tf32}: " + "  ".join(f"{k} {v:.2f}" for k, v in r.items()), flush=True)

for amp, cl, tf in [(False, False, False), (False, False, True), (True, False, False), (False, True, True), (True, True, False)]:
    run(amp, cl, tf)
