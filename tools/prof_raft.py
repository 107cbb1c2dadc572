#!/usr/bin/env python
"""ncu target (developer tool): ONE eager rc.raft_flow_sequence call on a run of 10 frames at 1080p under fp16 autocast
(what one GOP costs), bracketed by cudaProfilerStart/Stop:
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/prof_raft.py"""
import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import torch
from torchvision.models.optical_flow import raft_large
import rdvc_corr_b200 as rc

dev = torch.device("cuda", 0)
n, h, w = int(os.environ.get("PAIRS", 9)), 1088, 1920
torch.manual_seed(0)
model = raft_large(weights=None, corr_block=rc.TVCorrBlock()).eval().to(dev)
g = torch.Generator(device=dev).manual_seed(0)
base = torch.rand(1, 3, h + 64, w + 64, device=dev, generator=g)
frames = torch.cat([base[:, :, i:i + h, 2 * i:2 * i + w] for i in range(n + 1)], 0).contiguous()
with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
    for _ in range(2):
        rc.raft_flow_sequence(model, frames, 12)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    flow = rc.raft_flow_sequence(model, frames, 12)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("ok", tuple(flow.shape))
