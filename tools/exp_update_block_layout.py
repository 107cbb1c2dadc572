#!/usr/bin/env python
"""Developer experiment: 12 iterations of the STOCK update block (TV:raft.py:278-285) at 1080p / B = 9 under fp16 autocast,
captured in a CUDA graph, with NCHW tensors (as RAFT runs it) vs channels_last module + inputs.  The ncu launch list of a
P-frame (profiles/r02raft_launches_summary.csv) shows cuDNN's nchw<->nhwc transposes and the separate bias adds at
~40 % of the time; this measures whether the memory format removes them."""
import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import torch
from torchvision.models.optical_flow import raft_large

dev = torch.device("cuda", 0)
B, h, w = int(os.environ.get("B", 9)), 136, 240
torch.manual_seed(0)
model = raft_large(weights=None).eval().to(dev)
ub = model.update_block
g = torch.Generator(device=dev).manual_seed(0)
hidden0 = torch.tanh(torch.randn(B, 128, h, w, device=dev, generator=g))
context0 = torch.relu(torch.randn(B, 128, h, w, device=dev, generator=g))
corr0 = torch.randn(B, 324, h, w, device=dev, generator=g)
flow0 = torch.randn(B, 2, h, w, device=dev, generator=g)


def loop(hidden, context, corr, flow):
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        for _ in range(12):
            hidden, delta = ub(hidden, context, corr, flow)
            flow = flow + delta
    return hidden, flow


ref = None
for name, fmt in (("NCHW", torch.contiguous_format), ("channels_last", torch.channels_last)):
    ub.to(memory_format=fmt)
    args = [t.half().contiguous(memory_format=fmt) for t in (hidden0, context0, corr0)] + [flow0.contiguous(memory_format=fmt)]
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):
            loop(*args)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        out = loop(*args)
    graph.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        graph.replay()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    fl = out[1].float().contiguous()
    if ref is None:
        ref = fl.clone()
    print(f"{name:14s} 12 update iterations, B={B}: {ms:7.2f} ms ({ms / B:5.2f} ms per P-frame)   max |flow - NCHW| {float((fl - ref).abs().max()):.3e}"
          f"   out format channels_last={out[0].is_contiguous(memory_format=torch.channels_last)}", flush=True)
