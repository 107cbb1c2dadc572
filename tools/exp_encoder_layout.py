#!/usr/bin/env python
"""Developer experiment: RAFT's context encoder (BatchNorm, eval) and feature-encoder trunk (InstanceNorm) at 1080p under
fp16 autocast, CUDA-graph captured, NCHW (as RAFT runs them) vs channels_last module + input."""
import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import torch
from torchvision.models.optical_flow import raft_large

dev = torch.device("cuda", 0)
B, h, w = int(os.environ.get("B", 9)), 1088, 1920
torch.manual_seed(0)
model = raft_large(weights=None).eval().to(dev)
x0 = torch.rand(B, 3, h, w, device=dev) * 2 - 1
fe = model.feature_encoder
mods = {"context_encoder": model.context_encoder,
        "feature_trunk": torch.nn.Sequential(fe.convnormrelu, fe.layer1, fe.layer2, fe.layer3)}
for name, m in mods.items():
    ref = None
    for fname, fmt in (("NCHW", torch.contiguous_format), ("channels_last", torch.channels_last)):
        m.to(memory_format=fmt)
        x = x0.contiguous(memory_format=fmt)

        def run():
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
                return m(x)
        s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2):
                run()
        torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out = run()
        graph.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            graph.replay()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        o = out.float().contiguous()
        if ref is None:
            ref = o.clone()
        print(f"{name:16s} {fname:14s} B={B}: {ms:7.2f} ms ({ms / B:5.2f} per image)  max |out - NCHW| {float((o - ref).abs().max()):.3e}", flush=True)
        del graph, out
