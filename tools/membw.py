#!/usr/bin/env python
"""Calibrate HBM write / copy bandwidth with plain torch ops (developer tool)."""
import torch
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
nb = 5_659_776_000
x = torch.empty(nb // 4, dtype=torch.float32, device="cuda")
y = torch.empty(nb // 4, dtype=torch.float32, device="cuda")
ms = t(lambda: x.fill_(1.0)); print(f"fill  {nb/1e9:.2f} GB: {ms:.3f} ms -> {nb/ms/1e6:.0f} GB/s write")
ms = t(lambda: y.copy_(x)); print(f"copy  {nb/1e9:.2f} GB: {ms:.3f} ms -> {2*nb/ms/1e6:.0f} GB/s read+write")
ms = t(lambda: x.sum()); print(f"sum   {nb/1e9:.2f} GB: {ms:.3f} ms -> {nb/ms/1e6:.0f} GB/s read")
