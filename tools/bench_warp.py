#!/usr/bin/env python
"""Measurement of the fused flow-resize + warp kernel ("next" row f-4) at 1080p (developer/bench tool).

Frame 1920x1080 (C = 3), RAFT flow at 1920x1088 as in R:codec_processing.py:1446,1456.  Reports the
kernel's time (CUDA events, L2 flushed between iterations), its algorithmic bytes / time against the measured
HBM peak, and beside it the same two steps composed from the torch ops the reference uses, on the same GPU
and on this box's CPU cores.  Prints one JSON line.
"""
import json, os, sys, time
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import torch, torch.nn.functional as F
import rdvc_corr_b200 as rc

B, C, H, W, h_in, w_in = 1, 3, 1080, 1920, 1088, 1920
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev).manual_seed(0)
x = torch.rand(B, C, H, W, device=dev, generator=gen)
fl = 4.0 * torch.randn(B, 2, h_in // 8, w_in // 8, device=dev, generator=gen)
fl = F.interpolate(fl, size=(h_in, w_in), mode="bilinear", align_corners=False).contiguous()   # smooth, like a real flow
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def ref_ops(x, fl):
    """The reference's composition (resize_flow + WarpingLayer) with plain torch ops."""
    f = F.interpolate(fl, size=(H, W), mode="bilinear", align_corners=False, antialias=False)
    fs = torch.zeros_like(f); fs[:, 0] = f[:, 0] * (W / w_in); fs[:, 1] = f[:, 1] * (H / h_in)
    gy, gx = torch.meshgrid(torch.linspace(-1, 1, H, device=x.device), torch.linspace(-1, 1, W, device=x.device), indexing="ij")
    grid = torch.stack((gx, gy), 2).unsqueeze(0).repeat(B, 1, 1, 1)
    nf = torch.stack((fs[:, 0] / ((W - 1) / 2.0), fs[:, 1] / ((H - 1) / 2.0)), 3)
    return F.grid_sample(x, grid + nf, mode="bilinear", padding_mode="border", align_corners=True), fs


def timed(fn, n=20):
    for _ in range(3): fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]

# the kernel alone: back-to-back launches over 6 rotating buffer sets (6 x 84 MB >> 126 MB L2), one event pair
lib = rc._cabi.load()
sets = [(x.clone(), fl.clone(), torch.empty_like(x), torch.empty(B, 2, H, W, device=dev)) for _ in range(6)]
st = torch.cuda.current_stream().cuda_stream
def burst(n):
    for k in range(n):
        a_, f_, w_, o_ = sets[k % 6]
        lib.rdvc_motion_warp(a_.data_ptr(), f_.data_ptr(), B, C, H, W, h_in, w_in, w_.data_ptr(), o_.data_ptr(), st)
burst(12); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); burst(60); e1.record(); torch.cuda.synchronize()
ours = e0.elapsed_time(e1) / 60
ours_api = timed(lambda: rc.motion_warp(x, fl, (H, W)))
stock = timed(lambda: ref_ops(x, fl))
wa, fa = rc.motion_warp(x, fl, (H, W)); wb, fb = ref_ops(x, fl)
xc, fc = x.cpu(), fl.cpu()
torch.set_num_threads(os.cpu_count() or 1)
ref_ops(xc, fc); t0 = time.perf_counter(); ref_ops(xc, fc); cpu_ms = (time.perf_counter() - t0) * 1e3
bytes_alg = 4 * (2 * h_in * w_in + 2 * C * H * W + 2 * H * W) * B
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
print(json.dumps({
    "metric": "flow_resize_plus_warp_1080p", "ours_us": round(ours * 1e3, 1), "ours_through_python_api_us": round(ours_api * 1e3, 1), "torch_ops_same_gpu_us": round(stock * 1e3, 1),
    "torch_ops_cpu_ms": round(cpu_ms, 1), "cpu_cores": os.cpu_count(), "speedup_vs_torch_ops_gpu": round(stock / ours, 1),
    "algorithmic_bytes": bytes_alg, "achieved_gbs": round(bytes_alg / (ours * 1e-3) / 1e9, 1), "peak_gbs": peak,
    "frac_of_hbm_peak": round(bytes_alg / (ours * 1e-3) / 1e9 / peak, 3),
    "max_abs_diff_warped": float((wa - wb).abs().max()), "max_abs_diff_flow": float((fa - fb).abs().max()),
    "l2": "kernel: 6 rotating buffer sets (504 MB); torch ops / python api: 256 MB buffer zeroed between iterations"}))
