#!/usr/bin/env python
"""Measurement of the motion-compensation network ("next" row f-4) at 1080p (developer/bench tool).

One 1920x1080 frame: warped_ref, ref_frame (B, 3, H, W), flow (B, 2, H, W) -> refined frame, the call the reference
makes at R:codec_processing.py:1458.  Reports the whole network (1 pack + 8 convolution launches) with CUDA events
(a 256 MB buffer is zeroed between iterations to flush L2), each layer type alone, the algorithmic FLOPs and HBM bytes
against the measured peaks, and beside it the same network through PyTorch/cuDNN on the same GPU (fp32 as the reference
runs it; fp16 autocast + channels_last as the fastest stock setting) and on this box's CPU cores.  One JSON line.
"""
import json
import os
import sys
import time

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.nn.functional as F

import rdvc_corr_b200 as rc
from rdvc_corr_b200 import mcn as hm

B, H, W = int(os.environ.get("MCN_B", 1)), 1080, 1920
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
a = torch.rand(B, 3, H, W, device=dev, generator=g)
r = torch.rand(B, 3, H, W, device=dev, generator=g)
f = torch.randn(B, 2, H, W, device=dev, generator=g) * 4
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

z = np.load(os.path.join(ROOT, "tests", "golden", "mcn.npz"))
net = rc.MotionCompensationNetwork()
net.load_state_dict({k[6:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("state:")})
net = net.eval().to(dev)
layers = [(w.to(dev), b.to(dev)) for w, b in net.folded_layers()]


def stock(a, f, r, layers=layers):
    x = F.leaky_relu(F.conv2d(torch.cat([a, f, r], 1), layers[0][0], layers[0][1], padding=2), 0.2)
    for i in range(3):
        t = F.leaky_relu(F.conv2d(x, *layers[1 + 2 * i], padding=1), 0.2)
        x = F.leaky_relu(F.conv2d(t, *layers[2 + 2 * i], padding=1) + x, 0.2)
    return a * torch.sigmoid(F.conv2d(x, *layers[7], padding=2))


def timed(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


with torch.no_grad():
    if os.environ.get("MCN_KERNEL"):
        assert rc._cabi.load().rdvc_corr_set_option(14, int(os.environ["MCN_KERNEL"])) == 0
    sweep = {}
    for d in [int(v) for v in os.environ.get("MCN_PF_SWEEP", "").split(",") if v]:
        rc._cabi.load().rdvc_corr_set_option(13, d)
        sweep[d] = round(timed(lambda: net(a, f, r), n=20), 3)
    if sweep:
        print("prefetch-distance sweep (ms):", sweep, file=sys.stderr)
        rc._cabi.load().rdvc_corr_set_option(13, 0)
    ours_ms = timed(lambda: net(a, f, r), n=20)
    got = net(a, f, r)
    stock_fp32_ms = timed(lambda: stock(a, f, r))
    want = stock(a, f, r)
    torch.backends.cudnn.benchmark = True
    a_cl, f_cl, r_cl = (t.contiguous(memory_format=torch.channels_last) for t in (a, f, r))
    layers_cl = [(w.contiguous(memory_format=torch.channels_last), b) for w, b in layers]
    def amp_cl():
        with torch.autocast("cuda", dtype=torch.float16):
            return stock(a_cl, f_cl, r_cl, layers_cl)
    stock_amp_cl_ms = timed(amp_cl)
    def amp_nchw():
        with torch.autocast("cuda", dtype=torch.float16):
            return stock(a, f, r)
    stock_amp_ms = timed(amp_nchw)

    # per-layer timings of ours (planes prepared once)
    lib = rc._cabi.load()
    plane = hm.pack_input(a, f, r)
    _, dev_w, _, masks, biases = net._prepare(dev)
    t_pack = timed(lambda: hm.pack_input(a, f, r), n=10)
    x0 = hm.conv_layer(plane, dev_w[0], masks[0], torch.from_numpy(biases[0]), 5, 1, B, H, W)
    t_first = timed(lambda: hm.conv_layer(plane, dev_w[0], masks[0], torch.from_numpy(biases[0]), 5, 1, B, H, W), n=10)
    t_mid = timed(lambda: hm.conv_layer(x0, dev_w[1], masks[1], torch.from_numpy(biases[1]), 3, 1, B, H, W), n=10)
    t_mid_res = timed(lambda: hm.conv_layer(x0, dev_w[2], masks[2], torch.from_numpy(biases[2]), 3, 1, B, H, W, residual=plane), n=10)
    out = torch.empty_like(a)
    st = torch.cuda.current_stream().cuda_stream
    t_last = timed(lambda: lib.rdvc_mcn_conv_out(x0.data_ptr(), dev_w[7].data_ptr(), masks[7], biases[7].ctypes.data, 5, 3,
                                                 a.data_ptr(), out.data_ptr(), B, H, W, st), n=10)

    # CPU: the reference's own setting (fp32, all host cores), at a reduced size, scaled by pixels
    torch.set_num_threads(os.cpu_count() or 1)
    hc, wc = 270, 480
    ac, fc, rc_ = (t[:, :, :hc, :wc].cpu() for t in (a, f, r))
    lc = [(w.cpu(), b.cpu()) for w, b in layers]
    stock(ac, fc, rc_, lc)
    t0 = time.perf_counter(); stock(ac, fc, rc_, lc); cpu_ms = (time.perf_counter() - t0) * 1e3 * (H * W) / (hc * wc)

px = B * H * W
flops = 2 * px * (25 * 8 * 32 + 6 * 9 * 32 * 32 + 25 * 32 * 3)            # the network's own arithmetic
plane_b = B * H * ((W + 1) // 2) * 128
hbm = px * 4 * (8 + 3 + 3) + plane_b * (1 + 2 * 7 + 3)                      # fp32 in/out + every plane written once and read once (+3 residual reads)
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
hbm_peak = peaks.get("hbm_gbs", 6650.0)
print(json.dumps({
    "metric": "motion_compensation_network_1080p_ms", "B": B, "kernel_option_14": int(os.environ.get("MCN_KERNEL", 0)), "ours_ms": round(ours_ms, 3),
    "ours_layers_ms": {"pack_input": round(t_pack, 3), "conv5x5_8to32": round(t_first, 3), "conv3x3": round(t_mid, 3),
                       "conv3x3_residual": round(t_mid_res, 3), "conv5x5_32to3_sigmoid_mul": round(t_last, 3)},
    "torch_cudnn_fp32_ms": round(stock_fp32_ms, 3), "torch_cudnn_fp16_autocast_ms": round(stock_amp_ms, 3),
    "torch_cudnn_fp16_autocast_channels_last_ms": round(stock_amp_cl_ms, 3),
    "torch_cpu_fp32_ms_scaled_from_270x480": round(cpu_ms, 1), "cpu_cores": os.cpu_count(),
    "speedup_vs_cudnn_fp32": round(stock_fp32_ms / ours_ms, 2), "speedup_vs_best_cudnn": round(min(stock_amp_ms, stock_amp_cl_ms, stock_fp32_ms) / ours_ms, 2),
    "network_gflop": round(flops / 1e9, 1), "achieved_tflops_network": round(flops / (ours_ms * 1e-3) / 1e12, 1),
    "algorithmic_hbm_bytes": hbm, "achieved_gbs": round(hbm / (ours_ms * 1e-3) / 1e9, 1), "hbm_peak_gbs": hbm_peak,
    "frac_of_hbm_peak": round(hbm / (ours_ms * 1e-3) / 1e9 / hbm_peak, 3),
    "max_abs_diff_vs_cudnn_fp32": float((got - want).abs().max()),
    "l2": "256 MB buffer zeroed between timed iterations"}))
