#!/usr/bin/env python
"""Measurement of the frame-preparation kernel ("next" row, caller side) for a 1080p frame (developer/bench tool):
ours (uint8 upload + one kernel) vs the reference's path (TF.to_tensor + TF.resize(antialias=True) on the CPU, then
.to(device)), for the two RAFT input sizes.  Prints one JSON line per size."""
import json, os, sys, time
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import numpy as np, torch
import torchvision.transforms.functional as TF
import rdvc_corr_b200 as rc

dev = torch.device("cuda", 0)
rng = np.random.default_rng(1)      # a deterministic uint8 frame with smooth and noisy content
yy, xx = np.meshgrid(np.linspace(0, 6.0, 1080), np.linspace(0, 9.0, 1920), indexing="ij")
frame = np.clip(127.5 + 100.0 * np.sin(yy[..., None] + np.arange(3)) * np.cos(xx[..., None]) + rng.integers(-20, 21, (1080, 1920, 3)),
                0, 255).astype(np.uint8)
pinned = torch.from_numpy(frame).pin_memory()
torch.set_num_threads(os.cpu_count() or 1)
for size in ((1088, 1920), (368, 640)):
    def ours():
        return rc.frame_to_tensor(pinned, size, dev)
    def ref():
        return TF.resize(TF.to_tensor(frame), list(size), antialias=True).unsqueeze(0).to(dev)
    for fn in (ours, ref):
        for _ in range(3): fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20): ours()
    torch.cuda.synchronize(); t_ours = (time.perf_counter() - t0) / 20
    t0 = time.perf_counter()
    for _ in range(5): ref()
    torch.cuda.synchronize(); t_ref = (time.perf_counter() - t0) / 5
    d_t = torch.from_numpy(frame).to(dev)
    out = torch.empty(1, 3, *size, device=dev)
    lib = rc._cabi.load(); st = torch.cuda.current_stream().cuda_stream
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3): lib.rdvc_preprocess_frame(d_t.data_ptr(), 1080, 1920, 3, out.data_ptr(), size[0], size[1], st)
    e0.record()
    for _ in range(50): lib.rdvc_preprocess_frame(d_t.data_ptr(), 1080, 1920, 3, out.data_ptr(), size[0], size[1], st)
    e1.record(); torch.cuda.synchronize()
    print(json.dumps({"metric": "preprocess_frame_raft_1080p", "out_hw": size, "ours_host_to_device_tensor_ms": round(t_ours * 1e3, 3),
                      "kernel_us": round(e0.elapsed_time(e1) / 50 * 1e3, 1),
                      "reference_cpu_path_ms": round(t_ref * 1e3, 2), "cpu_cores": os.cpu_count(),
                      "speedup": round(t_ref / t_ours, 1),
                      "max_abs_diff": float((ours() - ref()).abs().max())}), flush=True)
