#!/usr/bin/env python
"""GPU bring-up diagnostics (developer tool, not a test): each stage runs in its own
process under a timeout so one faulting kernel cannot take the rest down.

    python tests/devtools/bringup.py all            # on the GPU box
    python tests/devtools/bringup.py lookup|build_small|build_odd|build_ref|perf
"""
from __future__ import annotations

import os
import subprocess
import sys
import time

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _imports():
    import numpy as np
    import torch
    import rdvc_corr_b200 as rc
    from oracle import corr_numpy as cn
    return np, torch, rc, cn


def stage_lookup():
    np, torch, rc, cn = _imports()
    from helpers import pyramid_from_levels, rel_max
    from oracle import corr_c as cc
    lib = rc._cabi.load()
    for (B, C, h, w) in [(2, 32, 18, 22), (1, 16, 46, 80)]:
        f1, f2 = cn.synth_fmaps(B, C, h, w, seed=3)
        flat = cc.build_pyramid(f1, f2, 4)
        levels = cc.split_levels(flat, B, h, w, 4)
        pyr = pyramid_from_levels(rc, levels, B, h, w)
        for sigma in (0.0, 0.3, 4.0, 40.0):
            co = cn.synth_coords(B, h, w, sigma, seed=1)
            ref = cc.index_pyramid(flat, co, 4, 4)
            for variant in (1, 2):
                lib.rdvc_corr_set_option(0, variant)
                out = rc.index_pyramid(pyr, torch.from_numpy(co).cuda(), 4)
                torch.cuda.synchronize()
                print(f"lookup {B}x{C}x{h}x{w} sigma={sigma} variant={variant} rel={rel_max(out.cpu().numpy(), ref):.3e}")
        lib.rdvc_corr_set_option(0, 0)


def _build_case(B, D, h, w, tile, vol_dtype_name="float32", levels=4, verbose=True, mode=1):
    np, torch, rc, cn = _imports()
    from helpers import bf16_round, rel_max, ref_pyramid_linear
    lib = rc._cabi.load()
    vol_dtype = getattr(torch, vol_dtype_name)
    f1, f2 = cn.synth_fmaps(B, D, h, w, seed=5)
    lib.rdvc_corr_set_option(1, tile)
    lib.rdvc_corr_set_option(4, mode)
    pyr = rc.build_pyramid(torch.from_numpy(f1).cuda(), torch.from_numpy(f2).cuda(), levels, vol_dtype)
    torch.cuda.synchronize()
    lib.rdvc_corr_set_option(1, 0)
    lib.rdvc_corr_set_option(4, 0)
    if mode == 1:
        ref = cn.build_pyramid(bf16_round(f1), bf16_round(f2), levels)  # same rounded operands, fp64 math
    else:
        ref = ref_pyramid_linear(f1, f2, levels)
    ok = True
    for l in range(levels):
        got = pyr.level(l)[:, 0].float().cpu().numpy()
        r = rel_max(got, ref[l])
        tol = 1e-4 if vol_dtype == torch.float32 else 8e-3
        flag = "OK " if r < tol else "BAD"
        ok &= r < tol
        print(f"build mode={mode} B{B} D{D} {h}x{w} tile={tile} {vol_dtype_name} level{l}: rel={r:.3e} {flag}")
        if r >= tol and verbose:
            bad = np.abs(got - ref[l]) > tol * np.abs(ref[l]).max()
            print("   bad fraction", bad.mean(), "nan", np.isnan(got).mean(), "zero", (got == 0).mean())
            rows = np.where(bad.reshape(bad.shape[0], -1).any(axis=1))[0]
            print("   bad query rows (first 20):", rows[:20], "count", rows.size, "of", bad.shape[0])
            ys, xs = np.where(bad.any(axis=0))
            print("   bad y:", np.unique(ys)[:32], "bad x:", np.unique(xs)[:32])
            i = rows[0] if rows.size else 0
            print("   got[row0, :2, :8]", got[i, :2, :8])
            print("   ref[row0, :2, :8]", ref[l][i, :2, :8])
    return ok


def stage_build_small():
    _build_case(1, 64, 16, 16, 1)       # exactly one 16x16 tile, 2 m-blocks, one k-slab
    _build_case(1, 256, 16, 16, 1)      # four k-slabs
    _build_case(1, 64, 16, 32, 2)       # 8x32 tiles


def stage_build_odd():
    _build_case(2, 64, 18, 22, 1)
    _build_case(2, 64, 18, 22, 2)
    _build_case(1, 128, 46, 80, 1, "bfloat16")
    _build_case(1, 128, 33, 47, 1, levels=3)


def stage_build_ref():
    _build_case(1, 256, 46, 80, 1)
    _build_case(1, 256, 46, 80, 2)
    _build_case(1, 256, 46, 80, 1, "bfloat16")


def stage_build_linear():
    _build_case(1, 64, 16, 16, 0, mode=2)
    _build_case(2, 64, 18, 22, 0, mode=2)
    _build_case(1, 128, 33, 47, 0, levels=3, mode=2)
    _build_case(1, 256, 46, 80, 0, mode=2)
    _build_case(1, 256, 46, 80, 0, "bfloat16", mode=2)
    _build_case(2, 128, 24, 40, 0, "bfloat16", mode=2)


def stage_perf():
    np, torch, rc, cn = _imports()
    lib = rc._cabi.load()
    B, D, h, w = 1, 256, 136, 240
    g = torch.Generator(device="cuda").manual_seed(0)
    f1 = torch.randn(B, D, h, w, device="cuda", generator=g)
    f2 = torch.randn(B, D, h, w, device="cuda", generator=g)
    co = torch.from_numpy(cn.synth_coords(B, h, w, 2.0, seed=1)).cuda()
    for vol in (torch.float32, torch.bfloat16):
        for mode, tile, msplit in ((2, 0, 0), (2, 0, 1), (2, 0, 3), (1, 1, 0), (1, 2, 0)):
            lib.rdvc_corr_set_option(4, mode)
            lib.rdvc_corr_set_option(1, tile)
            lib.rdvc_corr_set_option(2, msplit)
            blk = rc.TVCorrBlock(volume_dtype=vol)
            for _ in range(2):
                blk.build_pyramid(f1, f2)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                blk.build_pyramid(f1, f2)
            e1.record(); torch.cuda.synchronize()
            tb = e0.elapsed_time(e1) / 5
            nbytes = lib.rdvc_corr_pyramid_bytes(B, h, w, 4, rc.RDVC_DT_F32 if vol == torch.float32 else rc.RDVC_DT_BF16, rc.RDVC_LAYOUT_TILED)
            print(f"build 1080p {vol} mode={mode} tile={tile} msplit={msplit}: {tb:.3f} ms  ({nbytes / tb / 1e6:.0f} GB/s of pyramid bytes)")
            for variant in (1, 2):
                lib.rdvc_corr_set_option(0, variant)
                for _ in range(2):
                    blk.index_pyramid(co)
                e0.record()
                for _ in range(12):
                    blk.index_pyramid(co)
                e1.record(); torch.cuda.synchronize()
                print(f"   lookup variant={variant}: {e0.elapsed_time(e1) / 12 * 1000:.1f} us/iter")
            lib.rdvc_corr_set_option(0, 0)
            blk.release()
    lib.rdvc_corr_set_option(1, 0)
    lib.rdvc_corr_set_option(2, 0)
    lib.rdvc_corr_set_option(4, 0)
    # stock torchvision on the same GPU for scale
    from oracle import tv_corr as tv
    blk = tv.tv_corr_block()
    blk.build_pyramid(f1, f2); torch.cuda.synchronize()
    t0 = time.time(); blk.build_pyramid(f1, f2); torch.cuda.synchronize()
    t1 = time.time()
    for _ in range(3):
        blk.index_pyramid(centroids_coords=co)
    torch.cuda.synchronize(); t2 = time.time()
    print(f"torchvision CUDA fp32: build {1e3 * (t1 - t0):.2f} ms, lookup {1e3 * (t2 - t1) / 3:.2f} ms/iter")


STAGES = {"lookup": stage_lookup, "build_small": stage_build_small, "build_odd": stage_build_odd,
          "build_ref": stage_build_ref, "build_linear": stage_build_linear, "perf": stage_perf}

if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which == "all":
        for name in STAGES:
            print(f"===== {name} =====", flush=True)
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), name], timeout=240)
                print(f"===== {name}: exit {r.returncode} =====", flush=True)
            except subprocess.TimeoutExpired:
                print(f"===== {name}: TIMEOUT =====", flush=True)
    else:
        STAGES[which]()
