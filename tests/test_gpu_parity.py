"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI, against the oracles.

Tolerances (BASELINE.json north_star), all max-norm relative = max|a-b| / max|ref|:
  * correlation volume / pyramid vs the fp32 reference ......... 2e-2  (bf16 operands, fp32 accumulate)
  * same, vs an fp64 evaluation of the SAME rounded operands .... 1e-4  (fp32 out) / 8e-3 (bf16 out)
    (pooled levels at full size: 5e-4 -- the pooled fmap2 means are rounded to bf16 once, and a mean
     within fp32 summation-order distance of a bf16 rounding boundary may round the other way than the
     test's own avg_pool2d: ~2^-15 of the operand elements differ by one bf16 ulp)
  * lookup vs the oracle's lookup of the SAME pyramid ........... 1e-3  (observed ~1e-6)
  * RAFT flow end-point error vs stock torchvision, 12 updates .. 0.05 px mean
"""
import ctypes
import os

import numpy as np
import pytest
import torch

import rdvc_corr_b200 as rc
from helpers import bf16_round, fp16_round, pyramid_from_levels, ref_pyramid_linear, rel_max
from oracle import corr_c as cc
from oracle import corr_numpy as cn
from oracle import tv_corr as tv

pytestmark = pytest.mark.gpu

TOL_VOLUME = 2e-2
TOL_SAME_OPERANDS_F32 = 1e-4
TOL_SAME_OPERANDS_BF16 = 8e-3
TOL_SAME_OPERANDS_POOLED = 5e-4
TOL_LOOKUP = 1e-3
TOL_EPE = 0.05
SIGMAS = (0.0, 0.3, 4.0, 40.0)
MODES = {"fused": 1, "linear": 2}
ROW, TILED = rc.RDVC_LAYOUT_ROWMAJOR, rc.RDVC_LAYOUT_TILED
LAYOUTS = {"rowmajor": ROW, "tiled": TILED}
# lookup code paths: (pyramid layout, lookup variant option)
LOOKUP_PATHS = {"row_scalar": (ROW, 1), "row_vec": (ROW, 2), "tiled": (TILED, 0)}


@pytest.fixture(scope="module")
def lib():
    assert torch.cuda.is_available(), "GPU tests need a B200"
    L = rc._cabi.load()
    yield L
    for k in list(range(10)) + [12]:
        L.rdvc_corr_set_option(k, {3: 15, 5: 1}.get(k, 0))


def set_opts(lib, **kw):
    keys = {"lookup": 0, "tile": 1, "msplit": 2, "mode": 4, "tma": 5, "epi": 9, "pair": 12}
    for k, v in kw.items():
        assert lib.rdvc_corr_set_option(keys[k], v) == 0


def gpu(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


# ------------------------------------------------------------------ lookup
@pytest.mark.parametrize("path", list(LOOKUP_PATHS))
@pytest.mark.parametrize("shape", [(2, 8, 18, 22), (1, 8, 46, 80), (1, 4, 16, 16), (1, 4, 17, 19)])
def test_lookup_matches_oracle(lib, shape, path):
    B, C, h, w = shape
    layout, variant = LOOKUP_PATHS[path]
    f1, f2 = cn.synth_fmaps(B, C, h, w, seed=3)
    flat = cc.build_pyramid(f1, f2, 4)
    pyr = pyramid_from_levels(rc, cc.split_levels(flat, B, h, w, 4), B, h, w, layout=layout)
    set_opts(lib, lookup=variant)
    for sigma in SIGMAS:
        co = cn.synth_coords(B, h, w, sigma, seed=1)
        got = rc.index_pyramid(pyr, gpu(co), 4).cpu().numpy()
        assert got.shape == (B, 324, h, w)
        assert rel_max(got, cc.index_pyramid(flat, co, 4, 4)) < TOL_LOOKUP
    set_opts(lib, lookup=0)


@pytest.mark.parametrize("layout", list(LAYOUTS))
def test_lookup_golden_small_odd(lib, golden_dir, layout):
    """Pyramid and lookups straight from the torchvision-generated fixture."""
    g = np.load(os.path.join(golden_dir, "corr_small_odd.npz"))
    B, C, h, w = [int(x) for x in g["shape"]]
    pyr = pyramid_from_levels(rc, [g[f"level{l}"] for l in range(4)], B, h, w, layout=LAYOUTS[layout])
    for s in SIGMAS:
        co = cn.synth_coords(B, h, w, s, seed=1)
        got = rc.index_pyramid(pyr, gpu(co), 4).cpu().numpy()
        assert rel_max(got, g[f"lookup_sigma{s:g}"]) < TOL_LOOKUP


@pytest.mark.parametrize("layout", list(LAYOUTS))
@pytest.mark.parametrize("vol", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("radius,levels", [(3, 3), (1, 1), (2, 4)])
def test_lookup_other_radius_and_levels(lib, radius, levels, vol, layout):
    B, C, h, w = 2, 8, 18, 22
    f1, f2 = cn.synth_fmaps(B, C, h, w, seed=4)
    flat = cc.build_pyramid(f1, f2, levels)
    if vol == torch.bfloat16:
        flat = bf16_round(flat)
    pyr = pyramid_from_levels(rc, cc.split_levels(flat, B, h, w, levels), B, h, w, vol, layout=LAYOUTS[layout])
    co = cn.synth_coords(B, h, w, 2.5, seed=2)
    got = rc.index_pyramid(pyr, gpu(co), radius).cpu().numpy()
    assert got.shape == (B, levels * (2 * radius + 1) ** 2, h, w)
    assert rel_max(got, cc.index_pyramid(flat, co, levels, radius)) < TOL_LOOKUP


@pytest.mark.parametrize("layout", list(LAYOUTS))
@pytest.mark.parametrize("shape", [(1, 8, 24, 40), (2, 8, 18, 22), (1, 4, 17, 19)])
def test_lookup_bf16_volume(lib, shape, layout):
    B, C, h, w = shape
    f1, f2 = cn.synth_fmaps(B, C, h, w, seed=6)
    flat = bf16_round(cc.build_pyramid(f1, f2, 4))   # what a bf16 pyramid stores
    pyr = pyramid_from_levels(rc, cc.split_levels(flat, B, h, w, 4), B, h, w, torch.bfloat16,
                              layout=LAYOUTS[layout])
    for sigma in SIGMAS:
        co = cn.synth_coords(B, h, w, sigma, seed=3)
        got = rc.index_pyramid(pyr, gpu(co), 4).cpu().numpy()
        assert rel_max(got, cc.index_pyramid(flat, co, 4, 4)) < TOL_LOOKUP


def test_layouts_give_identical_lookups(lib):
    """The layout is storage only: the same pyramid values in either layout -> bit-identical features."""
    B, C, h, w = 2, 8, 18, 22
    f1, f2 = cn.synth_fmaps(B, C, h, w, seed=14)
    levels = cc.split_levels(cc.build_pyramid(f1, f2, 4), B, h, w, 4)
    co = gpu(cn.synth_coords(B, h, w, 6.0, seed=2))
    for vol in (torch.float32, torch.bfloat16):
        a = rc.index_pyramid(pyramid_from_levels(rc, levels, B, h, w, vol, layout=ROW), co, 4)
        b = rc.index_pyramid(pyramid_from_levels(rc, levels, B, h, w, vol, layout=TILED), co, 4)
        assert torch.equal(a, b)
    # and the tiled storage round-trips through level()/set_level()
    pyr = pyramid_from_levels(rc, levels, B, h, w, layout=TILED)
    for l in range(4):
        assert np.array_equal(pyr.level(l)[:, 0].cpu().numpy(), levels[l])


def test_lookup_integer_coords_are_exact_gathers(lib):
    B, C, h, w = 1, 8, 16, 24
    f1, f2 = cn.synth_fmaps(B, C, h, w, seed=9)
    lv = cc.split_levels(cc.build_pyramid(f1, f2, 1), B, h, w, 1)
    pyr = pyramid_from_levels(rc, lv, B, h, w)
    assert pyr.layout == TILED
    got = rc.index_pyramid(pyr, gpu(cn.make_coords_grid(B, h, w)), 4).cpu().numpy()
    q = 7 * w + 9
    for (i, j) in [(4, 4), (6, 2), (0, 8), (8, 0)]:
        assert got[0, i * 9 + j, 7, 9] == lv[0][q, 7 + j - 4, 9 + i - 4]
    assert np.all(got[0, 0 * 9 + 4, :, 0:4] == 0)  # dx = -4 at the left border: zero padding


# ------------------------------------------------------------------ build
BUILD_SHAPES = [(1, 64, 16, 16), (2, 64, 18, 22), (1, 128, 33, 47), (1, 256, 46, 80), (3, 64, 24, 40),
                (1, 192, 20, 28)]


@pytest.mark.parametrize("mode", ["fused", "linear"])
@pytest.mark.parametrize("shape", BUILD_SHAPES)
def test_build_fp32_volume(lib, shape, mode):
    B, D, h, w = shape
    f1, f2 = cn.synth_fmaps(B, D, h, w, seed=5)
    ref32 = cn.build_pyramid(f1, f2, 4)                       # fp64 math on the un-rounded inputs
    same = (cn.build_pyramid(bf16_round(f1), bf16_round(f2), 4) if mode == "fused"
            else ref_pyramid_linear(f1, f2, 4))               # fp64 math on what the kernel multiplies
    # fused mode: both tile shapes, row-major only; linear mode: both layouts
    for tile, layout in (((1, ROW), (2, ROW)) if mode == "fused" else ((0, ROW), (0, TILED))):
        set_opts(lib, mode=MODES[mode], tile=tile)
        pyr = rc.build_pyramid(gpu(f1), gpu(f2), 4, layout=layout)
        for l in range(4):
            got = pyr.level(l)[:, 0].cpu().numpy()
            assert got.shape == ref32[l].shape
            assert rel_max(got, same[l]) < TOL_SAME_OPERANDS_F32, (mode, tile, layout, l)
            assert rel_max(got, ref32[l]) < TOL_VOLUME, (mode, tile, layout, l)
        if layout == TILED:  # padding pixels of the tiled storage are exact zeros
            for l in range(4):
                hl, wl, tw, th, hp, wp = pyr._tiles(l)
                raw = pyr.storage(l)
                st = raw[:, : hp * wp].reshape(-1, hp // th, wp // tw, th, tw).permute(0, 1, 3, 2, 4).reshape(-1, hp, wp)
                assert not st[:, hl:, :].any() and not st[:, :, wl:].any() and not raw[:, hp * wp:].any()
    set_opts(lib, mode=0, tile=0)


@pytest.mark.parametrize("mode", ["fused", "linear"])
def test_build_bf16_volume(lib, mode):
    B, D, h, w = 1, 128, 46, 80
    f1, f2 = cn.synth_fmaps(B, D, h, w, seed=5)
    ref32 = cn.build_pyramid(f1, f2, 4)
    set_opts(lib, mode=MODES[mode])
    for layout in ((ROW,) if mode == "fused" else (ROW, TILED)):
        pyr = rc.build_pyramid(gpu(f1), gpu(f2), 4, torch.bfloat16, layout=layout)
        for l in range(4):
            got = pyr.level(l)[:, 0].float().cpu().numpy()
            assert rel_max(got, ref32[l]) < TOL_VOLUME
    set_opts(lib, mode=0)


@pytest.mark.parametrize("epi", [4, 8])
@pytest.mark.parametrize("vol", [torch.float32, torch.bfloat16])
def test_build_epilogue_shapes(lib, epi, vol):
    """Both epilogue shapes of the linear build (4 warps x 3 staging buffers, 8 warps x 1) for both
    storage types, on a shape with partial m-blocks, partial n-tiles and B > 1."""
    B, D, h, w = 2, 64, 18, 22
    f1, f2 = cn.synth_fmaps(B, D, h, w, seed=15)
    ref = ref_pyramid_linear(f1, f2, 4)
    tol = TOL_SAME_OPERANDS_F32 if vol == torch.float32 else TOL_SAME_OPERANDS_BF16
    set_opts(lib, epi=epi)
    for layout in (ROW, TILED):
        pyr = rc.build_pyramid(gpu(f1), gpu(f2), 4, vol, layout=layout)
        for l in range(4):
            assert rel_max(pyr.level(l)[:, 0].float().cpu().numpy(), ref[l]) < tol, (layout, l)
    set_opts(lib, epi=0)


@pytest.mark.parametrize("vol", [torch.float32, torch.bfloat16])
def test_build_cta_pair_kernel_is_bit_identical(lib, vol):
    """The opt-in CTA-pair build (tcgen05 cta_group::2, M = 256 over two SMs; option key 12 = 2) performs the
    same MMAs in the same K order as the single-CTA kernel: the pyramids must be bit-identical, including
    partial 256-row blocks (N = 396: the peer CTA's rows fall off the end), partial tiles and B > 1."""
    for (B, D, h, w) in [(2, 64, 18, 22), (1, 128, 33, 47), (1, 256, 46, 80)]:
        f1, f2 = cn.synth_fmaps(B, D, h, w, seed=17)
        set_opts(lib, pair=1)
        a = rc.build_pyramid(gpu(f1), gpu(f2), 4, vol).buffer.clone()
        set_opts(lib, pair=2)
        b = rc.build_pyramid(gpu(f1), gpu(f2), 4, vol).buffer.clone()
        assert torch.equal(a, b), (B, D, h, w)
    set_opts(lib, pair=0)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_build_half_inputs(lib, dtype):
    """Under the reference's default AMP the fmaps arrive in half precision.  fp16 inputs are
    multiplied AS fp16 (tcgen05 kind::f16 takes either format): nothing of the input is lost, so the volume is
    within 1e-3 of the fp32 reference instead of the 4e-3 of bf16 operands."""
    B, D, h, w = 1, 64, 24, 40
    f1, f2 = cn.synth_fmaps(B, D, h, w, seed=8)
    ref32 = cn.build_pyramid(f1, f2, 4)
    pyr = rc.build_pyramid(gpu(f1).to(dtype), gpu(f2).to(dtype), 4)
    rnd = fp16_round if dtype == torch.float16 else bf16_round
    same = ref_pyramid_linear(rnd(f1), rnd(f2), 4, round_fn=rnd)     # what the kernel multiplies, in fp64
    for l in range(4):
        got = pyr.level(l)[:, 0].cpu().numpy()
        assert rel_max(got, ref32[l]) < (1e-3 if dtype == torch.float16 else TOL_VOLUME), l
        assert rel_max(got, same[l]) < (TOL_SAME_OPERANDS_F32 if l == 0 else TOL_SAME_OPERANDS_POOLED), l


@pytest.mark.parametrize("levels", [1, 2, 3])
def test_build_fewer_levels(lib, levels):
    B, D, h, w = 1, 64, 18, 22
    f1, f2 = cn.synth_fmaps(B, D, h, w, seed=2)
    ref = ref_pyramid_linear(f1, f2, levels)
    pyr = rc.build_pyramid(gpu(f1), gpu(f2), levels)
    assert len(pyr.levels()) == levels
    for l in range(levels):
        assert rel_max(pyr.level(l)[:, 0].cpu().numpy(), ref[l]) < TOL_SAME_OPERANDS_F32


@pytest.mark.parametrize("tma", [0, 2])
@pytest.mark.parametrize("vol", [torch.float32, torch.bfloat16])
def test_build_other_store_paths(lib, tma, vol):
    """Linear mode without the wide TMA boxes: staged st.global (tma=0, the path odd-sized row-major
    levels take) and 32-row x 128-byte boxes (tma=2, 16-byte-aligned row pitches)."""
    B, D, h, w = 1, 64, 24, 40
    f1, f2 = cn.synth_fmaps(B, D, h, w, seed=12)
    ref = ref_pyramid_linear(f1, f2, 4)
    tol = TOL_SAME_OPERANDS_F32 if vol == torch.float32 else TOL_SAME_OPERANDS_BF16
    set_opts(lib, mode=2, tma=tma)
    for layout in (ROW, TILED):
        pyr = rc.build_pyramid(gpu(f1), gpu(f2), 4, vol, layout=layout)
        for l in range(4):
            assert rel_max(pyr.level(l)[:, 0].float().cpu().numpy(), ref[l]) < tol
    set_opts(lib, mode=0, tma=1)


def test_build_golden_ref_default(lib, golden_dir):
    """D=256, 46x80: RDVC's default RAFT size; sampled entries from torchvision's fp32 CorrBlock."""
    g = np.load(os.path.join(golden_dir, "corr_ref_default.npz"))
    B, D, h, w = [int(x) for x in g["shape"]]
    f1, f2 = cn.synth_fmaps(B, D, h, w, seed=int(g["seed"]))
    blk = rc.TVCorrBlock()
    blk.build_pyramid(gpu(f1), gpu(f2))
    for l in range(4):
        got = blk.corr_pyramid[l].reshape(-1)[torch.from_numpy(g[f"level{l}_idx"]).cuda()].cpu().numpy()
        assert rel_max(got, g[f"level{l}_val"]) < TOL_VOLUME
    own = [blk.corr_pyramid[l][:, 0].cpu().numpy() for l in range(4)]
    flat = np.concatenate([x.reshape(-1) for x in own])
    for s in SIGMAS:
        co = cn.synth_coords(B, h, w, s, seed=1)
        got = blk.index_pyramid(centroids_coords=gpu(co)).cpu().numpy()
        # vs the oracle's lookup of OUR pyramid: isolates the lookup kernel (1e-3) ...
        assert rel_max(got, cc.index_pyramid(flat, co, 4, 4)) < TOL_LOOKUP
        # ... and end to end vs the fixture (bf16 operand error carried through): 2e-2
        assert rel_max(got.reshape(-1)[::11], g[f"lookup_sigma{s:g}_stride11"]) < TOL_VOLUME


def test_build_rejects_unsupported(lib):
    f = torch.zeros(1, 32, 18, 22, device="cuda")
    with pytest.raises(ValueError, match="RDVC_E_UNSUPPORTED"):
        rc.build_pyramid(f, f, 4)
    with pytest.raises(ValueError, match="too small"):
        rc.build_pyramid(torch.zeros(1, 64, 8, 22, device="cuda"), torch.zeros(1, 64, 8, 22, device="cuda"), 4)


def test_build_is_linear_and_deterministic(lib):
    B, D, h, w = 1, 64, 24, 40
    f1, f2 = cn.synth_fmaps(B, D, h, w, seed=13)
    a = rc.build_pyramid(gpu(f1), gpu(f2), 4).buffer.clone()
    b = rc.build_pyramid(gpu(f1), gpu(f2), 4).buffer.clone()
    assert torch.equal(a, b)                                   # bit-identical reruns
    p2 = rc.build_pyramid(gpu(f1), gpu(f2 * 2.0), 4)           # power-of-two scaling is exact in bf16
    p1 = rc.build_pyramid(gpu(f1), gpu(f2), 4)
    for l in range(4):
        assert torch.equal(p2.level(l), p1.level(l) * 2.0)
    # the layout only permutes storage: row-major and tiled builds hold bit-identical values
    pr = rc.build_pyramid(gpu(f1), gpu(f2), 4, layout=ROW)
    for l in range(4):
        assert torch.equal(pr.level(l), p1.level(l))


# ------------------------------------------------------------------ full size: 1920x1088
def test_1080p_properties(lib):
    """BASELINE.json config 2 shape: the oracle cannot hold the 5.7 GB pyramid in seconds, so
    check size-independent properties + sampled rows against a torch fp32 matmul."""
    B, D, h, w = 1, 256, 136, 240
    N = h * w
    g = torch.Generator(device="cuda").manual_seed(0)
    f1 = torch.randn(B, D, h, w, device="cuda", generator=g)
    f2 = torch.randn(B, D, h, w, device="cuda", generator=g)
    blk = rc.TVCorrBlock()
    blk.build_pyramid(f1, f2)
    lv = blk.corr_pyramid
    assert [tuple(x.shape) for x in lv] == [(N, 1, 136, 240), (N, 1, 68, 120), (N, 1, 34, 60), (N, 1, 17, 30)]
    a = f1.to(torch.bfloat16).float().view(D, N)
    f2p = [f2] + [torch.nn.functional.avg_pool2d(f2, 2 ** l) for l in (1, 2, 3)]
    rows = torch.tensor([0, 1, 127, 128, 4097, 17000, N - 129, N - 1], device="cuda")
    for l in range(4):
        b = f2p[l].to(torch.bfloat16).float().view(D, -1)
        ref = (a[:, rows].t().double() @ b.double() / 16.0).float()
        got = lv[l][rows, 0].reshape(len(rows), -1)
        err = (got - ref).abs().max().item() / ref.abs().max().item()
        assert err < (TOL_SAME_OPERANDS_F32 if l == 0 else TOL_SAME_OPERANDS_POOLED), (l, err)
    # pyramid consistency: level l+1 is the 2x2 mean of level l (to bf16-operand accuracy)
    sl = slice(5000, 5256)
    for l in range(3):
        pooled = torch.nn.functional.avg_pool2d(lv[l][sl], 2)
        err = (pooled - lv[l + 1][sl]).abs().max().item() / lv[l + 1][sl].abs().max().item()
        assert err < TOL_VOLUME, (l, err)
    # checksum of checksums: sum over fmap2 pixels == fmap1 . sum(fmap2)
    tot = lv[0][rows, 0].double().sum(dim=(1, 2))
    ref = (a[:, rows].t().double() @ f2.to(torch.bfloat16).double().view(D, N).sum(dim=1)) / 16.0
    assert ((tot - ref).abs().max() / ref.abs().max()).item() < 1e-3
    # lookup at the identity grid: centre tap of level 0 is the volume diagonal
    co = gpu(cn.make_coords_grid(B, h, w))
    out = blk.index_pyramid(centroids_coords=co)
    assert out.shape == (B, 324, h, w) and out.is_contiguous() and out.dtype == torch.float32
    diag = lv[0].view(N, N).diagonal()
    assert torch.equal(out[0, 4 * 9 + 4].reshape(-1), diag)
    # and a drifting lookup against torchvision's own index_pyramid ON OUR PYRAMID
    co2 = gpu(cn.synth_coords(B, h, w, 3.0, seed=4))
    got = blk.index_pyramid(centroids_coords=co2)
    ref = tv.index_pyramid([x for x in lv], co2, 4)
    err = (got - ref).abs().max().item() / ref.abs().max().item()
    assert err < TOL_LOOKUP, err
    blk.release()


# ------------------------------------------------------------------ drop-in: torchvision RAFT
def _preprocess(frame_u8, size_hw):
    """R:codec_processing.py:751-759: to_tensor -> resize(antialias) -> [0,1], batch of 1."""
    import torchvision.transforms.functional as TF
    t = TF.resize(TF.to_tensor(frame_u8), list(size_hw), antialias=True)
    return t.unsqueeze(0)


def _seeded_raft(corr_block=None):
    from torchvision.models.optical_flow import raft_large
    torch.manual_seed(0)
    kw = {} if corr_block is None else {"corr_block": corr_block}
    return raft_large(weights=None, **kw).eval().cuda()


@pytest.mark.parametrize("vol_dtype", [torch.float32, torch.bfloat16])
def test_raft_flow_epe_vs_stock_torchvision(lib, golden_dir, vol_dtype):
    """BASELINE.json config 1 shape: R:im1.png / R:im2.png at 368x640, 12 updates, seed-0
    random-init raft_large (no weights offline), [0,1] inputs like the reference."""
    g = np.load(os.path.join(golden_dir, "frames_im1_im2.npz"))
    a = _preprocess(g["im1"], (368, 640)).cuda()
    b = _preprocess(g["im2"], (368, 640)).cuda()
    with torch.no_grad():
        ref = _seeded_raft()(a, b, num_flow_updates=12)[-1]
        blk = rc.TVCorrBlock(volume_dtype=vol_dtype)
        n0 = lib.rdvc_corr_launch_count()
        got = _seeded_raft(blk)(a, b, num_flow_updates=12)[-1]
        launched = lib.rdvc_corr_launch_count() - n0
    assert launched == 2 + 12, launched            # pack, build, 12 lookups
    epe = (got - ref).pow(2).sum(dim=1).sqrt()
    assert torch.isfinite(got).all()
    assert epe.mean().item() < TOL_EPE, epe.mean().item()
    # the CPU fixture (stock torchvision, CPU fp32) is a second, looser anchor
    cpu_ref = torch.from_numpy(g["flow_stride2"]).cuda()
    epe_cpu = (got[0, :, ::2, ::2] - cpu_ref).pow(2).sum(dim=0).sqrt()
    assert epe_cpu.mean().item() < 2 * TOL_EPE, epe_cpu.mean().item()


def test_raft_under_autocast(lib, golden_dir):
    """The reference's default GPU setting wraps RAFT in autocast (R:codec_processing.py:1436)."""
    g = np.load(os.path.join(golden_dir, "frames_im1_im2.npz"))
    a = _preprocess(g["im1"], (256, 448)).cuda()
    b = _preprocess(g["im2"], (256, 448)).cuda()
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        ref = _seeded_raft()(a, b, num_flow_updates=12)[-1]
        got = _seeded_raft(rc.TVCorrBlock())(a, b, num_flow_updates=12)[-1]
    epe = (got.float() - ref.float()).pow(2).sum(dim=1).sqrt()
    assert epe.mean().item() < TOL_EPE, epe.mean().item()


def test_raft_flow_runner_matches_forward(lib, golden_dir):
    """rc.raft_flow == RAFT.forward(...)[-1] with the same block: same modules, same order, the 11
    unused mask/upsample passes skipped (the caller keeps only flow_preds[-1], R:codec_processing.py:1444)."""
    g = np.load(os.path.join(golden_dir, "frames_im1_im2.npz"))
    a = _preprocess(g["im1"], (256, 448)).cuda()
    b = _preprocess(g["im2"], (256, 448)).cuda()
    model = _seeded_raft(rc.TVCorrBlock())
    with torch.no_grad():
        ref = model(a, b, num_flow_updates=12)
        got = rc.raft_flow(model, a, b, num_flow_updates=12)
        every = rc.raft_flow(model, a, b, num_flow_updates=12, all_predictions=True)
    assert torch.allclose(got, ref[-1], rtol=0, atol=1e-5)
    assert len(every) == 12 and all(torch.allclose(x, y, rtol=0, atol=1e-5) for x, y in zip(every, ref))
    with pytest.raises(TypeError):
        rc.raft_flow(_seeded_raft(), a, b)
    with pytest.raises(ValueError, match="divisible by 8"):
        rc.raft_flow(model, a[..., :250, :], b[..., :250, :])


def test_graphed_raft_flow_is_bit_identical(lib, golden_dir):
    """GraphedRaftFlow replays rc.raft_flow as one CUDA graph: same kernels, same order -> the same bits,
    for the captured pair and for new frames copied into the graph's static inputs."""
    g = np.load(os.path.join(golden_dir, "frames_im1_im2.npz"))
    a = _preprocess(g["im1"], (256, 448)).cuda()
    b = _preprocess(g["im2"], (256, 448)).cuda()
    model = _seeded_raft(rc.TVCorrBlock())
    runner = rc.GraphedRaftFlow(model, 12)
    with torch.no_grad():
        eager_ab = rc.raft_flow(model, a, b, 12)
        eager_ba = rc.raft_flow(model, b, a, 12)
    n0 = lib.rdvc_corr_launch_count()
    got_ab = runner(a, b)                 # warm-up + capture + replay
    captured = lib.rdvc_corr_launch_count() - n0
    got_ba = runner(b, a)                 # replay only: no launch goes through the C ABI again
    assert lib.rdvc_corr_launch_count() - n0 == captured
    assert torch.equal(got_ab, eager_ab) and torch.equal(got_ba, eager_ba)
    assert torch.equal(runner(a, b), eager_ab)
    with pytest.raises(TypeError):
        rc.GraphedRaftFlow(_seeded_raft())


def test_princeton_facade(lib):
    B, D, h, w = 1, 64, 24, 40
    f1, f2 = cn.synth_fmaps(B, D, h, w, seed=21)
    corr = rc.CorrBlock(gpu(f1), gpu(f2), num_levels=4, radius=4)
    co = cn.synth_coords(B, h, w, 1.5, seed=5)
    got = corr(gpu(co)).cpu().numpy()
    own = np.concatenate([x[:, 0].cpu().numpy().reshape(-1) for x in corr.corr_pyramid])
    assert got.shape == (B, 324, h, w)
    assert rel_max(got, cc.index_pyramid(own, co, 4, 4)) < TOL_LOOKUP


# ------------------------------------------------------------------ ABI details
def test_pair_host_entry_point(lib):
    """rdvc_corr_pair_host: host buffers in, host buffers out, same numbers as the device calls."""
    B, D, h, w, iters = 1, 64, 24, 40, 3
    f1, f2 = cn.synth_fmaps(B, D, h, w, seed=31)
    coords = np.stack([cn.synth_coords(B, h, w, 1.0 + i, seed=i) for i in range(iters)])
    out = np.empty((iters, B, 324, h, w), np.float32)
    fp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    rc_ = lib.rdvc_corr_pair_host(fp(f1), fp(f2), fp(coords), fp(out), B, D, h, w, 4, 4, iters, rc.RDVC_DT_F32)
    rc._cabi.check(rc_, "rdvc_corr_pair_host")
    blk = rc.TVCorrBlock()
    blk.build_pyramid(gpu(f1), gpu(f2))
    for i in range(iters):
        ref = blk.index_pyramid(centroids_coords=gpu(coords[i])).cpu().numpy()
        assert np.array_equal(out[i], ref)
    # the split call: two pairs in flight on the two slots, each result in its own host buffer
    out2 = [np.empty_like(out), np.empty_like(out)]
    f1b, f2b = cn.synth_fmaps(B, D, h, w, seed=32)
    for slot, (x1, x2) in enumerate(((f1, f2), (f1b, f2b))):
        rc._cabi.check(lib.rdvc_corr_pair_host_submit(fp(x1), fp(x2), fp(coords), fp(out2[slot]), B, D, h, w, 4, 4,
                                                      iters, rc.RDVC_DT_F32, slot), "submit")
    assert lib.rdvc_corr_pair_host_submit(fp(f1), fp(f2), fp(coords), fp(out2[0]), B, D, h, w, 4, 4, iters,
                                          rc.RDVC_DT_F32, 0) == -5          # slot 0 is still pending
    for slot in (0, 1):
        rc._cabi.check(lib.rdvc_corr_pair_host_wait(slot), "wait")
    assert np.array_equal(out2[0], out)
    blk.build_pyramid(gpu(f1b), gpu(f2b))
    for i in range(iters):
        assert np.array_equal(out2[1][i], blk.index_pyramid(centroids_coords=gpu(coords[i])).cpu().numpy())
    assert lib.rdvc_corr_pair_host_wait(1) == 0 and lib.rdvc_corr_pair_host_wait(5) == -5
    lib.rdvc_corr_release()


@pytest.mark.parametrize("mode,layout", [("fused", ROW), ("linear", ROW), ("linear", TILED)])
@pytest.mark.parametrize("vol", [torch.float32, torch.bfloat16])
def test_no_out_of_bounds_writes(lib, mode, layout, vol):
    """compute-sanitizer is closed on this pool, so bounds are checked with canaries: every byte
    outside the pyramid levels / lookup output / workspace (guard bands before and after, and the
    256-byte padding between levels) must survive a build + lookups on odd shapes."""
    CANARY, GUARD = 0xAB, 4096
    vd = rc.RDVC_DT_F32 if vol == torch.float32 else rc.RDVC_DT_BF16
    es = 4 if vol == torch.float32 else 2
    set_opts(lib, mode=MODES[mode])
    for (B, D, h, w) in [(2, 64, 18, 22), (1, 128, 33, 47), (1, 64, 17, 16)]:
        N = h * w
        f1, f2 = cn.synth_fmaps(B, D, h, w, seed=3)
        a, b = gpu(f1), gpu(f2)
        pb = lib.rdvc_corr_pyramid_bytes(B, h, w, 4, vd, layout)
        wb = lib.rdvc_corr_workspace_bytes(B, D, h, w)
        ob = B * 324 * N * 4
        bufs = {k: torch.full((n + 2 * GUARD,), CANARY, dtype=torch.uint8, device="cuda")
                for k, n in (("pyr", pb), ("ws", wb), ("out", ob))}
        ptr = {k: v.data_ptr() + GUARD for k, v in bufs.items()}
        assert all(p % 256 == 0 for p in ptr.values())
        st = torch.cuda.current_stream().cuda_stream
        rc._cabi.check(lib.rdvc_corr_build(a.data_ptr(), b.data_ptr(), B, D, h, w, rc.RDVC_DT_F32, ptr["pyr"],
                                           vd, layout, 4, ptr["ws"], wb, st), "build")
        co = gpu(cn.synth_coords(B, h, w, 5.0, seed=1))
        for variant in (1, 2):
            set_opts(lib, lookup=variant)
            rc._cabi.check(lib.rdvc_corr_lookup(ptr["pyr"], vd, layout, co.data_ptr(), B, h, w, 4, 4, ptr["out"],
                                                st), "lookup")
        torch.cuda.synchronize()
        for k, v in bufs.items():
            assert bool((v[:GUARD] == CANARY).all()) and bool((v[-GUARD:] == CANARY).all()), (k, "guard band")
        body = bufs["pyr"][GUARD:-GUARD]
        for l in range(4):
            off = lib.rdvc_corr_level_offset_bytes(B, h, w, l, vd, layout)
            end = off + B * N * lib.rdvc_corr_level_image_elems(h, w, l, vd, layout) * es
            nxt = lib.rdvc_corr_level_offset_bytes(B, h, w, l + 1, vd, layout)
            assert bool((body[end:nxt] == CANARY).all()), ("padding after level", l)
            lvl = body[off:end].view(vol).float()
            assert torch.isfinite(lvl).all()          # every element was written (0xABAB.. is finite but
            assert not bool((body[off:end] == CANARY).all())   # ... the level is not all canary)
        assert torch.isfinite(bufs["out"][GUARD:-GUARD].view(torch.float32)).all()
    set_opts(lib, mode=0, lookup=0)


def test_runs_on_callers_stream(lib):
    B, D, h, w = 1, 64, 24, 40
    f1, f2 = cn.synth_fmaps(B, D, h, w, seed=41)
    a, b = gpu(f1), gpu(f2)
    co = gpu(cn.synth_coords(B, h, w, 2.0, seed=1))
    ref = rc.CorrBlock(a, b)(co)
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        got = rc.CorrBlock(a, b)(co)
    s.synchronize()
    assert torch.equal(got, ref)


# ------------------------------------------------------------------ next row: flow resize + warp
TOL_FLOW_RESIZE = 1e-5    # relative to max(1, max|flow|): fp32 bilinear, same association as aten
TOL_WARP = 1e-3           # absolute, images in [0,1]: the reference rounds its sampling grid through
                          # linspace(-1,1,W) + flow/((W-1)/2) in fp32 (~1e-4 px at W = 1920); ours samples
                          # at the absolute coordinate j + dx


def test_motion_warp_matches_reference_fixtures(lib, golden_dir):
    """rdvc_motion_warp vs the outputs of the reference's own resize_flow + WarpingLayer."""
    from oracle import motion_warp as mw
    g = np.load(os.path.join(golden_dir, "motion_warp.npz"))
    names = sorted(k[: -len("_shape")] for k in g.files if k.endswith("_shape"))
    for name in names:
        B, C, H, W, h_in, w_in = [int(v) for v in g[f"{name}_shape"]]
        img, flow = mw.synth_case(B, C, H, W, h_in, w_in, float(g[f"{name}_sigma"]), seed=len(name))
        n0 = lib.rdvc_corr_launch_count()
        warped, f = rc.motion_warp(gpu(img), gpu(flow), (H, W))
        assert lib.rdvc_corr_launch_count() - n0 == 1                  # one launch for both steps
        scale = max(1.0, float(np.abs(g[f"{name}_flow_frame"]).max()))
        assert np.abs(f.cpu().numpy() - g[f"{name}_flow_frame"]).max() / scale < TOL_FLOW_RESIZE, name
        assert np.abs(warped.cpu().numpy() - g[f"{name}_warped"]).max() < TOL_WARP, name
        # the two reference-named entry points give the same numbers as the fused call
        f2 = rc.resize_flow(gpu(flow), (H, W))
        assert torch.equal(f2, f) if (h_in, w_in) != (H, W) else f2.data_ptr() != 0
        assert torch.equal(rc.WarpingLayer()(gpu(img), f), warped)


def test_motion_warp_vs_torch_ops_odd_shapes(lib):
    """Against the same torch ops the reference composes (interpolate + grid_sample), on the GPU, for
    shapes the fixtures do not cover (up- and down-scaling, C = 1..4, B = 3)."""
    import torch.nn.functional as F
    for (B, C, H, W, h_in, w_in) in [(3, 1, 31, 45, 17, 23), (1, 4, 64, 96, 72, 104), (2, 3, 50, 50, 25, 100)]:
        gen = torch.Generator(device="cuda").manual_seed(H * W)
        x = torch.rand(B, C, H, W, device="cuda", generator=gen)
        fl = 3.0 * torch.randn(B, 2, h_in, w_in, device="cuda", generator=gen)
        warped, f = rc.motion_warp(x, fl, (H, W))
        ref_f = F.interpolate(fl, size=(H, W), mode="bilinear", align_corners=False, antialias=False)
        ref_f = ref_f * torch.tensor([W / w_in, H / h_in], device="cuda").view(1, 2, 1, 1)
        assert ((f - ref_f).abs().max() / ref_f.abs().max().clamp(min=1)).item() < TOL_FLOW_RESIZE
        gy, gx = torch.meshgrid(torch.linspace(-1, 1, H, device="cuda"), torch.linspace(-1, 1, W, device="cuda"), indexing="ij")
        grid = torch.stack((gx, gy), 2)[None] + torch.stack((ref_f[:, 0] / ((W - 1) / 2), ref_f[:, 1] / ((H - 1) / 2)), 3)
        ref_w = F.grid_sample(x, grid, mode="bilinear", padding_mode="border", align_corners=True)
        assert (warped - ref_w).abs().max().item() < TOL_WARP


def test_motion_warp_1080p_properties(lib):
    """Frame 1920x1080, RAFT flow at 1920x1088 (R:codec_processing.py:1446): size-independent properties."""
    B, C, H, W, h_in, w_in = 1, 3, 1080, 1920, 1088, 1920
    gen = torch.Generator(device="cuda").manual_seed(7)
    x = torch.rand(B, C, H, W, device="cuda", generator=gen)
    zero = torch.zeros(B, 2, h_in, w_in, device="cuda")
    warped, f = rc.motion_warp(x, zero, (H, W))
    assert torch.equal(warped, x) and not f.any()                       # zero flow: identity, bit exact
    shift = zero.clone(); shift[:, 0] = 5.0; shift[:, 1] = -3.0 * h_in / H   # dy scales by H / h_in -> -3 px
    warped, f = rc.motion_warp(x, shift, (H, W))
    assert torch.allclose(f[:, 0], torch.full_like(f[:, 0], 5.0)) and torch.allclose(f[:, 1], torch.full_like(f[:, 1], -3.0), atol=1e-5)
    assert torch.allclose(warped[:, :, 3:, : W - 5], x[:, :, : H - 3, 5:], atol=2e-5)   # integer shift = copy
    assert torch.allclose(warped[:, :, 3:, W - 5:], x[:, :, : H - 3, W - 1:].expand(-1, -1, -1, 5), atol=2e-5)  # border
    lin = torch.arange(w_in, device="cuda", dtype=torch.float32).view(1, 1, 1, w_in).expand(B, 1, h_in, w_in)
    fl = torch.cat([lin * 0.001, lin * 0.0], 1).contiguous()             # resize is linear: exact on a ramp
    f = rc.resize_flow(fl, (H, W))
    assert torch.allclose(f[:, 0], (lin[:, 0, :H] * 0.001), atol=1e-5) and f.shape == (B, 2, H, W)


# ------------------------------------------------------------------ next row: frame preparation
TOL_PREP = 2e-5           # absolute on [0, 1]: fp32 triangle-filter sums in one pass instead of aten's two
                          # (observed up to 9e-6 at 1080 -> 1088)


def test_preprocess_matches_reference_fixtures(lib, golden_dir):
    """rdvc_preprocess_frame vs the outputs of the reference's own preprocess_frame_raft / _codec."""
    from oracle import preprocess as pp
    g = np.load(os.path.join(golden_dir, "preprocess.npz"))
    for name in sorted(k[: -len("_shape")] for k in g.files if k.endswith("_shape")):
        H, W, C, h, w = [int(v) for v in g[f"{name}_shape"]]
        frame = pp.synth_frame(H, W, C, seed=len(name))
        fr = frame if C > 1 else frame[:, :, 0]
        got = rc.preprocess_frame_raft(fr, (h, w), torch.device("cuda"))
        assert got.shape == (1, C, h, w) and got.dtype == torch.float32
        assert np.abs(got.cpu().numpy() - g[f"{name}_raft"]).max() < TOL_PREP, name
        assert np.abs(rc.preprocess_frame_codec(fr, "cuda").cpu().numpy() - g[f"{name}_codec"]).max() < 1e-7, name


def test_preprocess_1080p_vs_torchvision(lib):
    """Full size against the reference's ops run on the GPU: 1080p -> 1088x1920 (what the 1080p runs need) and
    1080p -> 368x640 (the reference's default RAFT size, 2.9x down with anti-aliasing)."""
    import torchvision.transforms.functional as TF
    from oracle import preprocess as pp
    frame = pp.synth_frame(1080, 1920, 3, seed=3)
    t = torch.from_numpy(frame).cuda().permute(2, 0, 1).float() / 255.0
    for size in ((1088, 1920), (368, 640), (1080, 1920)):
        n0 = lib.rdvc_corr_launch_count()
        got = rc.preprocess_frame_raft(frame, size, "cuda")
        assert lib.rdvc_corr_launch_count() - n0 == 1
        ref = TF.resize(t, list(size), antialias=True).unsqueeze(0)
        assert (got - ref).abs().max().item() < TOL_PREP, size
    assert torch.equal(rc.preprocess_frame_codec(frame, "cuda")[0], (t * 255.0).round() / 255.0) or \
        (rc.preprocess_frame_codec(frame, "cuda")[0] - t).abs().max().item() < 1e-7


# ------------------------------------------------------------------ randomized sweep over the parameter space
def test_random_configurations_against_oracle(lib):
    """24 seeded random configurations (batch, channels, odd sizes, levels, radius, storage type, layout):
    build vs the fp64 statement of what the kernel multiplies, lookup vs the C oracle on the library's pyramid."""
    rng = np.random.default_rng(2026)
    for case in range(24):
        levels = int(rng.integers(1, 5))
        lo = 2 * 2 ** (levels - 1)
        B, D = int(rng.integers(1, 3)), int(rng.choice([64, 128, 192, 256]))
        h, w = int(rng.integers(lo, 41)), int(rng.integers(lo, 49))
        radius = int(rng.integers(1, 5))
        vol = torch.float32 if rng.random() < 0.5 else torch.bfloat16
        layout = ROW if rng.random() < 0.4 else TILED
        sigma = float(rng.choice([0.0, 0.5, 3.0, 30.0]))
        tag = (case, B, D, h, w, levels, radius, str(vol), layout, sigma)
        f1, f2 = cn.synth_fmaps(B, D, h, w, seed=100 + case)
        pyr = rc.build_pyramid(gpu(f1), gpu(f2), levels, vol, layout=layout)
        ref = ref_pyramid_linear(f1, f2, levels)
        own = []
        for l in range(levels):
            got = pyr.level(l)[:, 0].float().cpu().numpy()
            tol = TOL_SAME_OPERANDS_BF16 if vol == torch.bfloat16 else (TOL_SAME_OPERANDS_F32 if l == 0 else TOL_SAME_OPERANDS_POOLED)
            assert rel_max(got, ref[l]) < tol, tag + (l,)
            own.append(got)
        co = cn.synth_coords(B, h, w, sigma, seed=case)
        got = rc.index_pyramid(pyr, gpu(co), radius).cpu().numpy()
        flat = np.concatenate([x.reshape(-1) for x in own])
        assert got.shape == (B, levels * (2 * radius + 1) ** 2, h, w)
        assert rel_max(got, cc.index_pyramid(flat, co, levels, radius)) < TOL_LOOKUP, tag
