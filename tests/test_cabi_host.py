"""CPU tests (-m "not gpu"): the C-ABI library loads and exports every symbol the header declares,
its size arithmetic and argument validation (no compute), and the Python host mirror's
interface/error behaviour (same surface as TV:raft.py:337-431)."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest
import torch

import rdvc_corr_b200 as rc
from oracle import corr_numpy as cn

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
HEADER = os.path.join(ROOT, "include", "rdvc_corr.h")


@pytest.fixture(scope="module")
def lib():
    rc._build.build()  # nvcc cross-compiles without a GPU
    return rc._cabi.load()


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rdvc_(?:corr|motion|preprocess|mcn)_\w+)\s*\(", src)))


def test_header_symbols_exported(lib):
    declared = header_functions()
    assert len(declared) >= 10
    out = subprocess.check_output(["nm", "-D", "--defined-only", rc._cabi.lib_path()], text=True)
    exported = set(re.findall(r"\bT (rdvc_(?:corr|motion|preprocess|mcn)_\w+)", out))
    missing = [f for f in declared if f not in exported]
    assert not missing, f"header declares symbols the library does not export: {missing}"
    # and the ctypes table binds exactly the header's functions
    assert sorted(rc._cabi.SYMBOLS) == declared
    for name in declared:
        assert hasattr(lib, name)


def test_header_is_plain_c_and_example_links(lib, tmp_path):
    """include/rdvc_corr.h is a C header (no C++ in the signatures): the plain-C host example compiles with
    -std=c99 -Wall -Werror and links against the shared library (not executed here: no GPU)."""
    src = os.path.join(ROOT, "examples", "host_pair.c")
    obj = str(tmp_path / "host_pair.o")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", src, "-o", obj])
    libdir = os.path.dirname(rc._cabi.lib_path())
    exe = str(tmp_path / "host_pair")
    subprocess.check_call(["gcc", obj, "-o", exe, "-L", libdir, "-lrdvc_corr", "-lm", "-Wl,-rpath," + libdir,
                           "-Wl,--allow-shlib-undefined"])
    assert os.path.exists(exe)


def test_library_is_sm100a_with_tcgen05_and_tma():
    """The shipped kernels are Blackwell-native: tcgen05.mma (UTCHMMA), TMEM loads (LDTM) and TMA
    loads/stores (UTMALDG/UTMASTG) must be in the SASS."""
    out = subprocess.run(["cuobjdump", "-sass", rc._cabi.lib_path()], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    sass = out.stdout
    assert "sm_100a" in sass
    for mnemonic in ("UTCHMMA", "LDTM", "UTMALDG", "UTMASTG"):
        assert mnemonic in sass, mnemonic


def test_version_and_sizes(lib):
    assert lib.rdvc_corr_version() == 102
    F32, BF16 = rc.RDVC_DT_F32, rc.RDVC_DT_BF16
    ROW, TILED = rc.RDVC_LAYOUT_ROWMAJOR, rc.RDVC_LAYOUT_TILED
    # 1080p: 136x240 -> 32640 query pixels, levels 136x240, 68x120, 34x60, 17x30
    assert lib.rdvc_corr_pyramid_bytes(1, 136, 240, 4, F32, ROW) == 5_659_776_000
    assert lib.rdvc_corr_pyramid_bytes(1, 136, 240, 4, BF16, ROW) == 2_829_888_000
    assert lib.rdvc_corr_level_offset_bytes(1, 136, 240, 0, F32, ROW) == 0
    assert lib.rdvc_corr_level_offset_bytes(1, 136, 240, 1, F32, ROW) == 32640 * 32640 * 4
    assert rc.corr_block.tile_shape(torch.float32) == (4, 4)      # 16-byte rows x 4 rows
    assert rc.corr_block.tile_shape(torch.bfloat16) == (8, 4)
    # tiled 1080p fp32: level images are whole tiles rounded up to 256 bytes: 68x120 -> +32 elements,
    # 34x60 -> 36x60 + 16, 17x30 -> 20x32 (+0.6 % bytes)
    n = 32640
    assert lib.rdvc_corr_pyramid_bytes(1, 136, 240, 4, F32, TILED) == 4 * n * (136 * 240 + 68 * 120 + 32 + 36 * 60 + 16 + 20 * 32)
    assert lib.rdvc_corr_level_image_elems(136, 240, 2, F32, TILED) == 36 * 60 + 16
    assert lib.rdvc_corr_level_image_elems(136, 240, 2, F32, ROW) == 34 * 60
    for (B, h, w) in [(2, 18, 22), (1, 46, 80), (3, 33, 47)]:
        for vd, es, tw in ((F32, 4, 4), (BF16, 2, 8)):
            for layout in (ROW, TILED):
                off = 0
                for l in range(4):
                    assert lib.rdvc_corr_level_offset_bytes(B, h, w, l, vd, layout) == off
                    assert off % 256 == 0
                    hl, wl = h >> l, w >> l
                    img = hl * wl
                    if layout == TILED:
                        img = (-(-hl // 4) * 4) * (-(-wl // tw) * tw)
                        img = -(-img * es // 256) * 256 // es          # whole 256-byte units
                    assert lib.rdvc_corr_level_image_elems(h, w, l, vd, layout) == img
                    n = B * h * w * img * es
                    off += (n + 255) // 256 * 256
                assert lib.rdvc_corr_pyramid_bytes(B, h, w, 4, vd, layout) == off
    ws = lib.rdvc_corr_workspace_bytes(1, 256, 136, 240)
    assert ws >= 2 * 32640 * 256 * 2 and ws % 256 == 0


def test_argument_validation_returns_negative_codes(lib):
    """Bad arguments are rejected before any CUDA call (works without a GPU)."""
    buf = ctypes.create_string_buffer(4096)
    p = ctypes.cast(buf, ctypes.c_void_p)
    F32 = rc.RDVC_DT_F32
    big = 1 << 40

    def build(**kw):
        a = dict(f1=p, f2=p, B=1, D=256, h=46, w=80, idt=F32, pyr=p, vdt=F32, lay=1, L=4, ws=p, wsb=big, st=None)
        a.update(kw)
        return lib.rdvc_corr_build(a["f1"], a["f2"], a["B"], a["D"], a["h"], a["w"], a["idt"], a["pyr"],
                                   a["vdt"], a["lay"], a["L"], a["ws"], a["wsb"], a["st"])

    assert build(f1=None) == -1 and "null" in rc._cabi.last_error()
    assert build(B=0) == -2
    assert build(h=15) == -3 and "too small" in rc._cabi.last_error()   # TV:raft.py:376
    assert build(h=7, L=3) == -3
    assert build(idt=7) == -4
    assert build(vdt=rc.RDVC_DT_F16) == -4
    assert build(D=100) == -5
    assert build(D=512) == -5
    assert build(L=5) == -5
    assert build(lay=7) == -5
    assert build(wsb=16) == -6
    misaligned = ctypes.c_void_p(p.value + 8)
    assert build(pyr=misaligned) == -7
    assert lib.rdvc_corr_lookup(None, F32, 1, p, 1, 46, 80, 4, 4, p, None) == -1
    assert lib.rdvc_corr_lookup(p, F32, 1, p, 1, 46, 80, 4, 9, p, None) == -5
    assert lib.rdvc_corr_lookup(p, F32, 0, p, 1, 46, 80, 4, 9, p, None) == -5
    assert lib.rdvc_corr_lookup(p, F32, 3, p, 1, 46, 80, 4, 4, p, None) == -5
    assert lib.rdvc_corr_lookup(p, F32, 1, p, 1, 8, 80, 4, 4, p, None) == -3
    assert lib.rdvc_corr_set_option(99, 0) == -5
    assert lib.rdvc_corr_pair_host(None, p, p, p, 1, 256, 46, 80, 4, 4, 12, F32) == -1
    assert lib.rdvc_corr_pair_host(p, p, p, p, 1, 256, 46, 80, 4, 4, 0, F32) == -2
    assert lib.rdvc_corr_pair_host_submit(p, p, p, None, 1, 256, 46, 80, 4, 4, 12, F32, 0) == -1
    assert lib.rdvc_corr_pair_host_submit(p, p, p, p, 1, 256, 46, 80, 4, 4, 12, F32, 2) == -5
    assert lib.rdvc_corr_pair_host_wait(0) == 0 and lib.rdvc_corr_pair_host_wait(-1) == -5
    with pytest.raises(ValueError, match="RDVC_E_TOO_SMALL"):
        rc._cabi.check(build(h=15), "rdvc_corr_build")
    # rdvc_motion_warp(prev, flow, B, C, H, W, h_in, w_in, warped, flow_out, stream)
    assert lib.rdvc_motion_warp(p, None, 1, 3, 8, 8, 8, 8, p, p, None) == -1
    assert lib.rdvc_motion_warp(p, p, 1, 3, 8, 8, 8, 8, None, p, None) == -1      # prev without warped
    assert lib.rdvc_motion_warp(None, p, 1, 0, 8, 8, 8, 8, None, None, None) == -1  # nothing to do
    assert lib.rdvc_motion_warp(p, p, 1, 3, 0, 8, 8, 8, p, p, None) == -2
    assert lib.rdvc_motion_warp(p, p, 1, 3, 70000, 8, 8, 8, p, p, None) == -5
    # rdvc_preprocess_frame(frame_hwc, H, W, C, out, h_out, w_out, stream)
    assert lib.rdvc_preprocess_frame(None, 8, 8, 3, p, 8, 8, None) == -1
    assert lib.rdvc_preprocess_frame(p, 8, 0, 3, p, 8, 8, None) == -2
    assert lib.rdvc_preprocess_frame(p, 8, 8, 5, p, 8, 8, None) == -5
    assert lib.rdvc_preprocess_frame(p, 800, 8, 3, p, 8, 8, None) == -5      # 100x down-scaling


# ------------------------------------------------------------------ Python host mirror
def test_tv_surface_matches_torchvision():
    from torchvision.models.optical_flow.raft import CorrBlock as TVRef
    ours, ref = rc.TVCorrBlock(num_levels=4, radius=4), TVRef(num_levels=4, radius=4)
    assert ours.out_channels == ref.out_channels == 324
    assert ours.num_levels == ref.num_levels and ours.radius == ref.radius
    for name in ("build_pyramid", "index_pyramid"):
        assert callable(getattr(ours, name))
    assert rc.TVCorrBlock(num_levels=3, radius=3).out_channels == 3 * 49


def test_value_errors_match_reference_messages():
    blk = rc.TVCorrBlock()
    from torchvision.models.optical_flow.raft import CorrBlock as TVRef
    ref = TVRef()
    a, b = torch.zeros(1, 256, 46, 80), torch.zeros(1, 256, 46, 81)
    with pytest.raises(ValueError) as e1:
        blk.build_pyramid(a, b)
    with pytest.raises(ValueError) as e2:
        ref.build_pyramid(a, b)
    assert str(e1.value) == str(e2.value)
    small = torch.zeros(1, 256, 8, 80)
    with pytest.raises(ValueError) as e1:
        blk.build_pyramid(small, small)
    with pytest.raises(ValueError) as e2:
        ref.build_pyramid(small, small)
    assert str(e1.value) == str(e2.value)
    with pytest.raises(RuntimeError, match="before build_pyramid"):
        rc.TVCorrBlock().index_pyramid(torch.zeros(1, 2, 46, 80))


def test_no_cpu_fallback():
    """CPU tensors must fail loudly: the product has no CPU or PyTorch fallback path."""
    f = torch.zeros(1, 256, 46, 80)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        rc.TVCorrBlock().build_pyramid(f, f)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        rc.CorrBlock(f, f, num_levels=4, radius=4)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(rc._cabi, "_lib", None)
    monkeypatch.setattr(rc._build, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="has not been built"):
        rc._cabi.load()


def test_product_never_imports_oracle():
    """The oracle is test infrastructure; nothing under the product package may reference it."""
    pkg = os.path.dirname(rc.__file__)
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(root, f)).read()
                assert "oracle" not in text.lower().replace("no oracle", ""), os.path.join(root, f)


def test_injects_into_torchvision_raft():
    """raft_large(corr_block=...) keeps our block (TV:raft.py:792) and sizes convcorr1 from
    out_channels (TV:raft.py:797)."""
    from torchvision.models.optical_flow import raft_large
    blk = rc.TVCorrBlock()
    model = raft_large(weights=None, corr_block=blk)
    assert model.corr_block is blk
    assert model.update_block.motion_encoder.convcorr1[0].in_channels == 324


def test_motion_warp_host_behaviour_matches_reference():
    """resize_flow / WarpingLayer keep the reference's early-outs and messages
    (R:codec_processing.py:782-797 and :336-340); compute needs the GPU and fails loudly without it."""
    assert rc.resize_flow(None, (4, 4)) is None
    f = torch.zeros(1, 2, 6, 8)
    assert rc.resize_flow(f, (6, 8)) is f                                   # :788-789: no resize needed
    with pytest.raises(ValueError, match="Flow tensor must have 2 channels, got 3"):
        rc.resize_flow(torch.zeros(1, 3, 6, 8), (4, 4))
    z = rc.resize_flow(torch.zeros(1, 2, 0, 8), (4, 5))
    assert z.shape == (1, 2, 4, 5) and not z.any()
    assert rc.resize_flow(f, (0, 5)).shape == (1, 2, 0, 5)
    x = torch.zeros(2, 3, 6, 8)
    with pytest.raises(ValueError) as e:
        rc.WarpingLayer()(x, torch.zeros(2, 2, 6, 9))
    assert str(e.value) == f"Input image (2,3,6,8) and flow ({torch.Size([2, 2, 6, 9])}) shape/channel mismatch."
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        rc.resize_flow(f, (12, 16))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        rc.WarpingLayer()(x, torch.zeros(2, 2, 6, 8))
    with pytest.raises(ValueError, match="does not match"):
        rc.motion_warp(x, torch.zeros(2, 2, 3, 4), (6, 9))


def test_preprocess_host_behaviour_matches_reference(capsys):
    """preprocess_frame_raft / _codec keep the reference's contract: a failure is printed and None is returned
    (R:codec_processing.py:760-761, :768-769) -- here: no GPU, and there is no CPU fallback to hide it."""
    frame = np.zeros((16, 24, 3), np.uint8)
    assert rc.preprocess_frame_raft(frame, (32, 48), torch.device("cpu")) is None
    assert "Error preprocessing frame for RAFT" in capsys.readouterr().out
    assert rc.preprocess_frame_codec(frame, "cpu") is None
    assert "Error preprocessing frame for Codec" in capsys.readouterr().out
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        rc.frame_to_tensor(frame, None, "cpu")


def test_coords_validation_without_gpu():
    pyr = rc.CorrPyramid(1, 46, 80, 4, torch.float32, torch.empty(0, dtype=torch.uint8))
    with pytest.raises(ValueError, match="coords should be"):
        rc.index_pyramid(pyr, torch.zeros(1, 3, 46, 80))
    with pytest.raises(ValueError, match="do not match"):
        rc.index_pyramid(pyr, torch.zeros(1, 2, 46, 81))
    with pytest.raises(RuntimeError, match="GPU only"):
        rc.index_pyramid(pyr, torch.zeros(1, 2, 46, 80))
