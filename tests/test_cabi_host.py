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
    return sorted(set(re.findall(r"\b(rdvc_(?:corr|motion|preprocess|mcn|conv1x1|ec|linear)\w*)\s*\(", src)))


def test_header_symbols_exported(lib):
    declared = header_functions()
    assert len(declared) >= 10
    out = subprocess.check_output(["nm", "-D", "--defined-only", rc._cabi.lib_path()], text=True)
    exported = set(re.findall(r"\bT (rdvc_(?:corr|motion|preprocess|mcn|conv1x1|ec|linear)\w*)", out))
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
    assert lib.rdvc_corr_version() == 103
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


# ------------------------------------------------------------------ round 2: validation, build identity, f-1 host side
def test_bad_volume_dtype_is_an_error_not_a_crash(lib):
    """A bad vol_dtype enum used to divide by zero inside the size queries (SIGILL with the TILED layout)."""
    for layout in (rc.RDVC_LAYOUT_ROWMAJOR, rc.RDVC_LAYOUT_TILED):
        for bad in (7, -1, rc.RDVC_DT_F16):
            assert lib.rdvc_corr_pyramid_bytes(1, 46, 80, 4, bad, layout) == 0
            assert lib.rdvc_corr_level_offset_bytes(1, 46, 80, 2, bad, layout) == 0
            assert lib.rdvc_corr_level_image_elems(46, 80, 0, bad, layout) == 0
    buf = np.zeros(16, np.float32)
    fp = buf.ctypes.data_as(ctypes.c_void_p)
    # validated before any size arithmetic, allocation or CUDA call: works without a GPU
    assert lib.rdvc_corr_pair_host_submit(fp, fp, fp, fp, 1, 64, 16, 16, 4, 4, 1, 7, 0) == -4
    assert b"vol_dtype" in lib.rdvc_corr_last_error()
    assert lib.rdvc_corr_pair_host_submit_ex(fp, fp, fp, fp, 1, 64, 16, 16, 4, 4, 1, rc.RDVC_DT_F32, rc.RDVC_DT_BF16, 0) == -4
    assert lib.rdvc_corr_pair_host(fp, fp, fp, fp, 1, 64, 16, 16, 4, 4, 1, 7) == -4


def test_library_is_built_from_this_tree(lib):
    """The content hash of csrc/ + include/ is compiled into the library; the loader refuses a mismatch."""
    info = lib.rdvc_corr_build_info().decode()
    assert f"RDVC_SRC_HASH={rc._build.source_hash()}" in info and "version=103" in info
    assert rc._build.embedded_hash(rc._cabi.lib_path()) == rc._build.source_hash()
    assert not rc._build.is_stale(rc._cabi.lib_path())
    if os.environ.get("RDVC_CORR_LIB") is None:
        assert "experiments=0" in info and not rc._cabi.has_experiments()


def test_stale_library_is_refused(tmp_path, monkeypatch):
    fake = tmp_path / "librdvc_corr.so"
    blob = open(rc._build.LIB_PATH, "rb").read().replace(rc._build.source_hash().encode(), b"0" * 64)
    fake.write_bytes(blob)
    monkeypatch.setenv("RDVC_CORR_LIB", str(fake))
    monkeypatch.setattr(rc._cabi, "_lib", None)
    with pytest.raises(RuntimeError, match="built from other sources"):
        rc._cabi.load()


def test_product_build_has_no_work_skipping_knobs(lib):
    if rc._cabi.has_experiments():
        pytest.skip("experiments build")
    for key, value in [(0, 3), (0, 4), (3, 7), (4, 1), (6, 1), (12, 2), (12, 3)]:
        assert lib.rdvc_corr_set_option(key, value) == -5, (key, value)
    for key, value in [(0, 0), (3, 15), (4, 2), (4, 0), (6, 0), (12, 0), (12, 1)]:
        assert lib.rdvc_corr_set_option(key, value) == 0, (key, value)
    syms = subprocess.check_output(["nm", "-C", rc._cabi.lib_path()], text=True)
    assert "corr_build2_kernel" not in syms                       # the CTA-pair build variant
    assert "corr_build_kernel<0," not in syms                     # MODE_FUSED
    assert "corr_conv1x1_kernel<float>" in syms and "corr_lookup_tiled_kernel<4, float, 2, 0>" in syms


def test_conv1x1_weight_packing(lib):
    """rdvc_conv1x1_pack_weights (pure host C): torchvision channel l*S*S + i*S + j lands in column l*PL + j*S + i
    of a row padded to whole 64-element k-blocks; everything else is zero."""
    assert lib.rdvc_corr_feat_pitch(4, 4) == 352 and lib.rdvc_corr_feat_pitch(4, 3) == 224
    assert lib.rdvc_corr_feat_pitch(3, 4) == 272 and lib.rdvc_corr_feat_pitch(5, 4) == 0
    for (L, r, cout, fd) in [(4, 4, 256, rc.RDVC_DT_BF16), (4, 3, 96, rc.RDVC_DT_F16), (2, 2, 32, rc.RDVC_DT_BF16)]:
        S = 2 * r + 1
        PL = (S * S + 7) // 8 * 8
        kp = lib.rdvc_corr_feat_pitch(L, r)
        kpw = (kp + 63) // 64 * 64
        assert lib.rdvc_conv1x1_packed_weight_bytes(cout, L, r) == cout * kpw * 2
        rng = np.random.default_rng(cout)
        wgt = rng.standard_normal((cout, L * S * S)).astype(np.float32)
        packed = np.full(cout * kpw, 0x7FFF, np.uint16)
        rc._cabi.check(lib.rdvc_conv1x1_pack_weights(wgt.ctypes.data_as(ctypes.c_void_p), cout, L, r, fd,
                                                     packed.ctypes.data_as(ctypes.c_void_p)), "pack")
        t = torch.from_numpy(packed.view(np.int16).reshape(cout, kpw))
        vals = t.view(torch.bfloat16 if fd == rc.RDVC_DT_BF16 else torch.float16).float().numpy()
        want = np.zeros((cout, kpw), np.float32)
        rnd = torch.bfloat16 if fd == rc.RDVC_DT_BF16 else torch.float16
        for l in range(L):
            for i in range(S):
                for j in range(S):
                    want[:, l * PL + j * S + i] = torch.from_numpy(wgt[:, l * S * S + i * S + j]).to(rnd).float().numpy()
        assert np.array_equal(vals, want)
    w = np.zeros((48, 324), np.float32)
    assert lib.rdvc_conv1x1_pack_weights(w.ctypes.data_as(ctypes.c_void_p), 48, 4, 4, rc.RDVC_DT_BF16,
                                         w.ctypes.data_as(ctypes.c_void_p)) == -5          # cout % 32 != 0
    assert lib.rdvc_conv1x1_packed_weight_bytes(48, 4, 4) == 0


def test_encoder_tail_host_side(lib):
    """rdvc_linear_pack_weights (host) and the argument validation of the row f-2 entry points (no GPU needed)."""
    rng = np.random.default_rng(5)
    wgt = rng.standard_normal((256, 128)).astype(np.float32)
    for dt, tdt in ((rc.RDVC_DT_BF16, torch.bfloat16), (rc.RDVC_DT_F16, torch.float16)):
        assert lib.rdvc_linear_packed_weight_bytes(256, 128) == 256 * 128 * 2
        packed = np.zeros(256 * 128, np.uint16)
        rc._cabi.check(lib.rdvc_linear_pack_weights(wgt.ctypes.data_as(ctypes.c_void_p), 256, 128, dt,
                                                    packed.ctypes.data_as(ctypes.c_void_p)), "pack")
        got = torch.from_numpy(packed.view(np.int16)).view(tdt).float().numpy().reshape(256, 128)
        assert np.array_equal(got, torch.from_numpy(wgt).to(tdt).float().numpy())
    assert lib.rdvc_linear_pack_weights(wgt.ctypes.data_as(ctypes.c_void_p), 256, 128, rc.RDVC_DT_F32,
                                        wgt.ctypes.data_as(ctypes.c_void_p)) == -4
    one = ctypes.c_void_p(256)          # a non-null, 256-byte aligned dummy: validation runs before any CUDA call
    big = 1 << 40
    assert lib.rdvc_corr_pack(None, one, 1, 128, 46, 80, rc.RDVC_DT_F32, rc.RDVC_DT_F32, rc.RDVC_LAYOUT_TILED, 4, one, big, None) == -1
    assert lib.rdvc_corr_pack(one, one, 1, 96, 46, 80, rc.RDVC_DT_F32, rc.RDVC_DT_F32, rc.RDVC_LAYOUT_TILED, 4, one, big, None) == -5
    assert lib.rdvc_corr_pack(one, one, 1, 128, 46, 80, rc.RDVC_DT_F32, rc.RDVC_DT_F32, rc.RDVC_LAYOUT_TILED, 4, one, 16, None) == -6
    assert lib.rdvc_corr_encoder_tail(one, big, 64, one, None, 256, 1, 46, 80, rc.RDVC_DT_BF16, rc.RDVC_DT_F32,
                                      rc.RDVC_LAYOUT_TILED, 4, one, big, None) == -5          # only 128 -> 256
    assert lib.rdvc_corr_encoder_tail(one, big, 128, one, None, 256, 1, 46, 80, rc.RDVC_DT_F32, rc.RDVC_DT_F32,
                                      rc.RDVC_LAYOUT_TILED, 4, one, big, None) == -4          # operands are 16-bit
    assert lib.rdvc_corr_encoder_tail(one, 16, 128, one, None, 256, 1, 46, 80, rc.RDVC_DT_BF16, rc.RDVC_DT_F32,
                                      rc.RDVC_LAYOUT_TILED, 4, one, big, None) == -6
    assert lib.rdvc_corr_build_packed(1, 256, 46, 80, rc.RDVC_DT_F32, one, rc.RDVC_DT_F32, rc.RDVC_LAYOUT_TILED, 4, one, big, None) == -4
    assert lib.rdvc_corr_build_packed(1, 256, 8, 80, rc.RDVC_DT_BF16, one, rc.RDVC_DT_F32, rc.RDVC_LAYOUT_TILED, 4, one, big, None) == -3
    assert lib.rdvc_conv1x1(one, rc.RDVC_DT_F32, one, None, 1, 46, 80, 4, 4, 256, 1, one, rc.RDVC_DT_F32, None) == -4
    assert lib.rdvc_conv1x1(one, rc.RDVC_DT_BF16, one, None, 1, 46, 80, 4, 4, 48, 1, one, rc.RDVC_DT_F32, None) == -5
    assert lib.rdvc_corr_lookup_ex(one, rc.RDVC_DT_F32, rc.RDVC_LAYOUT_ROWMAJOR, one, 1, 46, 80, 4, 4, one, rc.RDVC_DT_F16, 0, None) == -5
    assert lib.rdvc_corr_lookup_ex(one, rc.RDVC_DT_F32, rc.RDVC_LAYOUT_TILED, one, 1, 46, 80, 4, 4, one, rc.RDVC_DT_BF16, 0, None) == -4
    assert lib.rdvc_corr_feat_rows(1, 17, 19) == 328 and lib.rdvc_corr_feat_bytes(1, 17, 19, 4, 4) == 352 * 328 * 2


def test_forward_only_guard():
    """Gradients cannot flow through the kernels: a tensor that requires grad is refused (unless grad mode is off,
    as in the reference's call, R:codec_processing.py:1436) instead of silently cutting the graph."""
    t = torch.zeros(2, requires_grad=True)
    with pytest.raises(RuntimeError, match="forward-only"):
        rc._cabi.forward_only("x", None, t)
    rc._cabi.forward_only("x", None, t.detach())
    with torch.no_grad():
        rc._cabi.forward_only("x", t)


def test_to_channels_last_keeps_values_and_is_idempotent():
    """raft_flow stores the stock update block / context encoder weights NHWC: values unchanged, done once (a second
    call must not re-allocate -- CUDA graphs captured in between hold the addresses)."""
    from rdvc_corr_b200.raft_flow import _to_channels_last
    m = torch.nn.Sequential(torch.nn.Conv2d(8, 16, 3), torch.nn.BatchNorm2d(16), torch.nn.Conv2d(16, 4, 1))
    before = {k: v.clone() for k, v in m.state_dict().items()}
    _to_channels_last(m)
    assert m[0].weight.is_contiguous(memory_format=torch.channels_last)
    assert all(torch.equal(v, before[k]) for k, v in m.state_dict().items())
    ptrs = [p.data_ptr() for p in m.parameters()]
    _to_channels_last(m)
    assert ptrs == [p.data_ptr() for p in m.parameters()]
    x = torch.randn(2, 8, 12, 12)
    ref = torch.nn.Sequential(torch.nn.Conv2d(8, 16, 3), torch.nn.BatchNorm2d(16), torch.nn.Conv2d(16, 4, 1))
    ref.load_state_dict(before)
    assert torch.allclose(m.eval()(x), ref.eval()(x), atol=1e-5)
