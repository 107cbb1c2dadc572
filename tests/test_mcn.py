"""Motion-compensation network ("next" row f-4, R:codec_processing.py:369-406).

CPU part (-m "not gpu"): the numpy oracle against the fixture produced by the reference's own classes, the host-side
weight packing (rdvc_mcn_pack_weights, pure C host code) evaluated the way the kernel walks it, the parameter tree /
BatchNorm folding of the host mirror, and its error behaviour.
GPU part (-m gpu): single layers and the whole network through the C ABI against the oracle.

Tolerances (absolute; frames and refinement maps live in [0, 1], activations are O(1)):
  * oracle vs the reference's fp32 output ......................... 1e-5   (observed 3e-7)
  * kernel layer vs the oracle with the SAME fp16 roundings ........ 4e-3 x max|ref|  (one fp16 ulp of the
    largest activation is 2^-11 = 5e-4 relative; a result on a rounding boundary may go either way)
  * network output vs the oracle with the same roundings ........... 1e-3
  * network output vs the reference's fp32 output .................. 3e-3  (fp16 operands, fp32 accumulation;
    the fp16-rounding oracle itself sits 2-3e-4 from the fp32 reference)
"""
import numpy as np
import pytest
import torch

import rdvc_corr_b200 as rc
from oracle import mcn as om
from rdvc_corr_b200 import mcn as hm

TOL_ORACLE = 1e-5
TOL_LAYER_REL = 4e-3
TOL_NET_SAME_ROUNDING = 1e-3
TOL_NET_FP32 = 3e-3
CASES = ("tile_exact", "ragged", "tiny")


@pytest.fixture(scope="module")
def golden(golden_dir):
    z = np.load(golden_dir + "/mcn.npz")
    params = {k[6:]: z[k] for k in z.files if k.startswith("state:")}
    return z, params


def mirror_with_reference_weights(params, device="cpu"):
    net = rc.MotionCompensationNetwork()
    missing = net.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in params.items()}, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    return net.eval().to(device)


# ------------------------------------------------------------------ CPU
@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_output(golden, name):
    z, params = golden
    out = om.mcn_forward(params, z[name + ":warped"], z[name + ":flow"], z[name + ":ref"])
    assert np.abs(out - z[name + ":out"]).max() <= TOL_ORACLE


def test_fp16_rounding_oracle_is_close_to_fp32(golden):
    z, params = golden
    for name in CASES:
        out = om.mcn_forward(params, z[name + ":warped"], z[name + ":flow"], z[name + ":ref"], emulate_fp16=True)
        assert np.abs(out - z[name + ":out"]).max() <= 1e-3


@pytest.mark.parametrize("cout,cin,k", [(32, 32, 3), (32, 8, 5), (3, 32, 5), (32, 32, 5), (8, 5, 5)])
def test_packed_weights_reproduce_the_convolution(cout, cin, k):
    rng = np.random.default_rng(cout * 100 + cin * 10 + k)
    w = rng.standard_normal((cout, cin, k, k)).astype(np.float32) * 0.1
    packed_u8, mask = hm.pack_conv_weights(torch.from_numpy(w))
    nout = 64 if cout > 8 else 16
    assert packed_u8.numel() == 3 * k * nout * 64 * 2
    packed = packed_u8.numpy().view(np.float16)
    x = np.zeros((2, 32, 9, 12), np.float32)
    x[:, :cin] = rng.standard_normal((2, cin, 9, 12)).astype(np.float16)
    got = om.superpixel_gemm_conv(x, packed, k, nout, mask)[:, :cout]
    want = om.conv2d_same(x[:, :cin], w.astype(np.float16).astype(np.float64))
    assert np.abs(got - want).max() <= 1e-9 * max(1.0, np.abs(want).max())
    # ... and the way the one-box kernel walks them (L / R accumulators shifted on the output side)
    got_x = om.superpixel_gemm_conv_xhalo(x, packed, k, nout, mask)[:, :cout]
    assert np.abs(got_x - want).max() <= 1e-9 * max(1.0, np.abs(want).max())
    # k-steps: 16-channel groups beyond cin are skipped, and for 3x3 the outer super-pixel columns need one pixel only
    groups = (cin + 15) // 16
    for t in range(3 * k):
        dsx = t % 3 - 1
        bits = (mask >> (4 * t)) & 15
        expect = 0
        for p in range(2):
            reach = [2 * dsx + p - q for q in range(2)]
            if any(abs(d) <= k // 2 for d in reach):
                expect |= ((1 << groups) - 1) << (2 * p)
        assert bits == expect, (t, bin(bits), bin(expect))


def test_pack_weights_rejects_unsupported():
    with pytest.raises(ValueError):
        hm.pack_conv_weights(torch.zeros(32, 32, 7, 7))
    with pytest.raises(ValueError):
        hm.pack_conv_weights(torch.zeros(32, 33, 3, 3))
    with pytest.raises(ValueError):
        hm.pack_conv_weights(torch.zeros(64, 32, 3, 3))


def test_mirror_takes_the_reference_state_dict_and_folds_like_the_oracle(golden):
    _, params = golden
    net = mirror_with_reference_weights(params)
    assert sorted(net.state_dict().keys()) == sorted(params.keys())
    for (w, b), (wo, bo) in zip(net.folded_layers(), om.folded_layers(params, 3)):
        assert np.abs(w.numpy() - wo).max() <= 1e-6 and np.abs(b.numpy() - bo).max() <= 1e-6


def test_mirror_error_behaviour(golden):
    _, params = golden
    net = mirror_with_reference_weights(params)
    a, f = torch.rand(1, 3, 8, 8), torch.rand(1, 2, 8, 8)
    with pytest.raises(ValueError, match="Input sizes mismatch"):          # R:codec_processing.py:395-398
        net(a, f, torch.rand(1, 3, 8, 9))
    with pytest.raises(ValueError, match="Expected flow shape"):           # :399-400
        net(a, torch.rand(1, 3, 8, 8), a)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        net(a, f, a)
    with pytest.raises(RuntimeError, match="inference only"):
        net.train()(a, f, a)
    with pytest.raises(ValueError):
        rc.MotionCompensationNetwork(base_channels=64)


def test_cabi_argument_errors_need_no_gpu():
    """Argument checking happens before any CUDA call: negative RDVC_E_* codes, message in rdvc_corr_last_error."""
    lib = rc._cabi.load()
    assert lib.rdvc_mcn_plane_bytes(1, 1080, 1920) == 1080 * 960 * 128
    assert lib.rdvc_mcn_plane_bytes(2, 5, 7) == 5120            # 2 * 5 * 4 super-pixels * 128 B (a multiple of 1 KB)
    assert lib.rdvc_mcn_plane_bytes(1, 3, 3) == 1024            # 3 * 2 * 128 = 768, rounded up to 1 KB
    assert lib.rdvc_mcn_workspace_bytes(1, 1080, 1920) == 3 * 1080 * 960 * 128
    assert lib.rdvc_mcn_packed_weight_bytes(3, 32) == 9 * 64 * 128 and lib.rdvc_mcn_packed_weight_bytes(5, 3) == 15 * 16 * 128
    assert lib.rdvc_mcn_packed_weight_bytes(7, 32) == 0 and lib.rdvc_mcn_packed_weight_bytes(3, 33) == 0
    A, Bp, Cp = 4096, 8192, 12288                                # fake 16-byte-aligned device addresses (never dereferenced)
    conv = lib.rdvc_mcn_conv
    assert conv(None, Bp, 0, None, 3, 1, None, Cp, 1, 8, 8, None) == -1                      # RDVC_E_NULL
    assert conv(A, Bp, 0, None, 4, 1, None, Cp, 1, 8, 8, None) == -5                         # kernel size
    assert conv(A, Bp, 0, None, 3, 7, None, Cp, 1, 8, 8, None) == -5                         # activation
    assert conv(A, Bp, 0, None, 3, 1, None, A, 1, 8, 8, None) == -5                          # in place
    assert "in place" in rc._cabi.last_error()
    assert conv(A, Bp, 0, None, 3, 1, None, Cp, 0, 8, 8, None) == -2                         # RDVC_E_SHAPE
    assert conv(A + 4, Bp, 0, None, 3, 1, None, Cp, 1, 8, 8, None) == -7                     # RDVC_E_ALIGN
    out = lib.rdvc_mcn_conv_out
    assert out(A, Bp, 0, None, 3, 3, A, Cp, 1, 8, 8, None) == -5                             # the output layer is 5x5
    assert out(A, Bp, 0, None, 5, 9, A, Cp, 1, 8, 8, None) == -5                             # at most 8 channels
    fwd = lib.rdvc_mcn_forward
    import ctypes
    ptrs = (ctypes.c_void_p * 8)(*([Bp] * 8))
    masks = (ctypes.c_ulonglong * 8)(*([0] * 8))
    biases = (ctypes.c_float * 256)()
    assert fwd(A, A, A, 1, 8, 8, 3, ptrs, masks, biases, Cp, 10, A, None) == -6              # RDVC_E_WORKSPACE
    assert fwd(A, A, A, 1, 8, 8, -1, ptrs, masks, biases, Cp, 1 << 20, A, None) == -5
    assert fwd(A, A, None, 1, 8, 8, 3, ptrs, masks, biases, Cp, 1 << 20, A, None) == -1
    assert lib.rdvc_mcn_pack_input(A, A, A, 1, 3, 3, 3, 8, 8, Cp, None) == -5                # 9 input channels


# ------------------------------------------------------------------ GPU
@pytest.fixture(params=["auto", "three_boxes", "x_halo"])
def mcn_kernel(request):
    """Both convolution kernels (option key 14): three activation boxes per tile / one box with the x halo, and the
    default that picks per layer."""
    lib = rc._cabi.load()
    assert lib.rdvc_corr_set_option(14, {"auto": 0, "three_boxes": 1, "x_halo": 2}[request.param]) == 0
    yield request.param
    lib.rdvc_corr_set_option(14, 0)


def _layer_case(B, H, W, k, cin, seed, act, with_res):
    rng = np.random.default_rng(seed)
    x = np.zeros((B, 32, H, W), np.float32)
    x[:, :cin] = rng.standard_normal((B, cin, H, W)).astype(np.float32)
    w = (rng.standard_normal((32, cin, k, k)) * (1.0 / np.sqrt(cin * k * k))).astype(np.float32)
    b = rng.standard_normal(32).astype(np.float32) * 0.3
    res = rng.standard_normal((B, 32, H, W)).astype(np.float32) if with_res else None
    return x, w, b, res


@pytest.mark.gpu
@pytest.mark.parametrize("B,H,W,k,cin,act,with_res", [
    (1, 8, 32, 3, 32, True, False),      # exactly one tile
    (1, 16, 64, 3, 32, False, True),     # 2 x 2 tiles, residual, no activation
    (2, 21, 45, 3, 32, True, True),      # odd width, ragged tiles, B > 1
    (1, 19, 70, 5, 32, True, False),     # 5 x 5
    (1, 13, 38, 5, 8, True, False),      # the network's first layer shape (8 real channels)
    (1, 5, 7, 3, 32, True, True),        # smaller than a tile
    (1, 70, 330, 3, 32, True, True),     # more tiles than SMs: the persistent loop and both accumulators wrap
])
def test_conv_layer_matches_oracle(mcn_kernel, B, H, W, k, cin, act, with_res):
    x, w, b, res = _layer_case(B, H, W, k, cin, seed=H * 1000 + W + k, act=act, with_res=with_res)
    packed, mask = hm.pack_conv_weights(torch.from_numpy(w))
    plane = hm.plane_from_nchw(torch.from_numpy(x).cuda())
    rplane = None if res is None else hm.plane_from_nchw(torch.from_numpy(res).cuda())
    out = hm.conv_layer(plane, packed.cuda(), mask, torch.from_numpy(b), k, hm.ACT_LEAKY if act else hm.ACT_NONE,
                        B, H, W, residual=rplane)
    torch.cuda.synchronize()
    got = hm.plane_to_nchw(out, B, H, W).cpu().numpy()
    want = om.conv_layer(x[:, :cin], w, b, act, residual=res, emulate_fp16=True)
    assert np.abs(got - want).max() <= TOL_LAYER_REL * np.abs(want).max()
    if W % 2:   # the padding pixel of an odd-width row must hold zeros (it is the next layer's zero padding)
        full = out[:B * H * (W + 1) * 64].view(torch.float16).reshape(B, H, W + 1, 32)
        assert float(full[:, :, W].abs().max()) == 0.0


@pytest.mark.gpu
def test_reverse_tile_order_is_bit_identical(mcn_kernel):
    """RDVC_MCN_REVERSE_ORDER only changes the order in which a launch walks its tiles (L2 reuse between layers)."""
    B, H, W, k = 2, 37, 150, 3
    x, w, b, res = _layer_case(B, H, W, k, 32, seed=11, act=True, with_res=True)
    packed, mask = hm.pack_conv_weights(torch.from_numpy(w))
    plane, rplane = (hm.plane_from_nchw(torch.from_numpy(t).cuda()) for t in (x, res))
    outs = [hm.conv_layer(plane, packed.cuda(), mask, torch.from_numpy(b), k, hm.ACT_LEAKY | flag, B, H, W, residual=rplane)
            for flag in (0, hm.REVERSE_ORDER)]
    torch.cuda.synchronize()
    n = B * H * ((W + 1) // 2) * 128          # the plane proper (the buffer is rounded up to 1 KB)
    assert torch.equal(outs[0][:n], outs[1][:n])


@pytest.mark.gpu
def test_pack_input_matches_concat():
    rng = np.random.default_rng(5)
    B, H, W = 2, 11, 27
    a, f, r = (torch.from_numpy(rng.standard_normal((B, c, H, W)).astype(np.float32)).cuda() for c in (3, 2, 3))
    plane = hm.pack_input(a, f, r)
    torch.cuda.synchronize()
    full = plane[:B * H * (W + 1) * 64].view(torch.float16).reshape(B, H, W + 1, 32)
    want = torch.cat([a, f, r], 1).permute(0, 2, 3, 1).to(torch.float16)
    assert torch.equal(full[:, :, :W, :8], want)
    assert float(full[:, :, :, 8:].abs().max()) == 0.0 and float(full[:, :, W].abs().max()) == 0.0


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_network_matches_reference_fixture(mcn_kernel, golden, name):
    z, params = golden
    net = mirror_with_reference_weights(params, "cuda")
    args = [torch.from_numpy(z[f"{name}:{k}"]).cuda() for k in ("warped", "flow", "ref")]
    with torch.no_grad():
        out = net(*args)
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    same = om.mcn_forward(params, z[name + ":warped"], z[name + ":flow"], z[name + ":ref"], emulate_fp16=True)
    assert np.abs(got - same).max() <= TOL_NET_SAME_ROUNDING
    assert np.abs(got - z[name + ":out"]).max() <= TOL_NET_FP32


@pytest.mark.gpu
def test_network_full_size_against_torch_on_the_same_gpu(mcn_kernel, golden):
    """1080p: no CPU oracle finishes in seconds, so the comparison is against the same network evaluated by
    PyTorch/cuDNN in fp32 on the GPU (the definition is pinned by the fixture test above)."""
    _, params = golden
    net = mirror_with_reference_weights(params, "cuda")
    g = torch.Generator(device="cuda").manual_seed(3)
    B, H, W = 1, 1080, 1920
    a = torch.rand(B, 3, H, W, device="cuda", generator=g)
    r = torch.rand(B, 3, H, W, device="cuda", generator=g)
    f = torch.randn(B, 2, H, W, device="cuda", generator=g) * 4
    with torch.no_grad():
        got = net(a, f, r)
        # the same parameters through stock torch ops, fp32 (TF32 off)
        old = torch.backends.cudnn.allow_tf32
        torch.backends.cudnn.allow_tf32 = False
        try:
            F = torch.nn.functional
            layers = [(w.cuda(), b.cuda()) for w, b in net.folded_layers()]
            x = F.leaky_relu(F.conv2d(torch.cat([a, f, r], 1), layers[0][0], layers[0][1], padding=2), 0.2)
            for i in range(3):
                t = F.leaky_relu(F.conv2d(x, *layers[1 + 2 * i], padding=1), 0.2)
                x = F.leaky_relu(F.conv2d(t, *layers[2 + 2 * i], padding=1) + x, 0.2)
            want = a * torch.sigmoid(F.conv2d(x, *layers[7], padding=2))
        finally:
            torch.backends.cudnn.allow_tf32 = old
    assert float((got - want).abs().max()) <= TOL_NET_FP32
