"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy, fp64) of the reference's motion-compensation
network in inference mode.  Only tests/, __graft_entry__.smoke() and bench legs that report a CPU baseline
may import this; the product never does.

Follows R:codec_processing.py:
  * ConvNormAct            :117-156  Conv2d(bias=False, padding=k//2) -> BatchNorm2d -> LeakyReLU(0.2)
  * ResidualBlock          :190-217  ConvNormAct -> ConvNorm -> + input -> LeakyReLU(0.2)
  * MotionCompensationNetwork :369-406  cat(warped, flow, ref) -> ConvNormAct(k=5) -> ResidualBlock x n
                                        -> Conv2d(k=5, bias) -> Sigmoid ; refined = warped * map

``params`` is the network's ``state_dict()`` as numpy arrays (the reference's own key names).
Pinned by tests/golden/mcn.npz, produced by executing the reference's class definitions
(tests/golden/make_golden_mcn.py).

``emulate_fp16=True`` additionally rounds the folded weights, the network input and every layer's output
to fp16 -- the points where the CUDA path rounds -- so the kernel can be checked far below the
fp16-vs-fp32 difference.
"""
from __future__ import annotations

import numpy as np

BN_EPS = 1e-5          # nn.BatchNorm2d default, R:codec_processing.py:125 does not override it
LEAKY_SLOPE = 0.2      # R:codec_processing.py:110, :126


def conv2d_same(x: np.ndarray, w: np.ndarray) -> np.ndarray:
    """Cross-correlation with zero padding k//2, stride 1 (nn.Conv2d).  x: (B, Ci, H, W); w: (Co, Ci, k, k)."""
    B, Ci, H, W = x.shape
    Co, Ci2, k, k2 = w.shape
    assert Ci == Ci2 and k == k2 and k % 2 == 1
    r = k // 2
    xp = np.zeros((B, Ci, H + 2 * r, W + 2 * r), dtype=np.float64)
    xp[:, :, r:r + H, r:r + W] = x
    out = np.zeros((B, Co, H, W), dtype=np.float64)
    for ky in range(k):
        for kx in range(k):
            out += np.einsum("oc,bchw->bohw", w[:, :, ky, kx].astype(np.float64), xp[:, :, ky:ky + H, kx:kx + W])
    return out


def fold_bn(w, gamma, beta, mean, var, eps: float = BN_EPS):
    """Inference BatchNorm after a bias-free convolution == convolution with scaled weights plus a bias."""
    s = gamma.astype(np.float64) / np.sqrt(var.astype(np.float64) + eps)
    return w.astype(np.float64) * s[:, None, None, None], beta.astype(np.float64) - mean.astype(np.float64) * s


def _leaky(x):
    return np.where(x > 0, x, LEAKY_SLOPE * x)


def _h(x, on):
    return x.astype(np.float16).astype(np.float64) if on else x


def folded_layers(params: dict, num_res_blocks: int):
    """[(weight fp64, bias fp64)] in execution order: first conv, 2 per residual block, output conv."""
    def cn(prefix):
        return fold_bn(params[prefix + ".conv.weight"], params[prefix + ".norm.weight"], params[prefix + ".norm.bias"],
                       params[prefix + ".norm.running_mean"], params[prefix + ".norm.running_var"])
    layers = [cn("network.0")]
    for r in range(num_res_blocks):
        layers.append(cn(f"network.{1 + r}.block.0"))
        layers.append(cn(f"network.{1 + r}.block.1"))
    last = 1 + num_res_blocks
    layers.append((params[f"network.{last}.weight"].astype(np.float64), params[f"network.{last}.bias"].astype(np.float64)))
    return layers


def mcn_forward(params: dict, warped: np.ndarray, flow: np.ndarray, ref: np.ndarray, num_res_blocks: int = 3,
                emulate_fp16: bool = False) -> np.ndarray:
    q = emulate_fp16
    layers = [(_h(w, q), b) for w, b in folded_layers(params, num_res_blocks)]
    x = _h(np.concatenate([warped, flow, ref], axis=1).astype(np.float64), q)
    w, b = layers[0]
    x = _h(_leaky(conv2d_same(x, w) + b[None, :, None, None]), q)
    for r in range(num_res_blocks):
        w1, b1 = layers[1 + 2 * r]
        w2, b2 = layers[2 + 2 * r]
        t = _h(_leaky(conv2d_same(x, w1) + b1[None, :, None, None]), q)
        x = _h(_leaky(conv2d_same(t, w2) + b2[None, :, None, None] + x), q)
    w, b = layers[-1]
    m = 1.0 / (1.0 + np.exp(-(conv2d_same(x, w) + b[None, :, None, None])))
    return warped.astype(np.float64) * m


def conv_layer(x: np.ndarray, w: np.ndarray, b, act: bool, residual=None, emulate_fp16: bool = True) -> np.ndarray:
    """One layer as the kernel computes it: act(conv(x) + b [+ residual]) with fp16 operands / fp16 result."""
    q = emulate_fp16
    y = conv2d_same(_h(x.astype(np.float64), q), _h(w.astype(np.float64), q))
    if b is not None:
        y = y + np.asarray(b, dtype=np.float64)[None, :, None, None]
    if residual is not None:
        y = y + _h(residual.astype(np.float64), q)
    if act:
        y = _leaky(y)
    return _h(y, q)


def superpixel_gemm_conv(x: np.ndarray, packed: np.ndarray, ksize: int, nout: int, kmask: int) -> np.ndarray:
    """Evaluate a convolution from the PACKED per-tap matrices exactly the way the kernel walks them
    (taps over (dy, dsx), rows (q, co), columns (p, ci), k-steps gated by ``kmask``) -- the CPU check of
    rdvc_mcn_pack_weights.  x: (B, 32, H, W) with W even; returns (B, nout // 2, H, W)."""
    B, C, H, W = x.shape
    assert C == 32 and W % 2 == 0
    R = ksize // 2
    Wsp = W // 2
    sp = x.reshape(B, 32, H, Wsp, 2).transpose(0, 2, 3, 4, 1).reshape(B, H, Wsp, 64).astype(np.float64)   # [.., (p, ci)]
    pad = np.zeros((B, H + 2 * R, Wsp + 2, 64))
    pad[:, R:R + H, 1:1 + Wsp] = sp
    taps = packed.reshape(ksize * 3, nout, 64).astype(np.float64)
    acc = np.zeros((B, H, Wsp, nout))
    for dy in range(-R, R + 1):
        for dsx in (-1, 0, 1):
            t = (dy + R) * 3 + dsx + 1
            a = pad[:, R + dy:R + dy + H, 1 + dsx:1 + dsx + Wsp]
            for k in range(4):
                if (kmask >> (4 * t + k)) & 1:
                    acc += a[..., 16 * k:16 * k + 16] @ taps[t][:, 16 * k:16 * k + 16].T
    co = nout // 2
    return acc.reshape(B, H, Wsp, 2, co).transpose(0, 4, 1, 2, 3).reshape(B, co, H, W)


def superpixel_gemm_conv_xhalo(x: np.ndarray, packed: np.ndarray, ksize: int, nout: int, kmask: int) -> np.ndarray:
    """The same convolution evaluated the way the ONE-BOX kernel (csrc/mcn_convx_sm100.cuh) organises it: every tap
    multiplies the unshifted rows; the column-offset -1 taps accumulate into L (what a super-pixel sends to its right
    neighbour), the +1 taps into R (to its left neighbour), and the result is C[sp] + L[sp-1] + R[sp+1].  For 3x3 only
    the q = 0 rows of the L taps and the q = 1 rows of the R taps are used, as in the kernel (N = 32 MMAs)."""
    B, C, H, W = x.shape
    assert C == 32 and W % 2 == 0
    R = ksize // 2
    Wsp = W // 2
    sp = x.reshape(B, 32, H, Wsp, 2).transpose(0, 2, 3, 4, 1).reshape(B, H, Wsp, 64).astype(np.float64)
    pad = np.zeros((B, H + 2 * R, Wsp + 2, 64))                 # one halo column each side, R halo rows
    pad[:, R:R + H, 1:1 + Wsp] = sp
    taps = packed.reshape(ksize * 3, nout, 64).astype(np.float64)
    half = nout // 2
    acc_c = np.zeros((B, H, Wsp + 2, nout))
    acc_l = np.zeros((B, H, Wsp + 2, nout))
    acc_r = np.zeros((B, H, Wsp + 2, nout))
    for dy in range(-R, R + 1):
        a = pad[:, R + dy:R + dy + H]                           # unshifted in x, halo columns included
        for dsx, acc in ((0, acc_c), (-1, acc_l), (1, acc_r)):
            t = (dy + R) * 3 + dsx + 1
            rows = slice(0, nout)
            if R == 1 and dsx == -1:
                rows = slice(0, half)
            if R == 1 and dsx == 1:
                rows = slice(half, nout)
            for k in range(4):
                if (kmask >> (4 * t + k)) & 1:
                    acc[..., rows] += a[..., 16 * k:16 * k + 16] @ taps[t][rows, 16 * k:16 * k + 16].T
    out = acc_c[:, :, 1:1 + Wsp] + acc_l[:, :, 0:Wsp] + acc_r[:, :, 2:2 + Wsp]
    co = nout // 2
    return out.reshape(B, H, Wsp, 2, co).transpose(0, 4, 1, 2, 3).reshape(B, co, H, W)
