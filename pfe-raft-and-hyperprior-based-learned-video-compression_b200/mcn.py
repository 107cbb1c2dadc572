"""Host-side mirror of the reference's motion-compensation network ("next" row f-4, after resize_flow and the warp):

* ``MotionCompensationNetwork(input_channels=8, output_channels=3, base_channels=32, num_res_blocks=3)``
  with ``forward(warped_ref, flow, ref_frame)``                R:codec_processing.py:369-406 (called at :1458)

The module has the reference's parameter tree (``network.0.conv.weight``, ``network.1.block.0.norm.running_var``,
``network.4.bias`` ...), so ``load_state_dict`` takes a checkpoint of the reference's class unchanged.  ``forward`` is
inference only: the BatchNorms are folded into the convolutions on the host, the folded weights are packed once per
parameter version (``rdvc_mcn_pack_weights``) and the whole network runs as 1 + 8 CUDA launches behind
``rdvc_mcn_forward`` (tcgen05 implicit GEMMs over fp16 activations, fp32 accumulation).  There is no fallback: CPU
tensors, training mode and shapes the kernels do not cover raise.
"""
from __future__ import annotations

import ctypes
from typing import List, Optional

import numpy as np
import torch
from torch import Tensor, nn

from . import _cabi

ACT_NONE, ACT_LEAKY = 0, 1
REVERSE_ORDER = 0x100   # RDVC_MCN_REVERSE_ORDER: OR into `act`
_C = 32   # channels of the activation layout == the only base_channels the kernels cover


class _ConvNorm(nn.Module):
    """Parameter container with the key names of the reference's ConvNormAct (:117-156): ``conv`` (no bias), ``norm``."""

    def __init__(self, cin: int, cout: int, k: int):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, kernel_size=k, stride=1, padding=k // 2, bias=False)
        self.norm = nn.BatchNorm2d(cout)


class _ResBlock(nn.Module):
    """Key names of the reference's ResidualBlock (:190-217): ``block.0`` and ``block.1``."""

    def __init__(self, ch: int):
        super().__init__()
        self.block = nn.Sequential(_ConvNorm(ch, ch, 3), _ConvNorm(ch, ch, 3))


def fold_conv_norm(m: _ConvNorm):
    """(weight, bias) fp32 CPU tensors of conv followed by inference BatchNorm."""
    w = m.conv.weight.detach().double().cpu()
    n = m.norm
    s = n.weight.detach().double().cpu() / torch.sqrt(n.running_var.detach().double().cpu() + n.eps)
    b = n.bias.detach().double().cpu() - n.running_mean.detach().double().cpu() * s
    return (w * s[:, None, None, None]).float(), b.float()


def pack_conv_weights(weight: Tensor):
    """``rdvc_mcn_pack_weights`` on a (cout, cin, k, k) fp32 CPU tensor -> (uint8 CPU tensor of the packed fp16 tap
    matrices, k-step mask).  Pure host code: no GPU needed."""
    lib = _cabi.load()
    w = np.ascontiguousarray(weight.detach().cpu().float().numpy())
    cout, cin, k, k2 = w.shape
    if k != k2:
        raise ValueError(f"square kernels only, got {k} x {k2}")
    nbytes = lib.rdvc_mcn_packed_weight_bytes(k, cout)
    if nbytes == 0:
        raise ValueError(f"unsupported convolution for the MCN kernels: cout={cout} cin={cin} k={k}")
    packed = np.empty(nbytes, dtype=np.uint8)
    mask = ctypes.c_ulonglong(0)
    _cabi.check(lib.rdvc_mcn_pack_weights(w.ctypes.data, cout, cin, k, packed.ctypes.data, ctypes.byref(mask)),
                "rdvc_mcn_pack_weights")
    return torch.from_numpy(packed), int(mask.value)


def _need_cuda(*ts: Tensor):
    if any(not t.is_cuda for t in ts):
        raise RuntimeError("rdvc_corr_b200 runs on an sm_100 GPU only; got CPU tensors. There is no CPU fallback.")


def pack_input(warped: Tensor, flow: Tensor, ref: Tensor) -> Tensor:
    """cat(warped, flow, ref) -> activation plane (uint8 view of [B][H][ceil(W/2)][2][32] fp16)."""
    _need_cuda(warped, flow, ref)
    lib = _cabi.load()
    B, _, H, W = warped.shape
    a, f, r = (t.detach().float().contiguous() for t in (warped, flow, ref))
    plane = torch.empty(lib.rdvc_mcn_plane_bytes(B, H, W), dtype=torch.uint8, device=warped.device)
    with torch.cuda.device(warped.device):
        rc = lib.rdvc_mcn_pack_input(a.data_ptr(), f.data_ptr(), r.data_ptr(), B, a.shape[1], f.shape[1], r.shape[1], H, W,
                                     plane.data_ptr(), torch.cuda.current_stream(warped.device).cuda_stream)
    _cabi.check(rc, "rdvc_mcn_pack_input")
    return plane


def plane_from_nchw(x: Tensor) -> Tensor:
    """(B, C <= 32, H, W) float CUDA tensor -> activation plane (test / tooling helper, plain torch ops)."""
    _need_cuda(x)
    B, C, H, W = x.shape
    Wp = W + (W & 1)
    full = torch.zeros((B, H, Wp, _C), dtype=torch.float16, device=x.device)
    full[:, :, :W, :C] = x.permute(0, 2, 3, 1).to(torch.float16)
    nbytes = _cabi.load().rdvc_mcn_plane_bytes(B, H, W)
    plane = torch.zeros(nbytes, dtype=torch.uint8, device=x.device)
    plane[:full.numel() * 2] = full.reshape(-1).view(torch.uint8)
    return plane


def plane_to_nchw(plane: Tensor, B: int, H: int, W: int) -> Tensor:
    """Activation plane -> (B, 32, H, W) fp32, plus nothing else (the padding pixel is dropped)."""
    Wp = W + (W & 1)
    full = plane[:B * H * Wp * _C * 2].view(torch.float16).reshape(B, H, Wp, _C)
    return full[:, :, :W].permute(0, 3, 1, 2).float().contiguous()


def conv_layer(plane_in: Tensor, packed_dev: Tensor, kmask: int, bias: Optional[Tensor], ksize: int, act: int,
               B: int, H: int, W: int, residual: Optional[Tensor] = None) -> Tensor:
    """One 32 -> 32 layer on activation planes (``rdvc_mcn_conv``)."""
    _need_cuda(plane_in, packed_dev)
    lib = _cabi.load()
    out = torch.empty_like(plane_in)
    bias_np = None if bias is None else np.ascontiguousarray(bias.detach().cpu().float().numpy())
    if bias_np is not None and bias_np.size != _C:
        raise ValueError(f"bias must have {_C} entries")
    with torch.cuda.device(plane_in.device):
        rc = lib.rdvc_mcn_conv(plane_in.data_ptr(), packed_dev.data_ptr(), kmask,
                               None if bias_np is None else bias_np.ctypes.data, ksize, act,
                               None if residual is None else residual.data_ptr(), out.data_ptr(), B, H, W,
                               torch.cuda.current_stream(plane_in.device).cuda_stream)
    _cabi.check(rc, "rdvc_mcn_conv")
    return out


class MotionCompensationNetwork(nn.Module):
    """R:codec_processing.py:369-406.  Same constructor, same parameter names, same ``forward`` contract and errors."""

    def __init__(self, input_channels: int = 3 + 2 + 3, output_channels: int = 3, base_channels: int = 32,
                 num_res_blocks: int = 3):
        super().__init__()
        if base_channels != _C:
            raise ValueError(f"the B200 kernels cover base_channels == {_C} (the reference's default), got {base_channels}")
        if input_channels != 8 or output_channels != 3:
            raise ValueError("the B200 kernels cover the reference's default 8 input / 3 output channels")
        layers: List[nn.Module] = [_ConvNorm(input_channels, base_channels, 5)]
        layers += [_ResBlock(base_channels) for _ in range(num_res_blocks)]
        layers.append(nn.Conv2d(base_channels, output_channels, kernel_size=5, padding=2))
        layers.append(nn.Sigmoid())
        self.network = nn.Sequential(*layers)
        self.num_res_blocks = num_res_blocks
        self._packed = None       # (key, device weights, ptr array, masks, biases)

    # -- weights ---------------------------------------------------------------------------------------------
    def _conv_norms(self) -> List[_ConvNorm]:
        out = [self.network[0]]
        for r in range(self.num_res_blocks):
            out += [self.network[1 + r].block[0], self.network[1 + r].block[1]]
        return out

    def _version_key(self, device):
        vs = [(t.data_ptr(), t._version) for t in list(self.parameters()) + list(self.buffers())]
        return (str(device), tuple(vs))

    def folded_layers(self):
        """[(weight, bias)] fp32 CPU tensors in execution order (BatchNorms folded)."""
        layers = [fold_conv_norm(m) for m in self._conv_norms()]
        last = self.network[1 + self.num_res_blocks]
        layers.append((last.weight.detach().float().cpu(), last.bias.detach().float().cpu()))
        return layers

    def _prepare(self, device):
        key = self._version_key(device)
        if self._packed is not None and self._packed[0] == key:
            return self._packed
        folded = self.folded_layers()
        dev_w, masks = [], []
        biases = np.zeros((len(folded), _C), dtype=np.float32)
        for i, (w, b) in enumerate(folded):
            packed, mask = pack_conv_weights(w)
            dev_w.append(packed.to(device))
            masks.append(mask)
            biases[i, :b.numel()] = b.numpy()
        ptrs = (ctypes.c_void_p * len(dev_w))(*[t.data_ptr() for t in dev_w])
        mask_arr = (ctypes.c_ulonglong * len(masks))(*masks)
        self._packed = (key, dev_w, ptrs, mask_arr, np.ascontiguousarray(biases))
        return self._packed

    # -- forward ---------------------------------------------------------------------------------------------
    def forward(self, warped_ref: Tensor, flow: Tensor, ref_frame: Tensor) -> Tensor:
        if not (warped_ref.size() == ref_frame.size() and warped_ref.size()[-2:] == flow.size()[-2:]):
            raise ValueError("Input sizes mismatch in MotionCompensationNetwork. "
                             f"Warped: {warped_ref.shape}, Flow: {flow.shape}, Ref: {ref_frame.shape}")
        if flow.dim() != 4 or flow.shape[1] != 2:
            raise ValueError(f"Expected flow shape (B, 2, H, W), got {flow.shape}")
        if warped_ref.shape[1] != 3:
            raise ValueError(f"Expected 3-channel frames, got {warped_ref.shape}")
        if self.training:
            raise RuntimeError("rdvc_corr_b200.MotionCompensationNetwork is inference only (BatchNorm is folded); call .eval()")
        _need_cuda(warped_ref, flow, ref_frame)
        _cabi.forward_only("MotionCompensationNetwork", warped_ref, flow, ref_frame)
        lib = _cabi.load()
        dev = warped_ref.device
        _, _, ptrs, masks, biases = self._prepare(dev)
        B, _, H, W = warped_ref.shape
        a, f, r = (t.detach().float().contiguous() for t in (warped_ref, flow, ref_frame))
        ws_bytes = lib.rdvc_mcn_workspace_bytes(B, H, W)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        out = torch.empty_like(a)
        with torch.cuda.device(dev):
            rc = lib.rdvc_mcn_forward(a.data_ptr(), f.data_ptr(), r.data_ptr(), B, H, W, self.num_res_blocks, ptrs, masks,
                                      biases.ctypes.data, ws.data_ptr(), ws_bytes, out.data_ptr(),
                                      torch.cuda.current_stream(dev).cuda_stream)
        _cabi.check(rc, "rdvc_mcn_forward")
        return out.to(warped_ref.dtype)
