"""Host-side mirror of the two reference functions between RAFT and the motion-compensation
network ("next" row f-4): same names, arguments and error behaviour as the reference's

* ``resize_flow(flow_tensor, target_hw)``          R:codec_processing.py:772-818
* ``WarpingLayer().forward(x, flow)``              R:codec_processing.py:322-367

plus the fused call the P-frame path actually wants (:1446 + :1456 in one launch):

* ``motion_warp(prev_frame, flow_at_raft_res, frame_hw) -> (warped_prev, flow_at_frame_res)``

All three are one CUDA kernel behind ``rdvc_motion_warp`` (include/rdvc_corr.h); there is no fallback.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import Tensor, nn

from . import _cabi


def _launch(prev: Optional[Tensor], flow: Tensor, H: int, W: int, want_flow: bool):
    if not flow.is_cuda or (prev is not None and not prev.is_cuda):
        raise RuntimeError("rdvc_corr_b200 runs on an sm_100 GPU only; got CPU tensors. There is no CPU fallback.")
    _cabi.forward_only("motion_warp / WarpingLayer / resize_flow", prev, flow)
    lib = _cabi.load()
    dev = flow.device
    f = flow.detach().to(torch.float32).contiguous()
    B, _, h_in, w_in = f.shape
    x = None if prev is None else prev.detach().to(torch.float32).contiguous()
    C = 0 if x is None else x.shape[1]
    warped = None if x is None else torch.empty_like(x)
    flow_out = torch.empty((B, 2, H, W), dtype=torch.float32, device=dev) if want_flow else None
    with torch.cuda.device(dev):
        rc = lib.rdvc_motion_warp(None if x is None else x.data_ptr(), f.data_ptr(), B, C, H, W, h_in, w_in,
                                  None if warped is None else warped.data_ptr(),
                                  None if flow_out is None else flow_out.data_ptr(),
                                  torch.cuda.current_stream(dev).cuda_stream)
    _cabi.check(rc, "rdvc_motion_warp")
    return warped, flow_out


def resize_flow(flow_tensor: Optional[Tensor], target_hw: Tuple[int, int]) -> Optional[Tensor]:
    """R:codec_processing.py:772-818, same early-outs: ``None`` in -> ``None``; wrong channel count ->
    ``ValueError``; equal sizes -> the input itself; a zero-area side -> zeros."""
    if flow_tensor is None:
        return None
    B, C, H_in, W_in = flow_tensor.shape
    if C != 2:
        raise ValueError(f"Flow tensor must have 2 channels, got {C}")
    H_out, W_out = target_hw
    if (H_in, W_in) == (H_out, W_out):
        return flow_tensor
    if H_in == 0 or W_in == 0 or H_out == 0 or W_out == 0:
        return torch.zeros(B, C, H_out, W_out, device=flow_tensor.device, dtype=flow_tensor.dtype)
    return _launch(None, flow_tensor, H_out, W_out, True)[1].to(flow_tensor.dtype)


class WarpingLayer(nn.Module):
    """R:codec_processing.py:322-367: warp ``x`` (B, C, H, W) by ``flow`` (B, 2, H, W), dx in channel 0."""

    def forward(self, x: Tensor, flow: Tensor) -> Tensor:
        B, C, H, W = x.size()
        if flow.size()[-2:] != (H, W) or flow.size()[1] != 2:
            raise ValueError(
                f"Input image ({B},{C},{H},{W}) and flow ({flow.shape}) shape/channel mismatch."
            )
        return _launch(x, flow, H, W, False)[0].to(x.dtype)


def motion_warp(prev_frame: Tensor, flow_at_raft_res: Tensor, frame_hw: Tuple[int, int]) -> Tuple[Tensor, Tensor]:
    """``resize_flow`` + ``WarpingLayer`` in one launch: (warped previous frame, flow at frame resolution)."""
    B, C, H, W = prev_frame.shape
    if tuple(frame_hw) != (H, W):
        raise ValueError(f"frame_hw {tuple(frame_hw)} does not match the previous frame ({H}, {W})")
    if flow_at_raft_res.dim() != 4 or flow_at_raft_res.shape[1] != 2:
        raise ValueError(f"Flow tensor must have 2 channels, got {flow_at_raft_res.shape[1]}")
    if flow_at_raft_res.shape[0] != B:
        raise ValueError("batch sizes of frame and flow differ")
    warped, flow = _launch(prev_frame, flow_at_raft_res, H, W, True)
    return warped, flow
