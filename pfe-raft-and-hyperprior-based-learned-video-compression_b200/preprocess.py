"""Host-side mirror of the reference's two frame-preparation helpers ("next" row, caller side of the path):

* ``preprocess_frame_raft(frame_np_rgb, resize_shape_hw, device)``   R:codec_processing.py:751-761
* ``preprocess_frame_codec(frame_np_rgb, device)``                   R:codec_processing.py:763-769

Same names, arguments, result shape ``(1, C, H, W)`` fp32 in [0, 1] and the same error behaviour (a failure is
printed and ``None`` is returned).  The reference runs ``to_tensor`` + anti-aliased ``resize`` on the CPU and
uploads the float tensor; here the uint8 frame is uploaded (a quarter of the bytes) and one CUDA kernel
(``rdvc_preprocess_frame``) converts and resizes.  There is no fallback.
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import torch
from torch import Tensor

from . import _cabi


def frame_to_tensor(frame, out_hw: Optional[Sequence[int]], device) -> Tensor:
    """uint8 HWC frame (numpy array or torch tensor, host or device) -> (1, C, h, w) fp32 on ``device``."""
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("rdvc_corr_b200 runs on an sm_100 GPU only; got a CPU device. There is no CPU fallback.")
    t = torch.from_numpy(np.ascontiguousarray(frame)) if isinstance(frame, np.ndarray) else frame
    if t.dtype != torch.uint8:
        raise ValueError(f"frame must be uint8, got {t.dtype}")
    if t.dim() == 2:
        t = t.unsqueeze(-1)
    if t.dim() != 3:
        raise ValueError(f"frame must be (H, W, C), got {tuple(t.shape)}")
    H, W, C = t.shape
    h, w = (H, W) if out_hw is None else (int(out_hw[0]), int(out_hw[1]))
    t = t.to(dev, non_blocking=True).contiguous()
    out = torch.empty((1, C, h, w), dtype=torch.float32, device=dev)
    lib = _cabi.load()
    with torch.cuda.device(dev):
        rc = lib.rdvc_preprocess_frame(t.data_ptr(), H, W, C, out.data_ptr(), h, w,
                                       torch.cuda.current_stream(dev).cuda_stream)
    _cabi.check(rc, "rdvc_preprocess_frame")
    return out


def preprocess_frame_raft(frame_np_rgb, resize_shape_hw, device) -> Optional[Tensor]:
    """R:codec_processing.py:751-761."""
    try:
        return frame_to_tensor(frame_np_rgb, resize_shape_hw, device)
    except Exception as e:  # the reference prints and returns None (:760-761)
        print(f"Error preprocessing frame for RAFT: {e}")
        return None


def preprocess_frame_codec(frame_np_rgb, device) -> Optional[Tensor]:
    """R:codec_processing.py:763-769."""
    try:
        return frame_to_tensor(frame_np_rgb, None, device)
    except Exception as e:
        print(f"Error preprocessing frame for Codec: {e}")
        return None
