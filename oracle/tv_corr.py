"""TEST INFRASTRUCTURE ONLY -- the reference's own implementation of the path.

RDVC ships no CorrBlock; its encoder imports RAFT from torchvision
(R:codec_processing.py:48-53) and so runs
``torchvision.models.optical_flow.raft.CorrBlock`` (TV:raft.py:337-431).
torchvision is part of the image (0.26.0+cu128) both here and on the GPU box,
so this wrapper is the live "reference" arm: the parity oracle on CPU, the
``cpu_baseline`` (kind "reference") and ``bench.py --impl reference``.
"""
from __future__ import annotations

import torch


def tv_corr_block(num_levels: int = 4, radius: int = 4):
    from torchvision.models.optical_flow.raft import CorrBlock
    return CorrBlock(num_levels=num_levels, radius=radius)


@torch.no_grad()
def build_pyramid(fmap1: torch.Tensor, fmap2: torch.Tensor, num_levels: int = 4):
    """Returns the list of level tensors, each (B*N, 1, h_l, w_l)."""
    blk = tv_corr_block(num_levels=num_levels)
    blk.build_pyramid(fmap1, fmap2)
    return blk.corr_pyramid


@torch.no_grad()
def index_pyramid(levels, coords: torch.Tensor, radius: int = 4):
    blk = tv_corr_block(num_levels=len(levels), radius=radius)
    blk.corr_pyramid = list(levels)
    return blk.index_pyramid(centroids_coords=coords)


@torch.no_grad()
def build_and_lookup(fmap1, fmap2, coords_list, num_levels: int = 4, radius: int = 4):
    """One frame pair: build once, then one lookup per entry of coords_list."""
    blk = tv_corr_block(num_levels=num_levels, radius=radius)
    blk.build_pyramid(fmap1, fmap2)
    return [blk.index_pyramid(centroids_coords=c) for c in coords_list]
