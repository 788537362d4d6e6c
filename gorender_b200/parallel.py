"""Multi-GPU partitioning: one process per GPU, launched with
torchrun; `torch.distributed` is the plumbing.

Two modes, both named in BASELINE.json's north_star:

* frame-parallel — a batch of poses is split into contiguous blocks, one per
  rank; every rank holds the whole scene and renders its frames into its own
  device framebuffers.  No data-path collective.
* sort-first strips — rank g rasterises the tile-aligned row strip
  `strip_rows(H, G, g)` of one large frame (geometry is replicated; K1/K2 run
  for the whole scene on every rank, bins are clamped to the strip), then the
  strips are gathered to rank 0: the one real exchange step of the path
  (NCCL gather over NVLink on GPUs; gloo in the CPU tests of the host logic).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np

from ._cabi import GRB_TILE


def pose_block(num_poses: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous block [begin, end) of poses owned by `rank`; sizes differ by at most one."""
    base, extra = divmod(num_poses, world_size)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def strip_rows(height: int, world_size: int, rank: int, tile: int = GRB_TILE) -> Tuple[int, int]:
    """Tile-aligned row range [y0, y1) of `rank`'s strip; tile rows are dealt out as evenly as
    possible, the last strip ends at `height`.  Ranks beyond the number of tile rows get (h, h)."""
    tile_rows = (height + tile - 1) // tile
    b, e = pose_block(tile_rows, world_size, rank)
    return min(b * tile, height), min(e * tile, height)


def gather_strips_to_rank0(color, depth, height: int, group=None):
    """Gather every rank's strip of (H, W, 4) uint8 colour / (H, W) float32 depth torch tensors
    into rank 0's full-frame tensors, in place.  Strips have different heights, so this is a
    grouped send/recv (== ncclGather with per-rank counts)."""
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if world == 1:
        return
    ops = []
    if rank == 0:
        for src in range(1, world):
            y0, y1 = strip_rows(height, world, src)
            if y1 > y0:
                ops.append(dist.P2POp(dist.irecv, color[y0:y1], src, group))
                ops.append(dist.P2POp(dist.irecv, depth[y0:y1], src, group))
    else:
        y0, y1 = strip_rows(height, world, rank)
        if y1 > y0:
            ops.append(dist.P2POp(dist.isend, color[y0:y1], 0, group))
            ops.append(dist.P2POp(dist.isend, depth[y0:y1], 0, group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()


class TorchFrameBuffer:
    """Device framebuffer whose memory is owned by torch tensors (`color` (frames,H,W,4) uint8,
    `depth` (frames,H,W) float32) and wrapped for the C ABI (`grb_framebuffer_wrap`) — what the
    NCCL gather of the strip mode operates on."""

    def __init__(self, width: int, height: int, frames: int, device, torch_device):
        import torch

        from .renderer import FrameBuffer

        self.color = torch.empty((frames, height, width, 4), dtype=torch.uint8, device=torch_device)
        self.depth = torch.empty((frames, height, width), dtype=torch.float32, device=torch_device)
        self.fb = FrameBuffer(width, height, frames, device, device_color=self.color.data_ptr(),
                              device_depth=self.depth.data_ptr())


def draw_strip(renderer, packed, height: int, world_size: int, rank: int, frame0: int = 0):
    """Rasterise this rank's strip of the frame described by `packed` (grb_object[1][nobj]).
    Returns the row range, or None when the rank owns no rows."""
    y0, y1 = strip_rows(height, world_size, rank)
    if y1 <= y0:
        return None
    renderer.draw_packed(packed, frame0, rows=(y0, y1), sync=False)
    return y0, y1


def frame_parallel_blocks(num_poses: int, world_size: int) -> List[Tuple[int, int]]:
    return [pose_block(num_poses, world_size, r) for r in range(world_size)]
