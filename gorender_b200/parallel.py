"""Multi-GPU partitioning: one process per GPU, launched with
torchrun; `torch.distributed` is the plumbing.

Two modes, both named in BASELINE.json's north_star:

* frame-parallel — a batch of poses is split into contiguous blocks, one per
  rank; every rank holds the whole scene and renders its frames into its own
  device framebuffers.  No data-path collective.
* sort-first strips — rank g renders a tile-aligned row strip of one large
  frame.  Geometry is replicated; the setup kernel of a rank retires every
  256-face block whose projected bounds miss its rows after eight vertex
  transforms (setup.cu, block_rejected), so a rank's geometry work follows its
  share of the screen.  `StripGroup` is the B200 form of the exchange step:
  rank 0's framebuffer is shared with the other processes (CUDA IPC), every
  rank's raster kernel stores its rows straight into it over NVLink, and the
  hand-off is a device-side flag per rank — no gather, no host in the loop.
  Strips are balanced by the busy tiles of a probe frame, not by equal rows.
  `gather_strips_to_rank0` (grouped NCCL send/recv on torch-owned framebuffers,
  round 1's form; gloo in the CPU tests of the host logic) is kept beside it.
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


def balanced_strip_rows(row_weights: Sequence[float], world_size: int, height: int,
                        tile: int = GRB_TILE) -> List[Tuple[int, int]]:
    """Contiguous tile-aligned row ranges [y0, y1), one per rank, with about equal total weight.
    `row_weights[r]` is the cost of tile row r (e.g. its number of busy tiles in a probe frame); rows
    of zero weight are free and go to whichever neighbour the cut leaves them with.  Every rank gets a
    (possibly empty) range, the ranges tile [0, height) in rank order."""
    w = np.asarray(row_weights, dtype=np.float64)
    n = len(w)
    assert n == (height + tile - 1) // tile
    total = float(w.sum())
    if total <= 0:
        return [strip_rows(height, world_size, r, tile) for r in range(world_size)]
    csum = np.concatenate([[0.0], np.cumsum(w)])
    cuts = [0]
    for r in range(1, world_size):
        target = total * r / world_size
        k = int(np.searchsorted(csum, target, side="left"))      # first k with csum[k] >= target
        if k > 0 and abs(csum[k - 1] - target) <= abs(csum[min(k, n)] - target):
            k -= 1
        cuts.append(min(max(k, cuts[-1]), n))
    cuts.append(n)
    return [(min(cuts[r] * tile, height), min(cuts[r + 1] * tile, height)) for r in range(world_size)]


class StripGroup:
    """Sort-first strips over the ranks of a torch.distributed process group (one process per GPU of one node).

    Rank 0 owns `nbuf` device framebuffers and shares them (CUDA IPC handles travel through the process
    group as host bytes); the other ranks open them.  `draw(k, packed)` renders this rank's rows of the frame(s):
    rank 0 straight into buffer k, every other rank into a framebuffer of its own, whose rows it then pushes into
    buffer k with a mirror laid over rank 0's memory — only the tiles that are busy (or were, in rank 0's copy)
    cross NVLink, on the copy stream, while the render stream is already working on the next call — and raises
    its flag.  On rank 0 `draw` also makes the render stream wait for every rank's flag, so whatever rank 0
    queues next (a host mirror update, a read-back) sees the whole frame.  `release(k)` (rank 0) tells the others
    that buffer k may be overwritten; their next push into it waits for that on the device.  Nothing blocks the
    host.  A buffer may hold several consecutive frames of an animation (`frames`), drawn by one call per rank:
    the hand-off and launch costs are then paid once per call, not once per frame."""

    CONSUMED_SLOT = 63

    def __init__(self, device, width: int, height: int, nbuf: int = 2, group=None, rows: Optional[List[Tuple[int, int]]] = None,
                 frames: int = 1):
        import torch.distributed as dist

        from ._cabi import GRB_PLANE_COLOR, GRB_PLANE_DEPTH
        from .renderer import FrameBuffer, Mirror, Renderer

        self.dist, self.group = dist, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.dev, self.width, self.height = device, width, height
        self.rows = rows if rows is not None else [strip_rows(height, self.world, r) for r in range(self.world)]
        assert len(self.rows) == self.world
        self.frames = frames
        if self.rank == 0:
            self.fbs = [FrameBuffer(width, height, frames, device) for _ in range(nbuf)]
            handles = [fb.ipc_export() for fb in self.fbs] if self.world > 1 else [None] * nbuf
        else:
            handles = [None] * nbuf
        if self.world > 1:
            box = [handles]
            dist.broadcast_object_list(box, src=0, group=group)
            handles = box[0]
        self.local, self.push = [], []
        if self.rank != 0:
            self.fbs = [FrameBuffer(width, height, frames, device, ipc_handle=h) for h in handles]     # rank 0's memory
            self.local = [FrameBuffer(width, height, frames, device) for _ in range(nbuf)]             # what this rank renders into
            self.push = [(Mirror(device, width, height, frames, GRB_PLANE_COLOR, target=fb),
                          Mirror(device, width, height, frames, GRB_PLANE_DEPTH, target=fb)) for fb in self.fbs]
        self.renderers = [Renderer(fb) for fb in (self.local if self.rank != 0 else self.fbs)]
        self.uses = [0] * nbuf
        self.released = [0] * nbuf

    def set_rows(self, rows: List[Tuple[int, int]]) -> None:
        """A new partition: what a rank knows about the tiles of rank 0's copy no longer covers the rows it renders."""
        assert len(rows) == self.world
        self.rows = rows
        for mc, mz in self.push:
            mc.invalidate()
            mz.invalidate()

    def draw(self, k: int, packed, timeout_ms: int = 5000):
        """This rank's strip of the frame(s) `packed` (grb_object[frames][nobj]; a buffer holds `frames` consecutive
        frames of an animation, each split the same way) into buffer k; asynchronous."""
        shared, r = self.fbs[k], self.renderers[k]
        n = self.uses[k] + 1
        y0, y1 = self.rows[self.rank]
        nf = packed.shape[0]
        if y1 > y0:
            r.draw_packed(packed, 0, rows=(y0, y1) if self.world > 1 else None, sync=False)
        if self.world > 1:
            if self.rank != 0:
                if n > 1:
                    shared.wait_signals(self.CONSUMED_SLOT, 1, n - 1, timeout_ms, on_copy_stream=True)   # rank 0 is done with the buffer's previous frames
                if y1 > y0:
                    mc, mz = self.push[k]
                    self.local[k].update_mirrors_async(0, nf, mc, mz, rows=(y0, y1))    # copy stream, behind the draw
                shared.signal(self.rank, n, on_copy_stream=True)
            else:
                shared.signal(0, n)
                shared.wait_signals(0, self.world, n, timeout_ms)
        self.uses[k] = n
        return y0, y1

    def release(self, k: int) -> None:
        """Rank 0: everything queued so far that reads buffer k comes before the other ranks' next writes into it."""
        if self.world > 1 and self.rank == 0 and self.released[k] != self.uses[k]:
            self.fbs[k].signal(self.CONSUMED_SLOT, self.uses[k])
            self.released[k] = self.uses[k]

    def close_local(self) -> None:
        """Not collective: drop this rank's handles (after a failed collective set-up)."""
        self.dev.synchronize()
        for mc, mz in self.push:
            mc.close()
            mz.close()
        for fb in self.fbs + self.local:
            fb.close()
        self.fbs, self.local, self.push = [], [], []

    def close(self) -> None:
        """Collective: the other ranks unmap rank 0's memory before rank 0 frees it."""
        self.dev.synchronize()
        if self.world > 1 and self.rank != 0:
            for mc, mz in self.push:
                mc.close()
                mz.close()
            for fb in self.fbs + self.local:
                fb.close()
        if self.world > 1:
            self.dist.barrier(group=self.group)
        if self.rank == 0:
            for fb in self.fbs:
                fb.close()
        self.fbs, self.local, self.push = [], [], []


def gather_strips_to_rank0(color, depth, height: int, group=None, rows: Optional[List[Tuple[int, int]]] = None):
    """Gather every rank's strip of (H, W, 4) uint8 colour / (H, W) float32 depth torch tensors
    into rank 0's full-frame tensors, in place.  Strips have different heights, so this is a
    grouped send/recv (== ncclGather with per-rank counts).  `rows`: the ranks' row ranges
    (default: equal tile rows, `strip_rows`)."""
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if world == 1:
        return
    if rows is None:
        rows = [strip_rows(height, world, r) for r in range(world)]
    ops = []
    if rank == 0:
        for src in range(1, world):
            y0, y1 = rows[src]
            if y1 > y0:
                ops.append(dist.P2POp(dist.irecv, color[y0:y1], src, group))
                ops.append(dist.P2POp(dist.irecv, depth[y0:y1], src, group))
    else:
        y0, y1 = rows[rank]
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
