"""Multi-GPU frame partition: interleaved row tiles, one process per GPU, NCCL gather.

The reference renders shuffled 48x48 chunks on a rayon pool inside one process
(src/image_buffer.rs:48-97); pixels are independent, so the frame shards with no data-path
exchange.  Here image rows are dealt out in tiles of `tile_rows` rows, tile t -> rank t % world
(interleaved because cost is very non-uniform: glass spheres vs. the black border).  Every rank
renders its tiles into a compact device slab (rt_render_device); the only collective is the gather
of the finished slabs to rank 0 over NVLink (torch.distributed / NCCL), after which rank 0 places
all rows with rt_place_tiles (one kernel; the hit mask is `pixel != 0`).  Per-pixel sampler streams are keyed by (seed, x, y, light), so
the image does not depend on the world size.
"""
from __future__ import annotations

import ctypes as C

from . import _lib as L
from .renderer import DeviceScene, RenderConfig

DEFAULT_TILE_ROWS = 2  # fine interleave: ranks' costs within ~2 % of each other on the headline frame (tools/gpu_tiles_onegpu.py)


def padded_rows(rc: RenderConfig, tile_rows: int, world: int) -> int:
    """Rows of the largest slab of any rank (gather buffers are padded to this)."""
    return max(rc.with_tiles(tile_rows, r, world).local_rows() for r in range(world))


class TiledRenderer:
    """One rank's share of a frame + the gather to rank 0."""

    def __init__(self, scene: DeviceScene, rc: RenderConfig, rank: int, world: int, tile_rows: int = DEFAULT_TILE_ROWS):
        import torch

        self.torch = torch
        self.scene, self.rank, self.world, self.tile_rows = scene, rank, world, tile_rows
        self.rc_full = rc
        self.rc = rc.with_tiles(tile_rows, rank, world) if world > 1 else rc
        self.width, self.height = rc.cfg.width, rc.cfg.height
        self.rows = self.rc.local_rows()
        self.pad_rows = padded_rows(rc, tile_rows, world) if world > 1 else self.rows
        dev = torch.device("cuda", scene.device)
        self.slab = torch.zeros(self.pad_rows * self.width, dtype=torch.int32, device=dev)
        self.mask = torch.zeros(self.pad_rows * self.width, dtype=torch.uint8, device=dev)
        self.frame = self.frame_mask = self.gathered = None
        if rank == 0:
            self.frame = torch.zeros(self.height * self.width, dtype=torch.int32, device=dev)
            self.frame_mask = torch.zeros(self.height * self.width, dtype=torch.uint8, device=dev)
            if world > 1:  # one contiguous buffer: slab r at r * pad_rows * width
                self.gathered = torch.zeros(world * self.pad_rows * self.width, dtype=torch.int32, device=dev)
                self.gather_list = list(self.gathered.view(world, -1).unbind(0))

    def render_local(self) -> None:
        """Enqueue this rank's tiles on torch's current stream (no host sync)."""
        stream = self.torch.cuda.current_stream().cuda_stream
        self.scene.render_device(self.rc, self.slab.data_ptr(), self.mask.data_ptr(), stream)

    def gather(self, with_mask: bool = True) -> None:
        """Gather the slabs on rank 0 (ONE collective: a slab pixel is 0xFFrrggbb for a hit and 0 for a
        miss, so the hit mask travels inside it) and place all rows into the full frame (one kernel)."""
        torch = self.torch
        lib = L.load()
        stream = torch.cuda.current_stream().cuda_stream
        if self.world == 1:
            self.frame.copy_(self.slab)
            if with_mask:
                self.frame_mask.copy_(self.mask)
            return
        import torch.distributed as dist

        dist.gather(self.slab, self.gather_list if self.rank == 0 else None, dst=0)
        if self.rank == 0:
            cfg = self.rc_full.with_tiles(self.tile_rows, 0, self.world).cfg
            L.check(lib.rt_place_tiles(C.byref(cfg), self.gathered.data_ptr(), self.pad_rows * self.width, self.frame.data_ptr(),
                                       self.frame_mask.data_ptr() if with_mask else None, stream))

    def check(self) -> None:
        """rt_check_frame: the frames enqueued by render_local() since the last check are valid (raises otherwise)."""
        self.scene.check()

    def step(self) -> None:
        self.render_local()
        self.gather()


def tile_rows_of_rank(rc: RenderConfig, tile_rows: int, rank: int, world: int) -> list[int]:
    """Image rows owned by `rank` in slab order (host-side bookkeeping; used by tests)."""
    lib = L.load()
    cfg = rc.with_tiles(tile_rows, rank, world).cfg
    return [lib.rt_tile_row(C.byref(cfg), j) for j in range(lib.rt_tile_local_rows(C.byref(cfg)))]
