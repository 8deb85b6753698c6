"""Row-tile bookkeeping of the multi-GPU partition (host logic, no GPU) and a world_size-2 gloo
run of the gather: every image row is owned by exactly one rank and lands at its place."""
import ctypes as C
import os
import socket
import subprocess
import sys
import textwrap
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.parametrize("height,tile_rows,world", [(640, 2, 2), (950, 4, 8), (1350, 8, 4), (7, 3, 2), (5, 8, 4), (640, 1, 3)])
def test_tiles_partition_the_rows(pkg, height, tile_rows, world):
    from hslu_i.ba_raytracing.f2501_raytracer_b200.distributed import padded_rows, tile_rows_of_rank

    rc = pkg.RenderConfig.from_features("", 0, height=height)
    owned = [tile_rows_of_rank(rc, tile_rows, r, world) for r in range(world)]
    flat = sorted(sum(owned, []))
    assert flat == list(range(height))  # every row exactly once
    for r, rows in enumerate(owned):
        assert rows == sorted(rows)
        assert all((y // tile_rows) % world == r for y in rows)
    assert padded_rows(rc, tile_rows, world) == max(map(len, owned))
    # world 1 / tiling off: identity
    assert tile_rows_of_rank(rc, tile_rows, 0, 1) == list(range(height))


@pytest.mark.parametrize("height,tile_rows,world", [(640, 2, 8), (950, 4, 8), (23, 3, 2), (5, 8, 4), (641, 1, 3)])
def test_place_tiles_mapping_inverts_the_tile_bookkeeping(pkg, height, tile_rows, world):
    """place_tiles_kernel (rt_kernels.cuh) maps image row -> (rank, local row) with
    t = row / tile_rows, rank = t % world, local = (t / world) * tile_rows + row % tile_rows;
    that must be the inverse of rt_tile_row(), which every rank uses to place its rows."""
    from hslu_i.ba_raytracing.f2501_raytracer_b200.distributed import tile_rows_of_rank

    rc = pkg.RenderConfig.from_features("", 0, height=height)
    owned = [tile_rows_of_rank(rc, tile_rows, r, world) for r in range(world)]
    for row in range(height):
        t, r_in = divmod(row, tile_rows)
        rank, local = t % world, (t // world) * tile_rows + r_in
        assert owned[rank][local] == row


def test_invalid_tiling_is_rejected(lib, pkg):
    rc = pkg.RenderConfig.from_features("", 0).with_tiles(4, 5, 2)  # rank >= world
    # validation happens before any device work
    assert lib.rt_render_device(C.c_void_p(1), C.byref(rc.cfg), C.c_void_p(1), None, None, None) in (
        pkg._lib.RT_ERR_INVALID_ARGUMENT,)


WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, {root!r})
    import numpy as np, torch, torch.distributed as dist
    import hslu_i.ba_raytracing.f2501_raytracer_b200 as pkg
    from hslu_i.ba_raytracing.f2501_raytracer_b200.distributed import padded_rows, tile_rows_of_rank
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    W, H, T = 37, 23, 3
    rc = pkg.RenderConfig.from_features("", 0, width=W, height=H)
    rows = tile_rows_of_rank(rc, T, rank, world)
    pad = padded_rows(rc, T, world)
    # each rank "renders" its rows: pixel value = image row * 1000 + x (what rt_render_device's compact slab holds)
    slab = torch.zeros(pad * W, dtype=torch.int32)
    for j, y in enumerate(rows):
        slab[j * W:(j + 1) * W] = torch.arange(W, dtype=torch.int32) + 1000 * y
    gathered = [torch.zeros_like(slab) for _ in range(world)] if rank == 0 else None
    dist.gather(slab, gathered, dst=0)
    if rank == 0:
        frame = np.full((H, W), -1, np.int64)
        for r in range(world):
            for j, y in enumerate(tile_rows_of_rank(rc, T, r, world)):  # same mapping rt_scatter_tile_rows applies on the GPU
                frame[y] = gathered[r][j * W:(j + 1) * W].numpy()
        expect = np.arange(W)[None, :] + 1000 * np.arange(H)[:, None]
        assert np.array_equal(frame, expect), "gathered frame is wrong"
        print("GATHER_OK")
    dist.destroy_process_group()
""")


def test_gather_of_row_tiles_world2_gloo(pkg, tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=str(ROOT)))
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r), LOCAL_RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "GATHER_OK" in outs[0]
