"""CPU, world_size 2 and 3 over gloo: the host side of the z-slab path.

Each rank takes its slab's (points, cells) counts - here from the oracle's per-slice counters, on the
GPU box from cub_count - all-gathers them, and must arrive at the id bases the single raster loop of the
reference reaches when it enters the rank's first slice."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from util import ROOT, gyroid, oracle, pkg


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, halo, result_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        P, O = pkg(), oracle()
        vol = gyroid((37, 20, 33), 9.0)
        ref = O.cuberille(vol, 0.0, triangles=True, project=False)
        slab = P.slabs.plan_slabs(vol.shape[0], world, halo)[rank]
        # what cub_count returns for this rank's own range
        n_points = int(ref.points_before_slice[slab.own_z1] - ref.points_before_slice[slab.own_z0])
        n_cells = int(ref.cells_before_slice[slab.own_z1] - ref.cells_before_slice[slab.own_z0])
        counts = P.slabs.all_gather_counts(n_points, n_cells)
        pbase, cbase = P.slabs.exclusive_bases(counts, rank)
        assert len(counts) == world
        assert pbase == int(ref.points_before_slice[slab.own_z0]), (rank, pbase)
        assert cbase == int(ref.cells_before_slice[slab.own_z0]), (rank, cbase)
        assert sum(c[0] for c in counts) == ref.points.shape[0] and sum(c[1] for c in counts) == ref.cells.shape[0]
        # the rank's cells, re-based, are exactly its segment of the reference's cell array
        seg = ref.cells[cbase:cbase + n_cells]
        assert seg.shape[0] == n_cells
        # vertices referenced by the segment that were created by lower slabs have ids below the base
        if n_cells:
            assert int(seg.max()) < pbase + n_points
        open(os.path.join(result_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,halo", [(2, 2), (3, 8)])
def test_count_allgather_gives_reference_id_bases(world, halo, tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, halo, str(tmp_path)), nprocs=world, join=True)
    assert sorted(os.listdir(tmp_path)) == [f"ok{r}" for r in range(world)]


def test_plan_slabs_covers_the_image_once():
    P = pkg()
    for nz, world, halo in [(10, 1, 2), (64, 8, 2), (100, 7, 8), (8, 8, 2)]:
        slabs = P.slabs.plan_slabs(nz, world, halo)
        assert slabs[0].own_z0 == 0 and slabs[-1].own_z1 == nz
        for a, b in zip(slabs, slabs[1:]):
            assert a.own_z1 == b.own_z0
        for s in slabs:
            assert s.own_z0 < s.own_z1
            assert s.local_z0 == max(0, s.own_z0 - halo) and s.local_z1 == min(nz, s.own_z1 + halo)
    with pytest.raises(ValueError):
        P.slabs.plan_slabs(10, 2, 1)
    with pytest.raises(ValueError):
        P.slabs.plan_slabs(3, 4, 2)
