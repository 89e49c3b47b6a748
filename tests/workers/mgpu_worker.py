"""Worker of tests/test_gpu_multi.py (one process per GPU, launched by torch.distributed.run).

Every rank takes its z-slab (own range + the halo cub_projection_halo asks for) of ONE seeded volume, runs the
hot path with no host round trip (cub_count_async -> cub_comm_exchange_counts over NCCL -> cub_emit_async), and
the meshes are gathered with cub_comm_gather_mesh.  Rank 0 compares the gathered mesh with the CPU oracle's
mesh of the whole volume, bit for bit, and every rank checks that it received the same bytes.

usage: mgpu_worker.py <result_dir> <case>
"""
import hashlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from util import gyroid, oracle, pkg, smooth_volume  # noqa: E402

CASES = {
    # name: (volume factory, iso, triangles, project, cell_data, params)
    "smooth64_tri_proj": (lambda: smooth_volume((64, 64, 64), np.float32, seed=11), True, True, False),
    "smooth_u8_quads_celldata": (lambda: smooth_volume((40, 33, 70), np.uint8, seed=5), False, False, True),
    "gyroid_tri_noproj": (lambda: (gyroid((48, 40, 96), 17.0), 0.0), True, False, False),
    "smooth_i16_tri_proj_celldata": (lambda: smooth_volume((57, 31, 45), np.int16, seed=3), True, True, True),
}


def main():
    result_dir, case = sys.argv[1], sys.argv[2]
    rank, world, local_rank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist.init_process_group("nccl", device_id=dev)
    try:
        P, O = pkg(), oracle()
        make, tri, proj, cd = CASES[case]
        vol, iso = make()
        prm = P.capi.default_params()
        prm.iso_value, prm.generate_triangles, prm.project_vertices = float(iso), int(tri), int(proj)
        prm.save_pixel_as_cell_data = int(cd)
        prm.surface_distance_threshold, prm.step_length, prm.max_steps = 0.02, 0.24, 100
        halo = max(P.capi.projection_halo(prm))
        slab = P.slabs.plan_slabs(vol.shape[0], world, halo)[rank]
        stream = torch.cuda.Stream(device=dev)
        torch.cuda.set_stream(stream)
        h = P.capi.Handle(local_rank, stream.cuda_stream)
        h.set_volume(vol[slab.local_z0:slab.local_z1])
        h.set_slab(vol.shape[0], slab.local_z0, slab.own_z0, slab.own_z1)
        comm = P.slabs.create_comm(h, dev)
        for _ in range(2):  # the second pass runs with buffers sized by the first: the fully asynchronous path
            P.slabs.step_async(h, comm, prm, 4)
        pts, cells, cdata = P.slabs.gather_mesh(h, comm, want_cell_data=cd)
        pts, cells = pts.cpu().numpy(), cells.cpu().numpy().view(np.uint32)
        cdata = cdata.cpu().numpy().view(vol.dtype) if cd else None
        digest = hashlib.sha256(pts.tobytes() + cells.tobytes() + (cdata.tobytes() if cd else b"")).hexdigest()
        if rank == 0:
            ref = O.cuberille(vol, iso, triangles=tri, project=proj, cell_data=cd, thr=0.02, step=0.24, relax=0.95, max_steps=100)
            ok = (pts.shape == ref.points.shape and cells.shape == ref.cells.shape
                  and np.array_equal(cells.astype(np.uint64), ref.cells)
                  and np.array_equal(pts.view(np.uint32), ref.points.view(np.uint32))
                  and (not cd or np.array_equal(cdata, ref.cell_data)))
            msg = f"{case}: world {world}, {pts.shape[0]} points, {cells.shape[0]} cells, halo {halo}: " + ("bit-exact vs oracle" if ok else "MISMATCH")
            print(msg, flush=True)
            open(os.path.join(result_dir, "verdict"), "w").write(("ok " if ok else "FAIL ") + msg)
        open(os.path.join(result_dir, f"digest{rank}"), "w").write(digest)
        comm.close()
        h.close()
    finally:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
