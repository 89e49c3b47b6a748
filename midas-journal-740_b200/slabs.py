"""z-slab decomposition of the hot path across ranks.

The reference's GenerateData is one raster loop (txx:136-206); concatenating contiguous z-slabs keeps
its cell order and its first-touch vertex order.  Each rank runs the same kernels on its slab (with a
halo) and the ONLY exchange of the data path is an all-gather of two integers per rank: the exclusive
scan of (points, cells) counts gives the global id bases.  Host logic only - works with any
torch.distributed backend (nccl on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from dataclasses import dataclass


@dataclass(frozen=True)
class Slab:
    rank: int
    image_nz: int
    own_z0: int      # global half-open range of slices whose voxels this rank emits
    own_z1: int
    local_z0: int    # global half-open range of slices the rank holds (own range + halo, clipped to the image)
    local_z1: int


def plan_slabs(image_nz: int, world: int, halo: int = 2) -> list[Slab]:
    """Contiguous, near-equal z-slabs; `halo` >= 2 (classification + first-touch ownership of the shared
    corner plane); when vertices are projected they travel up to step / (1 - relax) = 5 x the largest spacing by
    default: use >= 8 for isotropic voxels, and 5 * max(spacing) / spacing_z + 3 in general."""
    if halo < 2:
        raise ValueError("the halo must be at least 2 slices")
    if world < 1 or image_nz < world:
        raise ValueError("need at least one slice per rank")
    out = []
    for r in range(world):
        z0, z1 = (image_nz * r) // world, (image_nz * (r + 1)) // world
        out.append(Slab(r, image_nz, z0, z1, max(0, z0 - halo), min(image_nz, z1 + halo)))
    return out


def exclusive_bases(counts: list[tuple[int, int]], rank: int) -> tuple[int, int]:
    """(point id base, cell id base) of `rank` from every rank's (points, cells)."""
    return sum(c[0] for c in counts[:rank]), sum(c[1] for c in counts[:rank])


def all_gather_counts(n_points: int, n_cells: int, device=None, stream=None) -> list[tuple[int, int]]:
    """All-gather of the per-rank (points, cells) counts: NCCL has no exclusive scan, so every rank gathers
    the 2 x world integers and sums its prefix locally.  Single process: no communication.
    `stream`: a side torch stream for the exchange, so that it does not queue behind kernels the caller has
    already launched on the current stream (Handle.emit_vertices runs meanwhile)."""
    import contextlib
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [(int(n_points), int(n_cells))]
    world = dist.get_world_size()
    ctx = torch.cuda.stream(stream) if stream is not None else contextlib.nullcontext()
    with ctx:
        mine = torch.tensor([n_points, n_cells], dtype=torch.int64, device=device)
        out = torch.empty(2 * world, dtype=torch.int64, device=device)
        dist.all_gather_into_tensor(out, mine)
        flat = out.cpu().tolist()
    return [(flat[2 * r], flat[2 * r + 1]) for r in range(world)]


def run_streamed(handles, vol_ptr: int, dtype, dims_xyz, params, n_slabs: int, out_points_ptr: int, out_cells_ptr: int,
                 id_bytes: int = 4, halo: int = 2, spacing=(1.0, 1.0, 1.0), origin=(0.0, 0.0, 0.0),
                 device_volume=None, streams=None):
    """One GPU, host volume in, host mesh out, streamed: the image is cut into `n_slabs` z-slabs that travel
    through `handles` (>= 2 capi.Handle objects, each on its own stream) round-robin, so the host->device copy
    of slab c+1, the kernels of slab c and the device->host copy of slab c-1 overlap.  Ids are global: the id
    base of a slab is the running sum of the counts of the slabs before it (the same exclusive scan the
    multi-GPU path gets from its all-gather).  `vol_ptr` / `out_*_ptr`: pinned host memory; the outputs must
    be large enough for the whole mesh.  Returns (n_points, n_cells).

    Two ways for the volume to travel:
    * default: every handle copies its slab *with* its halo slices into its own buffer (cub_set_volume with
      CUB_MEM_HOST); works for volumes larger than the device memory, re-sends 2 * halo slices per cut;
    * `device_volume` (a torch uint8 CUDA tensor of the volume's size) + `streams` (the torch streams the
      handles were created on): the volume is copied once, in n_slabs consecutive pieces on a copy stream, into
      that buffer and the handles borrow windows of it (CUB_MEM_DEVICE); nothing is sent twice."""
    import numpy as np
    from . import capi
    nx, ny, nz = dims_xyz
    item = np.dtype(dtype).itemsize
    slice_bytes = ny * nx * item
    slabs = plan_slabs(nz, n_slabs, halo)
    verts_per_cell = 3 if params.generate_triangles else 4
    cells_per_quad = 2 if params.generate_triangles else 1
    pbase = cbase = 0
    resident = device_volume is not None
    events = []
    if resident:
        import ctypes
        import torch
        assert streams is not None and len(streams) == len(handles), "resident mode needs the handles' torch streams"
        assert device_volume.is_cuda and device_volume.numel() * device_volume.element_size() >= nz * slice_bytes
        dev = device_volume.view(torch.uint8).reshape(-1)
        host = torch.from_numpy(np.ctypeslib.as_array((ctypes.c_uint8 * (nz * slice_bytes)).from_address(vol_ptr)))
        copy_stream = _copy_stream(dev.device)
        for st in streams:
            copy_stream.wait_stream(st)  # earlier work on the handles may still read the buffer
        done = 0
        with torch.cuda.stream(copy_stream):
            for s in slabs:  # piece c ends where slab c's window ends: after it, slab c has all it reads
                a, b = done * slice_bytes, s.local_z1 * slice_bytes
                if b > a:
                    dev[a:b].copy_(host[a:b], non_blocking=True)
                done = max(done, s.local_z1)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
                events.append(ev)

    def upload(c):
        s, h = slabs[c], handles[c % len(handles)]
        if resident:
            streams[c % len(handles)].wait_event(events[c])
            h.set_volume_ptr(device_volume.data_ptr() + s.local_z0 * slice_bytes, dtype, (nx, ny, s.local_z1 - s.local_z0),
                             capi.MEM_DEVICE, spacing, origin)
        else:
            h.set_volume_ptr(vol_ptr + s.local_z0 * slice_bytes, dtype, (nx, ny, s.local_z1 - s.local_z0), capi.MEM_HOST,
                             spacing, origin)
        h.set_slab(nz, s.local_z0, s.own_z0, s.own_z1)

    upload(0)
    for c in range(n_slabs):
        h = handles[c % len(handles)]
        if c + 1 < n_slabs:
            upload(c + 1)             # queued on the next handle's stream before this slab's count blocks the host
        n_pts, n_quads = h.count(params)
        h.set_id_base(pbase, cbase)
        h.emit(id_bytes)
        h.fetch_into(out_points_ptr + pbase * 12, out_cells_ptr + cbase * verts_per_cell * id_bytes, 0, capi.MEM_HOST,
                     sync=False)
        pbase += n_pts
        cbase += n_quads * cells_per_quad
    for h in handles:
        h.synchronize()
    return pbase, cbase


_COPY_STREAMS = {}


def _copy_stream(device):
    import torch
    key = (device.type, device.index)
    if key not in _COPY_STREAMS:
        _COPY_STREAMS[key] = torch.cuda.Stream(device=device)
    return _COPY_STREAMS[key]
