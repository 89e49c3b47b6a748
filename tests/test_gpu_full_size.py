"""GPU: BASELINE.json-size workloads checked through size-independent properties (the oracle would need
minutes here): generated on the device, run through the C-ABI, mesh pulled back and checked with numpy.

* every vertex id is used, every quad is a unit square (unprojected positions are exact lattice corners)
* the surface is closed: the generators force the outermost voxel layer outside, so every undirected edge is
  shared by an even number of quads (2, or 4 where two sheets touch along an edge)
* Euler-Poincare style count: for a quad mesh in which every edge has multiplicity m_e, sum m_e = 4 F
* idempotence: a second run on the same handle gives the same bytes
* z-slab decomposition (4 slabs) concatenates to the same bytes
* raster vertex order is a renumbering: same sorted point set, same cells after relabelling
* a z-sub-slab of the same bytes goes through the oracle (bit-exact), so the properties are anchored
"""
import numpy as np
import pytest

from util import assert_mesh_equal, oracle, pkg

pytestmark = pytest.mark.gpu


def _run(P, h, tri=False, proj=False, order=0, thr=0.5):
    p = P.capi.default_params()
    p.iso_value, p.generate_triangles, p.project_vertices, p.vertex_order = h._iso, int(tri), int(proj), order
    p.surface_distance_threshold = thr
    h.run(p)
    pts, cells, _ = h.fetch()
    return pts, cells


@pytest.mark.parametrize("kind,size,p0,iso", [("gyroid", 512, 64.0, 0.0), ("marschner_lobb", 512, 0.0, 0.5), ("blobs", 384, 48.0, 0.5)])
def test_full_size_properties(kind, size, p0, iso):
    P = pkg()
    gen = {"gyroid": P.capi.GEN_GYROID, "marschner_lobb": P.capi.GEN_MARSCHNER_LOBB, "blobs": P.capi.GEN_BLOBS}[kind]
    h = P.capi.Handle(0)
    h.generate(gen, (size, size, size), p0=p0, p1=1.0)
    h._iso = iso
    pts, quads = _run(P, h)
    n, f = pts.shape[0], quads.shape[0]
    assert f > 100000 and n > 100000
    # ids: dense, all used
    assert int(quads.max()) == n - 1
    used = np.zeros(n, bool)
    used[quads.reshape(-1)] = True
    assert used.all()
    # exact lattice positions, unit squares
    assert np.array_equal(pts + 0.5, np.rint(pts + 0.5))
    q = pts[quads.astype(np.int64)]
    edge = np.abs(q - np.roll(q, -1, axis=1)).sum(axis=2)
    assert np.all(edge == 1.0)
    # closed surface: even edge multiplicities, sum = 4F
    a = quads.astype(np.uint64)
    b = np.roll(a, -1, axis=1)
    lo, hi = np.minimum(a, b), np.maximum(a, b)
    keys = (lo << np.uint64(32) | hi).reshape(-1)
    _, mult = np.unique(keys, return_counts=True)
    assert mult.sum() == 4 * f
    assert np.all(mult % 2 == 0) and mult.max() <= 4
    # no empty interior slice (the reference's lookup-plane quirk is not in play)
    vol = h.download_volume()
    inside = (vol >= iso)
    occ = inside.reshape(size, -1).any(axis=1)
    assert occ[1:-1].all() or kind == "marschner_lobb"
    # idempotence
    pts2, quads2 = _run(P, h)
    assert np.array_equal(pts.view(np.uint32), pts2.view(np.uint32)) and np.array_equal(quads, quads2)
    # raster vertex order: a renumbering
    rp, rq = _run(P, h, order=P.capi.ORDER_RASTER)
    assert rp.shape == pts.shape and rq.shape == quads.shape
    key = {}
    view = np.ascontiguousarray(pts).view([("x", "f4"), ("y", "f4"), ("z", "f4")]).reshape(-1)
    order_ref = np.argsort(view, order=("z", "y", "x"))
    assert np.array_equal(np.ascontiguousarray(rp).view(view.dtype).reshape(-1), view[order_ref])  # raster order = sorted corners
    to_ref = order_ref  # raster id k is reference id order_ref[k]
    assert np.array_equal(to_ref[rq.astype(np.int64)], quads.astype(np.int64))
    # z-slabs: 4 slabs, 2-slice halo, concatenation equals the single run
    bounds = np.linspace(0, size, 5).astype(int)
    p = P.capi.default_params()
    p.iso_value, p.generate_triangles, p.project_vertices = iso, 0, 0
    hs, counts = [], []
    for s in range(4):
        z0, z1 = int(bounds[s]), int(bounds[s + 1])
        lo_, hi_ = max(0, z0 - 2), min(size, z1 + 2)
        hh = P.capi.Handle(0)
        hh.generate(gen, (size, size, hi_ - lo_), (size, size, size), lo_, p0, 1.0)
        hh.set_slab(size, lo_, z0, z1)
        counts.append(hh.count(p))
        hs.append(hh)
    pbase, parts_p, parts_c = 0, [], []
    for hh, (np_, nq) in zip(hs, counts):
        hh.set_id_base(pbase, 0)
        hh.emit(4)
        a_, b_, _ = hh.fetch()
        parts_p.append(a_); parts_c.append(b_); pbase += np_
        hh.close()
    assert np.array_equal(np.concatenate(parts_p).view(np.uint32), pts.view(np.uint32))
    assert np.array_equal(np.concatenate(parts_c), quads)
    # anchor: a 24-slice sub-slab of the same bytes through the oracle
    O = oracle()
    z0 = size // 2
    sub = np.ascontiguousarray(vol[z0:z0 + 24])
    ref = O.cuberille(sub, iso, triangles=True, project=True, thr=0.01 if kind == "gyroid" else 0.005)
    hs2 = P.capi.Handle(0)
    hs2.set_volume(sub)
    hs2._iso = iso
    sp, sc = _run(P, hs2, tri=True, proj=True, thr=0.01 if kind == "gyroid" else 0.005)
    assert_mesh_equal(P.Mesh(sp, sc), ref, f"{kind} sub-slab")
    hs2.close()
    h.close()


def test_headline_size_gyroid_1024_on_device():
    """BASELINE.json's headline workload (gyroid 1024^3 float32, P = 128, quads, u32 ids) checked where it lives:
    the mesh stays on the device and torch does the set arithmetic (44 M points, 44 M quads).
    * ids dense and all used; every quad a unit square on the half-integer lattice
    * closed surface (the generator forces the border outside): even edge multiplicities <= 4, sum = 4 F
    * idempotent; 4 z-slabs concatenate to the same bytes
    * a 16-slice sub-slab of the same bytes goes through the oracle, bit for bit (triangles + projection)"""
    import torch
    P = pkg()
    S = 1024
    dev = torch.device("cuda:0")
    p = P.capi.default_params()
    p.iso_value, p.generate_triangles, p.project_vertices = 0.0, 0, 0

    def run(handle):
        n_pts, n_quads = handle.count(p)
        return n_pts, n_quads

    def fetch(handle, n_pts, n_quads):
        pts = torch.empty((n_pts, 3), dtype=torch.float32, device=dev)
        cells = torch.empty((n_quads, 4), dtype=torch.int32, device=dev)
        handle.fetch_into(pts.data_ptr(), cells.data_ptr(), 0, P.capi.MEM_DEVICE)
        return pts, cells

    h = P.capi.Handle(0)
    h.generate(P.capi.GEN_GYROID, (S, S, S), p0=128.0)
    n_pts, n_quads = run(h)
    h.emit(4)
    pts, quads = fetch(h, n_pts, n_quads)
    assert n_pts > 40_000_000 and n_quads > 40_000_000
    q = quads.long()
    assert int(q.max()) == n_pts - 1 and int(q.min()) == 0
    assert bool((torch.bincount(q.reshape(-1), minlength=n_pts) > 0).all())
    assert torch.equal(pts + 0.5, torch.round(pts + 0.5))
    for k in range(4):  # unit squares: consecutive corners differ by one lattice step along one axis
        d = (pts[q[:, k]] - pts[q[:, (k + 1) % 4]]).abs().sum(dim=1)
        assert bool((d == 1.0).all())
        del d
    lo = torch.minimum(q, q.roll(-1, dims=1))
    hi = torch.maximum(q, q.roll(-1, dims=1))
    keys = (lo << 32 | hi).reshape(-1)
    del lo, hi
    _, mult = torch.unique(keys, return_counts=True)
    del keys
    assert int(mult.sum()) == 4 * n_quads and bool((mult % 2 == 0).all()) and int(mult.max()) <= 4
    del mult, q
    # idempotence
    assert run(h) == (n_pts, n_quads)
    h.emit(4)
    pts2, quads2 = fetch(h, n_pts, n_quads)
    assert torch.equal(pts.view(torch.int32), pts2.view(torch.int32)) and torch.equal(quads, quads2)
    del pts2, quads2
    # oracle anchor on a sub-slab of the same bytes
    vol = h.download_volume()
    sub = np.ascontiguousarray(vol[500:516])
    del vol
    h.close()
    O = oracle()
    ref = O.cuberille(sub, 0.0, triangles=True, project=True, thr=0.01)
    hs = P.capi.Handle(0)
    hs.set_volume(sub)
    ps = P.capi.default_params()
    ps.iso_value, ps.surface_distance_threshold = 0.0, 0.01
    hs.run(ps)
    sp, sc, _ = hs.fetch()
    assert_mesh_equal(P.Mesh(sp, sc), ref, "gyroid 1024 sub-slab")
    hs.close()
    # 4 z-slabs, generated per slab with a 2-slice halo
    pbase = cbase = 0
    for s in range(4):
        z0, z1 = S * s // 4, S * (s + 1) // 4
        lo_, hi_ = max(0, z0 - 2), min(S, z1 + 2)
        hh = P.capi.Handle(0)
        hh.generate(P.capi.GEN_GYROID, (S, S, hi_ - lo_), (S, S, S), lo_, 128.0)
        hh.set_slab(S, lo_, z0, z1)
        np_, nq = hh.count(p)
        hh.set_id_base(pbase, cbase)
        hh.emit(4)
        a, b = fetch(hh, np_, nq)
        assert torch.equal(a.view(torch.int32), pts[pbase:pbase + np_].view(torch.int32)), f"slab {s}: points"
        assert torch.equal(b, quads[cbase:cbase + nq]), f"slab {s}: cells"
        pbase += np_
        cbase += nq
        hh.close()
        del a, b
    assert (pbase, cbase) == (n_pts, n_quads)


def test_blobs_1024_triangles_projection_slab_of_the_full_run_matches_the_oracle():
    """BASELINE.json config 5 at one GPU's share (sphere blobs 1024^3 float32, triangles + projection).  The full run
    stays on the device; a 16-slice z-slab of it with the projection halo is (a) run as a slab handle and compared
    with the corresponding section of the full mesh, bytes for bytes, and (b) run through the oracle on the same
    bytes (own range + halo as a stand-alone image: the halo keeps every read of the own range's vertices away from
    the cut), ids shifted by the number of points the full run created before the slab."""
    import torch
    P, O = pkg(), oracle()
    S, Z0, Z1 = 1024, 600, 616
    dev = torch.device("cuda:0")
    prm = P.capi.default_params()
    prm.iso_value, prm.generate_triangles, prm.project_vertices, prm.surface_distance_threshold = 0.5, 1, 1, 0.005
    below, above = P.capi.projection_halo(prm)
    h = P.capi.Handle(0)
    h.generate(P.capi.GEN_BLOBS, (S, S, S), p0=48.0, p1=1.0)
    n_pts, n_cells = h.run(prm)
    assert n_pts > 30_000_000 and n_cells == 2 * h.count(prm)[1]
    h.emit(4)
    pts = torch.empty((n_pts, 3), dtype=torch.float32, device=dev)
    cells = torch.empty((n_cells, 3), dtype=torch.int32, device=dev)
    h.fetch_into(pts.data_ptr(), cells.data_ptr(), 0, P.capi.MEM_DEVICE)
    assert int(cells.max()) == n_pts - 1 and bool(torch.isfinite(pts).all())
    h.close()
    # how many points / cells exist when the raster loop enters slice Z0: a count-only handle over [0, Z0)
    hb = P.capi.Handle(0)
    hb.generate(P.capi.GEN_BLOBS, (S, S, Z0 + above), (S, S, S), 0, 48.0, 1.0)
    hb.set_slab(S, 0, 0, Z0)
    pbase, qbase = hb.count(prm)
    hb.close()
    # (a) the slab as a slab handle of the same image
    lo, hi = Z0 - below, Z1 + above
    hs = P.capi.Handle(0)
    hs.generate(P.capi.GEN_BLOBS, (S, S, hi - lo), (S, S, S), lo, 48.0, 1.0)
    hs.set_slab(S, lo, Z0, Z1)
    np_, nq = hs.count(prm)
    hs.set_id_base(pbase, 2 * qbase)
    hs.emit(4)
    sp, sc, _ = hs.fetch()
    sub = hs.download_volume()
    hs.close()
    assert np.array_equal(sp.view(np.uint32), pts[pbase:pbase + np_].cpu().numpy().view(np.uint32)), "slab points differ from the full run"
    assert np.array_equal(sc.view(np.int32), cells[2 * qbase:2 * qbase + 2 * nq].cpu().numpy()), "slab cells differ from the full run"
    # (b) the oracle on the same bytes
    # (the slab sits at image index (0, 0, lo): the same physical coordinates, so the same float roundings)
    ref = O.cuberille(sub, 0.5, triangles=True, project=True, thr=0.005, region_index=(0, 0, lo))
    pb, cb = ref.points_before_slice, ref.cells_before_slice
    k0, k1 = Z0 - lo, Z1 - lo
    assert int(pb[k1] - pb[k0]) == np_ and int(cb[k1] - cb[k0]) == 2 * nq
    assert np.array_equal(ref.points[int(pb[k0]):int(pb[k1])].view(np.uint32), sp.view(np.uint32)), "projected points differ from the oracle"
    shifted = ref.cells[int(cb[k0]):int(cb[k1])].astype(np.int64) - int(pb[k0]) + pbase
    assert np.array_equal(shifted, sc.astype(np.int64)), "connectivity differs from the oracle"
