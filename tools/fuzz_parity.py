"""Randomised parity run: random shapes, pixel types, fills and option combinations, CUDA path against the oracle
(bit-exact).  Usage: python tools/fuzz_parity.py [seconds] [seed] [big]   (big: long rows / many rows: several
sweep tiles, several 32-word segments per row, several packed-classify tasks per row)"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
from util import assert_mesh_equal, assert_mesh_equal_up_to_vertex_order, oracle, pkg, random_volume, run_filter, smooth_volume

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
big = len(sys.argv) > 3
rng = np.random.default_rng(seed)
P, O = pkg(), oracle()
dtypes = [np.uint8, np.int8, np.uint16, np.int16, np.uint32, np.int32, np.float32, np.float64]
t0, n = time.time(), 0
while time.time() - t0 < budget:
    if big:
        shape = (int(rng.integers(1, 10)), int(rng.integers(1, 70)), int(rng.choice([512, 544, 1024, 1056, 1152, 2048, 2080])) if rng.random() < 0.5 else int(rng.integers(300, 2300)))
    else:
      shape = (int(rng.integers(1, 28)), int(rng.integers(1, 36)), int(rng.choice([1, 2, 31, 32, 33, 64, 65, 96, 128, 130, 192, 256])) if rng.random() < 0.5 else int(rng.integers(1, 140)))
    dt = dtypes[int(rng.integers(len(dtypes)))]
    smooth = rng.random() < 0.4 and min(shape) >= 4 and dt not in (np.int8,)
    if smooth:
        vol, iso = smooth_volume(shape, dt, seed=int(rng.integers(1 << 30)))
    else:
        vol, iso = random_volume(shape, dt, seed=int(rng.integers(1 << 30)), fill=float(rng.choice([0.05, 0.3, 0.5, 0.7, 0.95])))
    tri, proj = bool(rng.integers(2)), bool(rng.integers(2)) and smooth
    cd, border, raster = bool(rng.integers(2)), bool(rng.integers(2)), rng.random() < 0.25
    idb = 8 if rng.random() < 0.3 else 4
    geo = dict(spacing=tuple(float(x) for x in rng.choice([0.5, 1.0, 1.25, 2.0], 3)), origin=tuple(float(x) for x in rng.integers(-5, 6, 3)))
    if rng.random() < 0.25:  # an oriented image: axis flips / permutations (exact) or a random rotation
        if rng.random() < 0.5:
            Dm = np.zeros((3, 3)); perm = rng.permutation(3)
            for i_ in range(3):
                Dm[i_, perm[i_]] = rng.choice([-1.0, 1.0])
        else:
            Dm = np.linalg.qr(rng.normal(size=(3, 3)))[0]
        geo["direction"] = tuple(float(x) for x in Dm.reshape(9))
        raster = False  # (the raster-order check sorts by physical coordinates)
    ridx = tuple(int(x) for x in rng.integers(-4, 5, 3)) if rng.random() < 0.3 else (0, 0, 0)
    method = int(rng.choice([0, 0, 1, 2])) if proj else 0   # the reference's alternative projection branches (txx:340-438)
    what = f"#{n} {np.dtype(dt).name} {shape} smooth={smooth} tri={tri} proj={proj} cd={cd} border={border} raster={raster} ids={idb} ridx={ridx} method={method} {geo}"
    kw = dict(triangles=tri, project=proj, cell_data=cd, thr=0.02, method=method)
    ref = O.cuberille(vol, iso, mode=O.CLOSED_FORM, border_faces=border, region_index=ridx, **kw, **geo)
    img = P.Image(vol, geo["spacing"], geo["origin"]); img.region_index = ridx
    if "direction" in geo:
        img.direction = geo["direction"]
    try:
        if raster:
            mesh = run_filter(img, iso, border_faces=border, raster_order=True, **kw)
            kw0 = dict(kw, project=False)
            mesh0 = run_filter(img, iso, border_faces=border, raster_order=True, **kw0)
            ref0 = O.cuberille(vol, iso, mode=O.CLOSED_FORM, border_faces=border, region_index=ridx, **kw0, **geo)
            if ref.points.shape[0]:
                assert_mesh_equal_up_to_vertex_order(mesh, ref, mesh0, ref0, what)
            else:
                assert mesh.points.shape[0] == 0 and mesh.cells.shape[0] == 0
        else:
            mesh = run_filter(img, iso, border_faces=border, id_bytes=idb, **kw)
            assert_mesh_equal(mesh, ref, what)
            # the same through z-slabs
            nz = shape[0]
            if nz >= 4 and rng.random() < 0.5:
                cuts = sorted(set([0, nz] + [int(x) for x in rng.integers(1, nz, int(rng.integers(1, 4)))]))
                prm = P.capi.default_params()
                prm.iso_value, prm.generate_triangles, prm.project_vertices = float(iso), int(tri), int(proj)
                prm.save_pixel_as_cell_data, prm.image_border_faces, prm.surface_distance_threshold = int(cd), int(border), 0.02
                prm.projection_method = method
                pts, cells, cds, pb, cb = [], [], [], 0, 0
                # a projected vertex travels up to step / (1 - relax) = 5 * max spacing, i.e. many slices when the z
                # spacing is the small one: give the slabs the whole image as halo then
                halo = nz if proj else 2
                dirn = geo.get("direction")
                for z0, z1 in zip(cuts[:-1], cuts[1:]):
                    lo, hi = max(0, z0 - halo), min(nz, z1 + halo)
                    h = P.capi.Handle(0)
                    h.set_volume(vol[lo:hi], geo["spacing"], geo["origin"], dirn); h.set_region_index(ridx); h.set_slab(nz, lo, z0, z1)
                    a, b = h.count(prm)
                    if rng.random() < 0.5:
                        h.emit_vertices()
                    h.set_id_base(pb, cb); h.emit(idb)
                    x, y, z = h.fetch(cd)
                    pts.append(x); cells.append(y); cds.append(z)
                    pb += a; cb += b * (2 if tri else 1); h.close()
                m2 = P.Mesh(np.concatenate(pts), np.concatenate(cells), np.concatenate(cds) if cd else None)
                assert_mesh_equal(m2, ref, what + f" slabs {cuts}")
            # the same through slabs.run_streamed (host volume in, host mesh out, several handles / streams)
            if nz >= 4 and not cd and "direction" not in geo and rng.random() < 0.2:
                import torch
                n_sl, n_h, resident = int(rng.integers(2, min(nz, 6) + 1)), int(rng.integers(2, 4)), bool(rng.integers(2))
                vt = torch.from_numpy(vol).pin_memory()
                vpc = 3 if tri else 4
                pts_h = torch.zeros((ref.points.shape[0] + 4, 3), dtype=torch.float32).pin_memory()
                cells_h = torch.zeros((ref.cells.shape[0] + 4, vpc), dtype=torch.int32).pin_memory()
                prm = P.capi.default_params()
                prm.iso_value, prm.generate_triangles, prm.project_vertices = float(iso), int(tri), int(proj)
                prm.image_border_faces, prm.surface_distance_threshold = int(border), 0.02
                prm.projection_method = method
                kw2 = {}
                if resident:
                    sts = [torch.cuda.Stream() for _ in range(n_h)]
                    hs = [P.capi.Handle(0, st.cuda_stream) for st in sts]
                    kw2 = dict(device_volume=torch.empty(vol.nbytes, dtype=torch.uint8, device="cuda"), streams=sts)
                else:
                    hs = [P.capi.Handle(0) for _ in range(n_h)]
                if ridx == (0, 0, 0):  # (run_streamed has no region index parameter)
                    a, b = P.slabs.run_streamed(hs, vt.data_ptr(), vol.dtype, (shape[2], shape[1], shape[0]), prm, n_sl, pts_h.data_ptr(),
                                                cells_h.data_ptr(), halo=nz if proj else 2, spacing=geo["spacing"], origin=geo["origin"], **kw2)
                    assert (a, b) == (ref.points.shape[0], ref.cells.shape[0]), what + " streamed counts"
                    assert_mesh_equal(P.Mesh(pts_h.numpy()[:a], cells_h.numpy()[:b].view(np.uint32)), ref, what + f" streamed {n_sl}/{n_h}/{resident}")
                for h in hs:
                    h.close()
    except Exception:
        print("FAILED:", what, flush=True)
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        np.save(os.path.join(ROOT, "gpurun_out", "fuzz_fail.npy"), vol)
        for t2 in (False, True):
            for c2 in (False, True):
                r2 = O.cuberille(vol, iso, mode=O.CLOSED_FORM, border_faces=border, region_index=ridx, triangles=t2, project=False, cell_data=c2, **geo)
                m2 = run_filter(img, iso, border_faces=border, triangles=t2, project=False, cell_data=c2)
                same_c = m2.cells.shape == r2.cells.shape and np.array_equal(m2.cells.astype(np.uint64), r2.cells)
                same_p = m2.points.shape == r2.points.shape and np.array_equal(m2.points, r2.points)
                print(f"   tri={t2} cd={c2}: cells equal {same_c}, points equal {same_p}, n {m2.points.shape[0]} vs {r2.points.shape[0]}, cells {m2.cells.shape[0]} vs {r2.cells.shape[0]}, iso {iso}", flush=True)
                if not same_c and m2.cells.shape == r2.cells.shape:
                    bad = np.nonzero((m2.cells.astype(np.uint64) != r2.cells).any(axis=1))[0]
                    print("      first bad cells", bad[:5], m2.cells[bad[0]], r2.cells[bad[0]], "n bad", bad.size)
        raise
    n += 1
print(f"fuzz ok: {n} cases in {time.time() - t0:.0f} s (seed {seed})")
