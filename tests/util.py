"""Shared helpers of the test-suite."""
from __future__ import annotations

import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

DATA = os.path.join(ROOT, "tests", "data")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pkg():
    return importlib.import_module("midas-journal-740_b200")


def oracle():
    import oracle_py
    return oracle_py


# The reference's known-answer table: Testing/CMakeLists.txt:10-331
# (ctest name, file, iso, expected points, expected cells, triangles, project, max steps);
# every row passes thr 0.2, step 0.24, relax 0.95.
KAT = [
    ("Cuberille_Blob0_00", "blob0", 200, 8, 6, 0, 0, 100),
    ("Cuberille_Blob1_01", "blob1", 200, 12, 10, 0, 0, 100),
    ("Cuberille_Blob2_01", "blob2", 200, 14, 12, 0, 0, 100),
    ("Cuberille_Blob3_01", "blob3", 200, 122, 180, 0, 0, 100),
    ("Cuberille_Blob4_01", "blob4", 200, 2124, 2122, 0, 1, 100),
    ("Cuberille_MarschnerLobb_01", "marschnerlobb", 55, 20524, 22104, 0, 1, 200),
    ("Cuberille_Fuel_01", "fuel", 15, 5302, 5316, 0, 0, 100),
    ("Cuberille_Fuel_02", "fuel", 15, 5302, 5316, 0, 1, 100),
    ("Cuberille_Fuel_03", "fuel", 15, 5302, 10632, 1, 1, 100),
    ("Cuberille_HydrogenAtom_01", "hydrogenAtom", 15, 29880, 29874, 0, 1, 100),
    ("Cuberille_Neghip_01", "neghip", 55, 15146, 15136, 0, 0, 100),
    ("Cuberille_Neghip_02", "neghip", 55, 15146, 15136, 0, 1, 100),
    ("Cuberille_Neghip_03", "neghip", 55, 15146, 30272, 1, 1, 100),
    ("Cuberille_Nucleon_01", "nucleon", 140, 3504, 3500, 0, 0, 100),
    ("Cuberille_Nucleon_02", "nucleon", 140, 3504, 3500, 0, 1, 100),
    ("Cuberille_Nucleon_03", "nucleon", 140, 3504, 7000, 1, 1, 100),
    ("Cuberille_Silicium_01", "silicium", 85, 20036, 20024, 0, 0, 100),
    ("Cuberille_Silicium_02", "silicium", 85, 20036, 20024, 0, 1, 100),
    ("Cuberille_Silicium_03", "silicium", 85, 20036, 40048, 1, 1, 100),
]
KAT_ARGS = dict(thr=0.2, step=0.24, relax=0.95)


def read_fixture(name: str):
    return pkg().read_mha(os.path.join(DATA, name + ".mha"))


def gyroid(n, period=16.0, dtype=np.float32, border=-2.0):
    """gyroid field on integer voxel coordinates, shape (nz, ny, nx), outermost layer forced outside (`border`)"""
    nz, ny, nx = (n, n, n) if np.isscalar(n) else n
    k = np.float32(2.0 * np.pi / period)
    z, y, x = np.meshgrid(np.arange(nz, dtype=np.float32), np.arange(ny, dtype=np.float32),
                          np.arange(nx, dtype=np.float32), indexing="ij")
    g = np.sin(k * x) * np.cos(k * y) + np.sin(k * y) * np.cos(k * z) + np.sin(k * z) * np.cos(k * x)
    g = g.astype(dtype)
    if border is not False:  # (False: the field runs up to the image border, for image_border_faces)
        g[0], g[-1] = border, border
        g[:, 0], g[:, -1] = border, border
        g[:, :, 0], g[:, :, -1] = border, border
    return np.ascontiguousarray(g)


def random_volume(shape, dtype, seed, fill=0.5):
    """iid noise: a worst case for the ownership rule (every corner configuration, inside voxels on
    the image border).  Returns (volume, iso)."""
    rng = np.random.default_rng(seed)
    dt = np.dtype(dtype)
    if np.issubdtype(dt, np.integer):
        info = np.iinfo(dt)
        lo, hi = max(info.min, -1000), min(info.max, 1000)
        vol = rng.integers(lo, hi + 1, size=shape, dtype=np.int64).astype(dt)
        iso = int(lo + (hi - lo) * (1.0 - fill))
    else:
        vol = rng.random(shape).astype(dt)
        iso = float(np.float32(1.0 - fill))
    return np.ascontiguousarray(vol), iso


def smooth_volume(shape, dtype, seed, scale=60.0):
    """band-limited random field (sum of a few random plane waves), scaled into the dtype's range"""
    rng = np.random.default_rng(seed)
    nz, ny, nx = shape
    z, y, x = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    f = np.zeros(shape, np.float64)
    for _ in range(6):
        k = rng.normal(size=3) * 0.35
        f += np.cos(k[0] * x + k[1] * y + k[2] * z + rng.uniform(0, 6.28))
    f = f / 6.0 * scale + scale  # in [0, 2*scale]
    dt = np.dtype(dtype)
    vol = np.rint(f).astype(dt) if np.issubdtype(dt, np.integer) else f.astype(dt)
    return np.ascontiguousarray(vol), (int(scale) if np.issubdtype(dt, np.integer) else float(scale))


def run_filter(img_or_vol, iso, *, triangles, project, cell_data=False, thr=0.5, step=-1.0, relax=0.95, max_steps=50,
               id_bytes=4, spacing=(1.0, 1.0, 1.0), origin=(0.0, 0.0, 0.0), raster_order=False, border_faces=False, method=0):
    """Drive the CUDA path through the filter mirror, with the reference driver's call sequence
    (Testing/CuberilleTest01.cxx:144-162)."""
    P = pkg()
    img = img_or_vol if isinstance(img_or_vol, P.Image) else P.Image(img_or_vol, spacing, origin)
    f = P.CuberilleImageToMeshFilter.New(id_bytes=id_bytes)
    f.SetInput(img)
    f.SetIsoSurfaceValue(iso)
    f.SetGenerateTriangleFaces(triangles)
    f.SetProjectVerticesToIsoSurface(project)
    f.SetSavePixelAsCellData(cell_data)
    f.SetRasterVertexOrder(raster_order)
    f.SetImageBorderFaces(border_faces)
    f.SetProjectionMethod(method)
    f.SetProjectVertexSurfaceDistanceThreshold(thr)
    if step >= 0:
        f.SetProjectVertexStepLength(step)
    f.SetProjectVertexStepLengthRelaxationFactor(relax)
    f.SetProjectVertexMaximumNumberOfSteps(max_steps)
    f.Update()
    return f.GetOutput()


def assert_mesh_equal(mesh, ref, what=""):
    """bit-exact: vertex count, connectivity in the reference's order, positions (float bits)"""
    assert mesh.points.shape == ref.points.shape, f"{what}: #points {mesh.points.shape} vs {ref.points.shape}"
    assert mesh.cells.shape == ref.cells.shape, f"{what}: #cells {mesh.cells.shape} vs {ref.cells.shape}"
    assert np.array_equal(mesh.cells.astype(np.uint64), ref.cells), f"{what}: connectivity differs"
    a, b = mesh.points.view(np.uint32), ref.points.view(np.uint32)
    if not np.array_equal(a, b):
        bad = np.nonzero((a != b).any(axis=1))[0]
        raise AssertionError(f"{what}: {bad.size} points differ, first {bad[0]}: {mesh.points[bad[0]]} vs {ref.points[bad[0]]}")
    if ref.cell_data is not None:
        assert mesh.cell_data is not None and np.array_equal(mesh.cell_data, ref.cell_data), f"{what}: cell data differs"


def assert_mesh_equal_up_to_vertex_order(mesh, ref, mesh_unprojected, ref_unprojected, what=""):
    """raster vertex order: same points, cells and cell order as the reference after renumbering the vertices.
    The renumbering is recovered from the UNPROJECTED meshes (lattice positions are unique and exact) and then
    applied to the meshes under test (vertex numbering does not depend on the projection)."""
    assert mesh.points.shape == ref.points.shape, f"{what}: #points {mesh.points.shape} vs {ref.points.shape}"
    assert mesh.cells.shape == ref.cells.shape, f"{what}: #cells"
    key = {p.tobytes(): i for i, p in enumerate(ref_unprojected.points)}
    assert len(key) == ref_unprojected.points.shape[0]
    to_ref = np.array([key[p.tobytes()] for p in mesh_unprojected.points], np.int64)
    assert np.array_equal(np.sort(to_ref), np.arange(to_ref.size)), f"{what}: not a renumbering"
    # raster order = sorted by (z, y, x) of the lattice corner
    o = np.lexsort((mesh_unprojected.points[:, 0], mesh_unprojected.points[:, 1], mesh_unprojected.points[:, 2]))
    assert np.array_equal(o, np.arange(o.size)), f"{what}: vertices are not in corner raster order"
    assert np.array_equal(mesh.points.view(np.uint32), ref.points[to_ref].view(np.uint32)), f"{what}: positions differ"
    assert np.array_equal(to_ref[mesh.cells.astype(np.int64)], ref.cells.astype(np.int64)), f"{what}: connectivity differs"
    if ref.cell_data is not None:
        assert np.array_equal(mesh.cell_data, ref.cell_data)
