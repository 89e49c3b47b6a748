"""CPU: the oracle against every known-answer test the reference holds for this path
(Testing/CMakeLists.txt:10-331; the driver asserts only #points / #cells,
Testing/CuberilleTest01.cxx:193-204) plus self-consistency of the restatement."""
import numpy as np
import pytest

from util import KAT, KAT_ARGS, gyroid, oracle, random_volume, read_fixture, smooth_volume


@pytest.mark.parametrize("row", KAT, ids=[r[0] for r in KAT])
def test_reference_known_answers(row):
    name, fixture, iso, exp_points, exp_cells, tri, proj, max_steps = row
    O = oracle()
    img = read_fixture(fixture)
    assert img.data.dtype == np.uint8 and img.spacing == (1.0, 1.0, 1.0)
    for mode in (O.LITERAL, O.CLOSED_FORM):
        m = O.cuberille(img.data, iso, triangles=tri, project=proj, mode=mode, max_steps=max_steps, **KAT_ARGS)
        assert m.points.shape[0] == exp_points
        assert m.cells.shape[0] == exp_cells
        assert np.isfinite(m.points).all()


@pytest.mark.parametrize("seed", range(24))
def test_literal_lookup_equals_closed_form_on_noise(seed):
    """the same on iid noise (every corner configuration, inside voxels on the image border, ragged shapes): the
    verbatim loop with its two lookup planes and the closed form number every vertex and cell alike, as long as no
    slice between two occupied ones is empty (the reference's lookup-plane rotation only advances on inside voxels)"""
    O = oracle()
    rng = np.random.default_rng(1000 + seed)
    shape = tuple(int(v) for v in rng.integers(1, 14, size=2)) + (int(rng.integers(1, 70)),)
    fill = rng.uniform(0.15, 0.85)
    vol = (rng.random(shape) < fill).astype(np.uint8) * 200
    for z in range(shape[0]):  # one inside voxel per slice
        vol[z, rng.integers(0, shape[1]), rng.integers(0, shape[2])] = 200
    for tri in (False, True):
        a = O.cuberille(vol, 100, triangles=tri, project=False, mode=O.LITERAL)
        b = O.cuberille(vol, 100, triangles=tri, project=False, mode=O.CLOSED_FORM)
        assert a.points.shape == b.points.shape and a.cells.shape == b.cells.shape, (shape, fill)
        assert np.array_equal(a.cells, b.cells), (shape, fill)
        assert np.array_equal(a.points.view(np.uint32), b.points.view(np.uint32)), (shape, fill)


@pytest.mark.parametrize("fixture,iso", [("fuel", 15), ("nucleon", 140), ("blob3", 200)])
def test_literal_lookup_equals_closed_form(fixture, iso):
    """the two-plane std::map emulation (txx:155-161,186-191) and the first-touch closed form give
    the same ids when no interior z-slice is empty"""
    O = oracle()
    img = read_fixture(fixture)
    a = O.cuberille(img.data, iso, triangles=True, project=True, mode=O.LITERAL, max_steps=100, **KAT_ARGS)
    b = O.cuberille(img.data, iso, triangles=True, project=True, mode=O.CLOSED_FORM, max_steps=100, **KAT_ARGS)
    assert np.array_equal(a.cells, b.cells)
    assert np.array_equal(a.points.view(np.uint32), b.points.view(np.uint32))


def test_unprojected_quads_are_unit_squares_and_order_is_first_touch():
    O = oracle()
    vol = gyroid(24, 12.0)
    m = O.cuberille(vol, 0.0, triangles=False, project=False)
    p = m.points[m.cells.astype(np.int64)]  # (n, 4, 3)
    edges = np.linalg.norm(p - np.roll(p, -1, axis=1), axis=2)
    assert np.all(edges == 1.0)
    # every coordinate is index - 0.5 exactly
    assert np.all((m.points + 0.5) == np.rint(m.points + 0.5))
    # every id is used, and a vertex is created no later than the voxel that first uses it: the largest id
    # seen up to any cell never exceeds (distinct ids so far) + 7 (the other corners of that voxel)
    flat = m.cells.reshape(-1).astype(np.int64)
    assert np.array_equal(np.unique(flat), np.arange(m.points.shape[0]))
    first = np.full(m.points.shape[0], flat.size, np.int64)
    np.minimum.at(first, flat, np.arange(flat.size))
    order = np.argsort(first, kind="stable")
    assert np.all(np.abs(order - np.arange(order.size)) <= 7)


def test_triangles_are_two_per_quad_and_share_the_quad_vertices():
    O = oracle()
    vol = gyroid(20, 10.0)
    q = O.cuberille(vol, 0.0, triangles=False, project=True, thr=0.01)
    t = O.cuberille(vol, 0.0, triangles=True, project=True, thr=0.01)
    assert t.cells.shape[0] == 2 * q.cells.shape[0]
    assert np.array_equal(q.points.view(np.uint32), t.points.view(np.uint32))
    tt = t.cells.reshape(-1, 6)
    for quad, tri in zip(q.cells[:500], tt[:500]):
        assert set(quad) == set(tri)
        assert (list(tri) == [quad[0], quad[1], quad[3], quad[1], quad[2], quad[3]]
                or list(tri) == [quad[0], quad[1], quad[2], quad[0], quad[2], quad[3]])


def test_border_voxels_make_no_faces():
    """ZeroFluxNeumann edge replicate: an inside voxel on the image border has no face there (txx:167, h:55-57)"""
    O = oracle()
    vol = np.full((4, 5, 6), 10, np.uint8)  # everything inside
    m = O.cuberille(vol, 5, triangles=False, project=False)
    assert m.points.shape[0] == 0 and m.cells.shape[0] == 0
    vol[1:3, 1:4, 1:5] = 0  # a hole: faces point into it
    m = O.cuberille(vol, 5, triangles=False, project=False)
    assert m.cells.shape[0] == 2 * (3 * 4 + 2 * 3 + 2 * 4)


def test_empty_slice_quirk_is_reproduced_by_literal_mode_only():
    """lastZ only advances on inside voxels, so an empty z-slice between two occupied
    ones makes the literal loop reuse the corner plane: 12 points instead of 16."""
    O = oracle()
    vol = np.zeros((10, 5, 5), np.uint8)
    vol[5, 2, 2] = 9
    vol[7, 2, 2] = 9
    a = O.cuberille(vol, 5, triangles=False, project=False, mode=O.LITERAL)
    b = O.cuberille(vol, 5, triangles=False, project=False, mode=O.CLOSED_FORM)
    assert a.points.shape[0] == 12 and b.points.shape[0] == 16
    assert a.cells.shape[0] == 12 and b.cells.shape[0] == 12


def test_nan_pixels_count_as_inside():
    O = oracle()
    vol = np.zeros((5, 5, 5), np.float32)
    vol[2, 2, 2] = np.nan  # !(nan < iso) -> inside, never "outside"
    m = O.cuberille(vol, 0.5, triangles=False, project=False)
    assert m.cells.shape[0] == 6 and m.points.shape[0] == 8


@pytest.mark.parametrize("dtype", [np.uint8, np.int8, np.uint16, np.int16, np.uint32, np.int32, np.float32, np.float64])
def test_all_pixel_types(dtype):
    O = oracle()
    vol, iso = random_volume((9, 10, 11), dtype, 5)
    a = O.cuberille(vol, iso, triangles=False, project=False, mode=O.LITERAL)
    b = O.cuberille(vol, iso, triangles=False, project=False, mode=O.CLOSED_FORM)
    assert a.cells.shape[0] > 0
    assert np.array_equal(a.cells, b.cells) and np.array_equal(a.points, b.points)
    bits = O.classify(vol, iso)
    inside = ~(vol < np.asarray(iso).astype(vol.dtype))
    unpacked = ((bits[:, :, :, None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(vol.shape[0], vol.shape[1], -1)
    assert np.array_equal(unpacked[:, :, : vol.shape[2]].astype(bool), inside)


def test_projection_moves_vertices_towards_the_iso_value():
    O = oracle()
    img = read_fixture("fuel")
    m0 = O.cuberille(img.data, 15, triangles=False, project=False)
    m1 = O.cuberille(img.data, 15, triangles=False, project=True, max_steps=100, **KAT_ARGS)
    v0, _ = O.sample(img.data, m0.points.astype(np.float64))
    v1, _ = O.sample(img.data, m1.points.astype(np.float64))
    assert np.mean(np.abs(v1 - 15) < 0.2) > 0.95
    assert np.mean(np.abs(v1 - 15)) < np.mean(np.abs(v0 - 15))
    assert np.max(np.linalg.norm(m1.points - m0.points, axis=1)) < 5.0


def test_image_border_faces_close_the_mesh():
    """opt-in: a neighbour outside the image is outside the surface (what padding with one outside layer gives);
    the reference never makes a face on the image border (h:55-57, txx:133)"""
    O = oracle()
    box = np.full((3, 4, 5), 9, np.uint8)  # everything inside
    for mode in (O.LITERAL, O.CLOSED_FORM):
        assert O.cuberille(box, 5, triangles=False, project=False, mode=mode).cells.shape[0] == 0
        m = O.cuberille(box, 5, triangles=False, project=False, mode=mode, border_faces=True)
        assert m.cells.shape[0] == 2 * (3 * 4 + 4 * 5 + 3 * 5) and m.points.shape[0] == 6 * 5 * 4 - 4 * 3 * 2
    # equals the default run on the explicitly padded volume, up to the shift of the origin
    rng = np.random.default_rng(11)
    vol = rng.integers(0, 256, size=(6, 7, 9), dtype=np.uint8)
    padded = np.zeros((8, 9, 11), np.uint8)
    padded[1:-1, 1:-1, 1:-1] = vol
    a = O.cuberille(vol, 128, triangles=True, project=False, mode=O.CLOSED_FORM, border_faces=True)
    b = O.cuberille(padded, 128, triangles=True, project=False, mode=O.CLOSED_FORM, origin=(-1.0, -1.0, -1.0))
    assert np.array_equal(a.cells, b.cells) and np.array_equal(a.points, b.points)


def test_oracle_params_struct_layout():
    """oracle_py._Params mirrors orc_params field for field (a drift would silently change what the oracle computes)"""
    import ctypes as C
    import os
    import re
    O = oracle()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "oracle", "cuberille_oracle.cpp")).read()
    body = text[text.index("struct orc_params {"):]
    body = body[:body.index("};")]
    names = re.findall(r"^\s*(?:double|int32_t|uint32_t|int64_t)\s+(\w+)", body, re.M)
    assert names == [n for n, _ in O._Params._fields_]
    assert C.sizeof(O._Params) == 8 + 4 * 4 + 3 * 8 + 2 * 4 + 3 * 8 + 9 * 8 + 2 * 4


# ---- oriented images: the oracle's restatement of ITK's direction-matrix semantics ---------------

def test_oracle_identity_direction_is_the_non_oriented_image():
    O = oracle()
    vol, iso = smooth_volume((12, 14, 16), np.uint8, seed=2)
    kw = dict(triangles=True, project=True, thr=0.05, spacing=(0.5, 1.0, 2.0), origin=(1.0, -2.0, 3.0))
    a = O.cuberille(vol, iso, **kw)
    b = O.cuberille(vol, iso, direction=(1, 0, 0, 0, 1, 0, 0, 0, 1), **kw)
    assert np.array_equal(a.points.view(np.uint32), b.points.view(np.uint32)) and np.array_equal(a.cells, b.cells)


def test_oracle_flipped_axis_positions():
    """TransformIndexToPhysicalPoint with D = diag(-1, 1, 1) into a float point, then the reference's axis-aligned
    half-spacing shift (txx:266-270), restated in numpy"""
    O = oracle()
    vol, iso = smooth_volume((9, 10, 11), np.uint8, seed=4)
    sp, og = (0.7, 1.3, 2.1), (0.1, -5.25, 3.0)
    D = (-1.0, 0, 0, 0, 1.0, 0, 0, 0, 1.0)
    plain = O.cuberille(vol, iso, triangles=False, project=False, mode=O.CLOSED_FORM)
    m = O.cuberille(vol, iso, triangles=False, project=False, mode=O.CLOSED_FORM, spacing=sp, origin=og, direction=D)
    assert np.array_equal(m.cells, plain.cells)
    idx = np.rint(plain.points + 0.5).astype(np.int64)   # lattice corner indices
    M = np.array(D, np.float64).reshape(3, 3) * np.array(sp)[None, :]
    exp = np.empty_like(plain.points)
    for a in range(3):
        p = np.full(idx.shape[0], np.float32(og[a]), np.float32)
        for j in range(3):
            p = (p.astype(np.float64) + M[a, j] * idx[:, j].astype(np.float64)).astype(np.float32)
        exp[:, a] = (p.astype(np.float64) - sp[a] / 2.0).astype(np.float32)
    assert np.array_equal(m.points.view(np.uint32), exp.view(np.uint32))


def test_oracle_axis_permutation_is_equivariant():
    """isotropic spacing, origin 0: with a permutation matrix as direction every operation of the oriented path is the
    non-oriented one up to added zeros, so the mesh is the permuted mesh, bit for bit - projection included"""
    O = oracle()
    vol, iso = smooth_volume((14, 15, 16), np.float32, seed=8)
    D = np.array([[0, 1, 0], [0, 0, 1], [1, 0, 0]], np.float64)   # physical axis i = index axis perm[i]
    kw = dict(triangles=True, project=True, thr=0.02, step=0.24, relax=0.95, max_steps=60, spacing=(2.0, 2.0, 2.0))
    a = O.cuberille(vol, iso, **kw)
    b = O.cuberille(vol, iso, direction=D.reshape(9), **kw)
    assert np.array_equal(a.cells, b.cells)
    assert np.array_equal(b.points.view(np.uint32), (a.points @ D.T.astype(np.float32)).view(np.uint32))


@pytest.mark.parametrize("method", [1, 2])
def test_oracle_alternate_projections_keep_counts_and_move_towards_the_surface(method):
    """USE_ADVANCED_PROJECTION / USE_LINESEARCH_PROJECTION (txx:340-438): same counts and connectivity classes as the
    default branch (projection never changes counts); the vertices end closer to the iso value than they started"""
    O = oracle()
    img = read_fixture("fuel")
    kw = dict(thr=0.2, step=0.24, relax=0.95, max_steps=100)
    flat = O.cuberille(img.data, 15, triangles=False, project=False)
    m = O.cuberille(img.data, 15, triangles=False, project=True, method=method, **kw)
    assert m.points.shape == flat.points.shape == (5302, 3) and np.array_equal(m.cells, flat.cells)
    v0, _ = O.sample(img.data, flat.points.astype(np.float64))
    v1, _ = O.sample(img.data, m.points.astype(np.float64))
    assert np.abs(v1 - 15).mean() < 0.5 * np.abs(v0 - 15).mean()
    assert np.isfinite(m.points).all()
