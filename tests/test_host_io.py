"""CPU: the file formats either side of the filter: MetaImage reader / writer and the legacy-VTK
polydata writer that stand in for itk::ImageFileReader / itk::VTKPolyDataWriter (Testing/CuberilleTest01.cxx:113-117,
180-187)."""
import numpy as np
import pytest

from util import pkg, read_fixture


@pytest.mark.parametrize("dtype", [np.uint8, np.int16, np.uint16, np.int32, np.float32, np.float64])
@pytest.mark.parametrize("compress", [True, False])
def test_write_mha_round_trip(tmp_path, dtype, compress):
    P = pkg()
    rng = np.random.default_rng(3)
    data = (rng.random((5, 7, 9)) * 100).astype(dtype)
    # an oriented image: rotation about z by 30 degrees combined with a flipped y axis (rows: itk direction D[i][j])
    c, s = np.cos(np.pi / 6), np.sin(np.pi / 6)
    direction = (c, s, 0.0, -s, -c, 0.0, 0.0, 0.0, 1.0)
    img = P.Image(data, (0.5, 1.25, 2.0), (-3.0, 4.5, 0.1), direction)
    path = str(tmp_path / "v.mha")
    P.write_mha(path, img, compress=compress)
    back = P.read_mha(path)
    assert back.data.dtype == np.dtype(dtype) and np.array_equal(back.data, data)
    assert back.spacing == img.spacing and back.origin == img.origin
    assert back.direction == tuple(float(v) for v in direction)  # exact: written with repr()
    # the file lists the axis vectors (columns of D)
    assert back.meta["TransformMatrix"].split()[:3] == [repr(float(c)), repr(float(-s)), repr(0.0)]


def test_read_mha_reference_fixture_is_identity_and_rewrites_identically(tmp_path):
    P = pkg()
    img = read_fixture("fuel")
    assert img.data.shape == (68, 68, 68) and img.data.dtype == np.uint8
    assert img.direction == (1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0) and img.spacing == (1.0, 1.0, 1.0)
    path = str(tmp_path / "fuel2.mha")
    P.write_mha(path, img)
    again = P.read_mha(path)
    assert np.array_equal(again.data, img.data) and again.origin == img.origin


def _read_vtk(path):
    lines = open(path).read().split("\n")
    assert lines[0].startswith("# vtk DataFile") and lines[2] == "ASCII" and lines[3] == "DATASET POLYDATA"
    i = next(k for k, l in enumerate(lines) if l.startswith("POINTS"))
    n = int(lines[i].split()[1])
    pts = np.array([[float(v) for v in lines[i + 1 + k].split()] for k in range(n)], np.float32).reshape(n, 3)
    j = next(k for k, l in enumerate(lines) if l.startswith("POLYGONS"))
    m, total = int(lines[j].split()[1]), int(lines[j].split()[2])
    rows = [[int(v) for v in lines[j + 1 + k].split()] for k in range(m)]
    cd = None
    if any(l.startswith("CELL_DATA") for l in lines):
        c = next(k for k, l in enumerate(lines) if l.startswith("LOOKUP_TABLE"))
        cd = np.array([float(lines[c + 1 + k]) for k in range(m)])
    return pts, rows, total, cd


@pytest.mark.parametrize("k", [3, 4])
def test_write_vtk_polydata_round_trip(tmp_path, k):
    P, rng = pkg(), np.random.default_rng(k)
    pts = (rng.random((50, 3)) * 37 - 5).astype(np.float32)
    cells = rng.integers(0, 50, size=(31, k)).astype(np.uint32)
    cd = rng.integers(0, 255, size=31).astype(np.uint8)
    path = str(tmp_path / "m.vtk")
    P.write_vtk_polydata(path, pts, cells, cd)
    p2, rows, total, cd2 = _read_vtk(path)
    assert np.array_equal(p2.view(np.uint32), pts.view(np.uint32))  # %.9g round-trips float32
    assert total == 31 * (k + 1) and all(r[0] == k for r in rows)
    assert np.array_equal(np.array([r[1:] for r in rows], np.uint32), cells)
    assert np.array_equal(cd2.astype(np.uint8), cd)


def test_write_vtk_polydata_empty_mesh(tmp_path):
    P = pkg()
    path = str(tmp_path / "e.vtk")
    P.write_vtk_polydata(path, np.zeros((0, 3), np.float32), np.zeros((0, 4), np.uint32))
    p2, rows, total, cd = _read_vtk(path)
    assert p2.shape == (0, 3) and rows == [] and total == 0 and cd is None
