"""ctypes binding of the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs import this module; the product package
never does (see the header of oracle/cuberille_oracle.cpp).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

DTYPE_CODES = {
    np.dtype(np.uint8): 0, np.dtype(np.int8): 1, np.dtype(np.uint16): 2, np.dtype(np.int16): 3,
    np.dtype(np.uint32): 4, np.dtype(np.int32): 5, np.dtype(np.float32): 6, np.dtype(np.float64): 7,
}

LITERAL = 0
CLOSED_FORM = 1
PROJECT_DEFAULT, PROJECT_ADVANCED, PROJECT_LINESEARCH = 0, 1, 2


class _Params(C.Structure):
    _fields_ = [
        ("iso_value", C.c_double),
        ("generate_triangles", C.c_int32),
        ("project_vertices", C.c_int32),
        ("save_pixel_as_cell_data", C.c_int32),
        ("mode", C.c_int32),
        ("surface_distance_threshold", C.c_double),
        ("step_length", C.c_double),
        ("step_relaxation", C.c_double),
        ("max_steps", C.c_uint32),
        ("image_border_faces", C.c_uint32),
        ("region_index", C.c_int64 * 3),
        ("direction", C.c_double * 9),
        ("projection_method", C.c_int32),
        ("reserved", C.c_int32),
    ]


def build(force: bool = False) -> str:
    """Compile liboracle.so with the committed Makefile (g++ only)."""
    src = os.path.join(_HERE, "cuberille_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_cuberille.restype = C.c_void_p
        L.orc_cuberille.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_double),
                                    C.POINTER(C.c_double), C.POINTER(_Params)]
        for name, rt in [("orc_num_points", C.c_uint64), ("orc_num_cells", C.c_uint64),
                         ("orc_verts_per_cell", C.c_int), ("orc_points", C.c_void_p),
                         ("orc_cells", C.c_void_p), ("orc_cell_data", C.c_void_p),
                         ("orc_cell_data_bytes", C.c_uint64), ("orc_step_length_used", C.c_double),
                         ("orc_points_before_slice", C.c_void_p), ("orc_cells_before_slice", C.c_void_p)]:
            getattr(L, name).restype = rt
            getattr(L, name).argtypes = [C.c_void_p]
        L.orc_free.restype = None
        L.orc_free.argtypes = [C.c_void_p]
        L.orc_classify.restype = C.c_int
        L.orc_classify.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_uint64), C.c_double, C.c_void_p, C.c_uint64]
        L.orc_project_points.restype = C.c_int
        L.orc_project_points.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_double),
                                         C.POINTER(C.c_double), C.POINTER(_Params), C.c_void_p, C.c_uint64]
        L.orc_sample.restype = C.c_int
        L.orc_sample.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_double),
                                 C.POINTER(C.c_double), C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


@dataclass
class Mesh:
    points: np.ndarray      # (n, 3) float32
    cells: np.ndarray       # (m, 3|4) uint64
    cell_data: np.ndarray | None
    step_length_used: float
    points_before_slice: np.ndarray | None = None   # (nz + 1,) points that exist when the loop enters slice z
    cells_before_slice: np.ndarray | None = None


def _geom(vol: np.ndarray, spacing, origin):
    assert vol.ndim == 3 and vol.flags.c_contiguous, "volume must be a C-contiguous (z, y, x) array"
    dims = (C.c_uint64 * 3)(vol.shape[2], vol.shape[1], vol.shape[0])
    sp = (C.c_double * 3)(*(spacing if spacing is not None else (1.0, 1.0, 1.0)))
    og = (C.c_double * 3)(*(origin if origin is not None else (0.0, 0.0, 0.0)))
    return dims, sp, og


def _params(iso, triangles, project, cell_data, mode, thr, step, relax, max_steps, border_faces=False,
            region_index=(0, 0, 0), direction=None, method=0) -> _Params:
    d = [float(v) for v in np.asarray(direction, np.float64).reshape(9)] if direction is not None else [0.0] * 9
    return _Params(float(iso), int(bool(triangles)), int(bool(project)), int(bool(cell_data)), int(mode),
                   float(thr), float(step), float(relax), int(max_steps), int(bool(border_faces)),
                   (C.c_int64 * 3)(*[int(v) for v in region_index]), (C.c_double * 9)(*d), int(method), 0)


def cuberille(vol: np.ndarray, iso, *, triangles=True, project=True, cell_data=False, mode=LITERAL,
              thr=0.5, step=-1.0, relax=0.95, max_steps=50, spacing=None, origin=None, border_faces=False,
              region_index=(0, 0, 0), direction=None, method=0) -> Mesh:
    """Run the oracle.  `vol` is indexed [z, y, x] (x fastest), like a MetaImage buffer."""
    L = lib()
    dims, sp, og = _geom(vol, spacing, origin)
    P = _params(iso, triangles, project, cell_data, mode, thr, step, relax, max_steps, border_faces, region_index, direction, method)
    h = L.orc_cuberille(vol.ctypes.data, DTYPE_CODES[vol.dtype], dims, sp, og, C.byref(P))
    if not h:
        raise RuntimeError("oracle: unsupported dtype")
    try:
        n, m, k = L.orc_num_points(h), L.orc_num_cells(h), L.orc_verts_per_cell(h)
        pts = np.empty((n, 3), np.float32)
        if n:
            C.memmove(pts.ctypes.data, L.orc_points(h), pts.nbytes)
        cells = np.empty((m, k), np.uint64)
        if m:
            C.memmove(cells.ctypes.data, L.orc_cells(h), cells.nbytes)
        cd = None
        if cell_data:
            cd = np.empty(m, vol.dtype)
            if m:
                assert L.orc_cell_data_bytes(h) == cd.nbytes
                C.memmove(cd.ctypes.data, L.orc_cell_data(h), cd.nbytes)
        nz = vol.shape[0]
        pb = np.empty(nz + 1, np.uint64)
        cb = np.empty(nz + 1, np.uint64)
        C.memmove(pb.ctypes.data, L.orc_points_before_slice(h), pb.nbytes)
        C.memmove(cb.ctypes.data, L.orc_cells_before_slice(h), cb.nbytes)
        return Mesh(pts, cells, cd, L.orc_step_length_used(h), pb, cb)
    finally:
        L.orc_free(h)


def classify(vol: np.ndarray, iso, words_per_row: int | None = None) -> np.ndarray:
    """1 bit / voxel inside mask, shape (z, y, words_per_row) uint32; padding bits zero."""
    L = lib()
    dims, _, _ = _geom(vol, None, None)
    wpr = words_per_row or (vol.shape[2] + 31) // 32
    out = np.zeros((vol.shape[0], vol.shape[1], wpr), np.uint32)
    rc = L.orc_classify(vol.ctypes.data, DTYPE_CODES[vol.dtype], dims, float(iso), out.ctypes.data, wpr)
    assert rc == 0
    return out


def project_points(vol: np.ndarray, iso, pts: np.ndarray, *, thr=0.5, step=-1.0, relax=0.95, max_steps=50,
                   spacing=None, origin=None, direction=None, method=0) -> np.ndarray:
    L = lib()
    dims, sp, og = _geom(vol, spacing, origin)
    P = _params(iso, True, True, False, LITERAL, thr, step, relax, max_steps, direction=direction, method=method)
    out = np.ascontiguousarray(pts, np.float32).copy()
    L.orc_project_points(vol.ctypes.data, DTYPE_CODES[vol.dtype], dims, sp, og, C.byref(P), out.ctypes.data,
                         out.shape[0])
    return out


def sample(vol: np.ndarray, pts: np.ndarray, spacing=None, origin=None):
    L = lib()
    dims, sp, og = _geom(vol, spacing, origin)
    p = np.ascontiguousarray(pts, np.float64)
    val = np.empty(p.shape[0], np.float64)
    grad = np.empty((p.shape[0], 3), np.float64)
    L.orc_sample(vol.ctypes.data, DTYPE_CODES[vol.dtype], dims, sp, og, p.ctypes.data, p.shape[0],
                 val.ctypes.data, grad.ctypes.data)
    return val, grad
