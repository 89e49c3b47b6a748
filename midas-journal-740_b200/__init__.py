"""cuberille-b200: B200-native (sm_100a) hot path of itk::CuberilleImageToMeshFilter::GenerateData.

Layout:
  csrc/                 hand-written CUDA kernels + the C-ABI (include/cuberille_c.h)
  capi.py               ctypes binding of the C-ABI
  filter.py             Python mirror of the reference filter's interface (tests / bench)
  mha.py                MetaImage reader/writer, VTK polydata writer (the test driver's IO)
  slabs.py              z-slab plan + the count all-gather of multi-GPU runs
The C++ drop-in adapter lives in include/itkCuberilleImageToMeshFilter.h.
"""
from . import capi, mha, slabs  # noqa: F401
from ._build import build  # noqa: F401
from .filter import CuberilleImageToMeshFilter, LinearInterpolateImageFunction, Mesh  # noqa: F401
from .mha import Image, read_mha, write_mha, write_vtk_polydata  # noqa: F401
