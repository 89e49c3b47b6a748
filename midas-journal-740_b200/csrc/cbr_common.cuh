// cbr_common.cuh — shared device-side definitions of libcuberille_cuda.so (sm_100a).
//
// Data layout in HBM:
//   volume   : [Zl][Y][X] pixels, x fastest (the itk::Image buffer, or a z-slab of it)
//   bitmask  : [Zl][Y][Wp] uint32, bit b of word w of a row = "voxel x=32w+b is inside"
//              ( inside == !(v < iso), txx:139-141 ).  Wp = roundup(ceil(X/32), 4) so every
//              row is 16-byte aligned; bits x >= X of the last valid word replicate bit X-1
//              (that makes the +x edge-replicate of txx:167 a plain shift); pad words are 0.
//   lattice  : [Zl+1][Y+1][EW] uint32 entries, one per corner word (EW = roundup(ceil((X+1)/32)+1, 4)); entry
//              (z, y, w) also stands for voxel word w of row y of slice z.  K2a writes cnt (owned corners |
//              faces << 10 | active corners << 20), act (active-corner mask) and own (the 8 ownership masks);
//              K2b turns cnt into vofs / fofs / cofs, the exclusive prefixes (in raster order over the scan
//              range) of owned corners, faces and active corners.
//
// The reference keeps a std::map per corner plane (h:243-313) to find out whether a corner
// already has a vertex.  Here ownership is a closed form of the 2x2x2 inside bits around a corner:
// a corner belongs to the first voxel, in raster order, among the <=8 voxels around it that
// activates it (vertexHasQuad, txx:164-173), and that voxel numbers its owned corners in
// local order 0..7 (txx:179-194).  k_sweep.cuh evaluates that rule for 32 corners at a time
// with bitwise operations on the inside words.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cbr {

struct Grid {
  int X, Y, Zl;  // local buffer size in voxels
  int Wx, Wp;    // valid / padded 32-bit words per row
  int zg0;       // global z index of local slice 0
  int Zg;        // z size of the whole image
};

__host__ __device__ inline size_t row_index(const Grid& g, int y, int zl) { return (size_t)zl * g.Y + y; }

// corner offsets of local vertex l (txx:236-254): bit0 -> x, bit1 -> y, bit2 -> z of kCorner[l]
//   0(0,0,0) 1(1,0,0) 2(1,1,0) 3(0,1,0) 4(0,0,1) 5(1,0,1) 6(1,1,1) 7(0,1,1)
__host__ __device__ constexpr int corner_ox(int l) { return (l == 1 || l == 2 || l == 5 || l == 6) ? 1 : 0; }
__host__ __device__ constexpr int corner_oy(int l) { return (l == 2 || l == 3 || l == 6 || l == 7) ? 1 : 0; }
__host__ __device__ constexpr int corner_oz(int l) { return (l >= 4) ? 1 : 0; }

// corners of face f in emission order (txx:197-202, 219-233)
__device__ __constant__ const int8_t kFaceCorners[6][4] = {{0, 4, 7, 3}, {0, 1, 5, 4}, {1, 2, 6, 5},
                                                           {2, 3, 7, 6}, {0, 3, 2, 1}, {4, 5, 6, 7}};

// Programmatic dependent launch (the kernels of a step are queued with cudaLaunchAttributeProgrammaticStreamSerialization):
// the first statement of every kernel of the hot path.  `wait` returns once the kernel launched before this one in the
// stream has completed and its writes are visible (a no-op for a plain launch); `launch_dependents` then lets the next
// kernel's CTAs be scheduled as this grid drains, up to their own `wait` - launch latency and prologue leave the
// critical path (a step of a thin z-slab is eight kernels of 5..150 us each).
__device__ __forceinline__ void pdl_enter() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

struct Geom {
  double spacing[3];
  double origin[3];
  // oriented images (a non-identity direction matrix): the direction cosines D, the index -> physical matrix
  // M = D * diag(spacing) and its inverse, all row-major; unused (the ITK 3.x expressions are evaluated) when
  // oriented == 0
  double dir[9];
  double m[9];
  double minv[9];
  int oriented;
};

// unprojected vertex position, AddVertex txx:265-270:
//   TransformIndexToPhysicalPoint:  identity direction (every test of the reference), the ITK 3.x form
//                                       p = (float)(spacing*index + origin)
//                                   oriented image, the ITK 4/5 form (the point is a Point<float>, so every
//                                   accumulation rounds to float)
//                                       p = (float)origin;  for j: p = (float)((double)p + M[a][j]*index[j])
//   then, on every PHYSICAL axis (the reference does not rotate the shift, txx:268-270):
//                                       p = (float)((double)p - spacing/2)
// (ORIENTED is a template parameter of the kernels: the non-oriented instantiation must not carry the other path)
template <bool ORIENTED>
__device__ __forceinline__ float corner_coord(const Geom& g, int a, int ix, int iy, int iz) {
  float p;
  if (!ORIENTED) {
    const int idx = a == 0 ? ix : (a == 1 ? iy : iz);
    p = (float)__dadd_rn(__dmul_rn(g.spacing[a], (double)idx), g.origin[a]);
  } else {
    p = (float)g.origin[a];
    p = (float)__dadd_rn((double)p, __dmul_rn(g.m[3 * a + 0], (double)ix));
    p = (float)__dadd_rn((double)p, __dmul_rn(g.m[3 * a + 1], (double)iy));
    p = (float)__dadd_rn((double)p, __dmul_rn(g.m[3 * a + 2], (double)iz));
  }
  return (float)__dadd_rn((double)p, -(g.spacing[a] / 2.0));
}

}  // namespace cbr
