// cub_common.cuh — shared device-side definitions of libcuberille_cuda.so (sm_100a).
//
// Data layout in HBM:
//   volume   : [Zl][Y][X] pixels, x fastest (the itk::Image buffer, or a z-slab of it)
//   bitmask  : [Zl][Y][Wp] uint32, bit b of word w of a row = "voxel x=32w+b is inside"
//              ( inside == !(v < iso), txx:139-141 ).  Wp = roundup(ceil(X/32), 4) so every
//              row is 16-byte aligned; bits x >= X of the last valid word replicate bit X-1
//              (that makes the +x edge-replicate of txx:167 a plain shift); pad words are 0.
//   vofs/fofs: [Zl][Y][Wp] uint32, exclusive prefix (in voxel-raster order over the scan
//              range) of the number of corner vertices OWNED / quads EMITTED by the voxels
//              before this word.
//
// The reference keeps a std::map per corner plane (h:243-313) to find out whether a corner
// already has a vertex.  Here ownership is a closed form of the 3x3x3 inside-neighbourhood:
// a corner belongs to the first voxel, in raster order, among the <=8 voxels around it that
// activates it (vertexHasQuad, txx:164-173), and that voxel numbers its owned corners in
// local order 0..7 (txx:179-194).  Everything below evaluates that rule for 32 voxels at a
// time with bitwise operations on the inside words.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cub {

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

// 27 inside-words around one word of 32 voxels: n[dz+1][dy+1][dx+1] holds, at bit b, the inside
// flag of voxel (x+dx, y+dy, z+dz) with every coordinate clamped to the image
// (ZeroFluxNeumannBoundaryCondition).
struct Nbhd {
  uint32_t n[3][3][3];
  uint32_t vx[3];  // bit mask: is (x+dx) inside the image      (dx = -1, 0, +1)
  uint32_t vy[3];  // 0 / ~0 : is (y+dy) inside the image
  uint32_t vz[3];  // 0 / ~0 : is (z+dz) inside the image
};

// shift one row's words to the x-1 / x+1 aligned views with edge replication
__device__ __forceinline__ void shift_lr(uint32_t c, uint32_t prev, uint32_t next, bool first, bool last,
                                         uint32_t& l, uint32_t& r) {
  l = (c << 1) | (first ? (c & 1u) : (prev >> 31));
  r = (c >> 1) | (last ? (c & 0x80000000u) : (next << 31));
}

__device__ __forceinline__ void set_validity(Nbhd& nb, const Grid& g, int w, int y, int zl) {
  const int zg = zl + g.zg0;
  nb.vx[0] = (w == 0) ? ~1u : ~0u;
  nb.vx[1] = (w == g.Wx - 1 && (g.X & 31)) ? ((1u << (g.X & 31)) - 1u) : ~0u;
  nb.vx[2] = (w == g.Wx - 1) ? (nb.vx[1] & ~(1u << ((g.X - 1) & 31))) : ~0u;
  nb.vy[0] = (y > 0) ? ~0u : 0u;
  nb.vy[1] = ~0u;
  nb.vy[2] = (y < g.Y - 1) ? ~0u : 0u;
  nb.vz[0] = (zg > 0) ? ~0u : 0u;
  nb.vz[1] = ~0u;
  nb.vz[2] = (zg < g.Zg - 1) ? ~0u : 0u;
}

// pointers to the 9 bitmask rows around (y, zl), clamped to the image (and, for memory safety,
// to the local buffer: the slab contract of cub_set_slab guarantees the clamp never bites)
__device__ __forceinline__ void row_pointers(const uint32_t* __restrict__ bits, const Grid& g, int y, int zl,
                                             const uint32_t* rp[3][3]) {
#pragma unroll
  for (int dz = -1; dz <= 1; ++dz) {
    int zg = zl + g.zg0 + dz;
    zg = zg < 0 ? 0 : (zg > g.Zg - 1 ? g.Zg - 1 : zg);
    int z = zg - g.zg0;
    z = z < 0 ? 0 : (z > g.Zl - 1 ? g.Zl - 1 : z);
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
      int yy = y + dy;
      yy = yy < 0 ? 0 : (yy > g.Y - 1 ? g.Y - 1 : yy);
      rp[dz + 1][dy + 1] = bits + row_index(g, yy, z) * (size_t)g.Wp;
    }
  }
}

// Loads the neighbourhood of word w of row (y, zl) with 27 scalar loads (L1/L2 resident).
__device__ __forceinline__ void load_nbhd(const uint32_t* __restrict__ bits, const Grid& g, int w, int y, int zl,
                                          Nbhd& nb) {
  const uint32_t* rp[3][3];
  row_pointers(bits, g, y, zl, rp);
  const bool first = (w == 0), last = (w == g.Wx - 1);
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      const uint32_t c = __ldg(rp[a][b] + w);
      const uint32_t p = first ? 0u : __ldg(rp[a][b] + w - 1);
      const uint32_t nx = last ? 0u : __ldg(rp[a][b] + w + 1);
      nb.n[a][b][1] = c;
      shift_lr(c, p, nx, first, last, nb.n[a][b][0], nb.n[a][b][2]);
    }
  set_validity(nb, g, w, y, zl);
}

// Face masks F[f] (bit b: voxel 32w+b emits face f; txx:164-173 with offsets txx:121-127) and
// ownership masks O[l] (bit b: voxel 32w+b creates the vertex of its local corner l, txx:179-194).
__device__ __forceinline__ void compute_masks(const Nbhd& nb, uint32_t F[6], uint32_t O[8]) {
  const uint32_t c = nb.n[1][1][1] & nb.vx[1];
  F[0] = c & ~nb.n[1][1][0];  // -x
  F[1] = c & ~nb.n[1][0][1];  // -y
  F[2] = c & ~nb.n[1][1][2];  // +x
  F[3] = c & ~nb.n[1][2][1];  // +y
  F[4] = c & ~nb.n[0][1][1];  // -z
  F[5] = c & ~nb.n[2][1][1];  // +z
#pragma unroll
  for (int l = 0; l < 8; ++l) {
    const int ox = corner_ox(l), oy = corner_oy(l), oz = corner_oz(l);
    // this voxel's position inside the 2x2x2 block around the corner, block raster order
    const int pv = (1 - oz) * 4 + (1 - oy) * 2 + (1 - ox);
    uint32_t earlier = 0;
#pragma unroll
    for (int p = 0; p < 8; ++p) {
      if (p < pv) {
        const int qx = p & 1, qy = (p >> 1) & 1, qz = p >> 2;
        const int dx = ox + qx - 1, dy = oy + qy - 1, dz = oz + qz - 1;  // block voxel relative to this voxel
        const int ex = dx + (qx == 0 ? 1 : -1);                           // its in-block neighbours
        const int ey = dy + (qy == 0 ? 1 : -1);
        const int ez = dz + (qz == 0 ? 1 : -1);
        const uint32_t u = nb.n[dz + 1][dy + 1][dx + 1];
        const uint32_t all_in = nb.n[dz + 1][dy + 1][ex + 1] & nb.n[dz + 1][ey + 1][dx + 1] & nb.n[ez + 1][dy + 1][dx + 1];
        earlier |= u & ~all_in & nb.vx[dx + 1] & nb.vy[dy + 1] & nb.vz[dz + 1];
      }
    }
    const uint32_t h = F[ox ? 2 : 0] | F[oy ? 3 : 1] | F[oz ? 5 : 4];  // vertexHasQuad[l]
    O[l] = h & ~earlier;
  }
}

// unprojected vertex position, AddVertex txx:265-270 (ITK 3.x form):
//   p = (float)(spacing*index + origin);  p = (float)((double)p - spacing/2)
__device__ __forceinline__ float corner_coord(double spacing, double origin, int idx) {
  float p = (float)__dadd_rn(__dmul_rn(spacing, (double)idx), origin);
  return (float)__dadd_rn((double)p, -(spacing / 2.0));
}

struct Geom {
  double spacing[3];
  double origin[3];
};

}  // namespace cub
