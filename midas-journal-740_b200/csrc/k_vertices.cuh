// k_vertices.cuh — K3b: one thread per vertex: its position, and the corner -> vertex-id map.
//
// Reference: AddVertex without the projection (txx:257-276): TransformIndexToPhysicalPoint, minus half a
// spacing, mesh->GetPoints()->InsertElement(id, vertex).
//
// K3a (k_sweep.cuh, MODE_ASSIGN) recorded, for every vertex id in the reference's creation order, which
// lattice corner it sits on (4 bytes: cx, cy and whether the corner is on the upper plane of its owner's
// slice; the owner slice follows from the id because ids are handed out slice by slice).  This kernel is
// perfectly balanced and coalesced: thread id reads 4 bytes,
// writes its 12-byte point, and stores its id at the corner's rank in corner-raster order
//     slot(corner) = cofs[corner word] + popc(act[corner word] & bits below the corner)
// (perm[slot] = id), which is where the face kernel looks it up.  In raster vertex order (the opt-in
// canonical order) ids ARE slots and k_points_raster writes the points straight from the active masks.
#pragma once
#include "cub_common.cuh"

namespace cub {

struct VertexArgs {
  const uint32_t* vtx;      // [n] packed corner of vertex id (scan-relative id): cx | cy << 16 | oz << 31
  const uint32_t* slice_first;  // [nz + 1] first id created by slice z_first + k; [nz] = UINT_MAX   (k_slice_index)
  const uint32_t* block_slice;  // [blocks of this launch] k of the block's first id                  (k_slice_index)
  int z_first;              // first local slice of the scan range
  size_t n;                 // ghost vertices + own vertices
  size_t first_point;       // ids below this one belong to the slab underneath: no point is written
  const uint32_t* act;      // entry lattice [Zl+1][EY][EW]
  const uint32_t* cofs;
  int EY, EW;
  int plane_lo, plane_hi;   // local corner planes that faces of this handle reference (inclusive)
  int coff[3];              // image index of lattice corner (0, 0, 0): slab offset, region index, minus the pad of image_border_faces
  Geom geom;
  float* points;            // indexed by scan-relative vertex id
  uint32_t* perm;           // [active corners of planes plane_lo..plane_hi] -> scan-relative vertex id
};

// ids are handed out slice by slice, so the owner slice of an id follows from the per-slice first ids: this
// one-off kernel compacts them out of the entry lattice and bisects once per k_vertices block, so that the
// 44 M vertex threads only step forward from their block's slice (almost always zero steps).
struct SliceIndexArgs {
  const uint32_t* vofs;     // entry lattice: vofs[z * plane_entries] = first id created by slice z
  size_t plane_entries;
  int z_first, nz;          // local slices [z_first, z_first + nz) of the scan range
  uint32_t* slice_first;    // [nz + 1]
  uint32_t* block_slice;    // [n_blocks]
  uint32_t n_blocks, ids_per_block;
};

__global__ void __launch_bounds__(256) k_slice_index(const SliceIndexArgs a) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t <= (uint32_t)a.nz) a.slice_first[t] = t < (uint32_t)a.nz ? __ldg(a.vofs + (size_t)(a.z_first + t) * a.plane_entries) : 0xffffffffu;
  if (t >= a.n_blocks) return;
  const uint32_t id = t * a.ids_per_block;
  int lo = 0, hi = a.nz - 1;  // largest k with first[k] <= id
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (__ldg(a.vofs + (size_t)(a.z_first + mid) * a.plane_entries) <= id) lo = mid; else hi = mid - 1;
  }
  a.block_slice[t] = (uint32_t)lo;
}

constexpr int kVertexPerThread = 4;                       // ids per thread: the loads of the 4 are in flight together
constexpr int kVertexBlockIds = 256 * kVertexPerThread;   // (the kernel is a chain of 3 dependent loads otherwise)

__global__ void __launch_bounds__(256) k_vertices(const VertexArgs a) {
  const size_t id0 = (size_t)blockIdx.x * kVertexBlockIds + threadIdx.x;
  uint32_t v[kVertexPerThread];
#pragma unroll
  for (int j = 0; j < kVertexPerThread; ++j) {
    const size_t id = id0 + (size_t)j * 256;
    v[j] = id < a.n ? __ldcs(a.vtx + id) : 0u;
  }
  int lo = (int)__ldg(a.block_slice + blockIdx.x);
  int cz[kVertexPerThread];
#pragma unroll
  for (int j = 0; j < kVertexPerThread; ++j) {
    const size_t id = id0 + (size_t)j * 256;
    if (id < a.n)
      while ((uint32_t)id >= __ldg(a.slice_first + lo + 1)) ++lo;
    cz[j] = a.z_first + lo + (int)(v[j] >> 31);
  }
  uint32_t co[kVertexPerThread], ac[kVertexPerThread];
#pragma unroll
  for (int j = 0; j < kVertexPerThread; ++j) {
    const size_t id = id0 + (size_t)j * 256;
    const int cx = (int)(v[j] & 0xffffu), cy = (int)((v[j] >> 16) & 0x7fffu);
    co[j] = ac[j] = 0;
    if (id < a.n && cz[j] >= a.plane_lo && cz[j] <= a.plane_hi) {
      const size_t e = ((size_t)cz[j] * a.EY + cy) * a.EW + (cx >> 5);
      co[j] = __ldg(a.cofs + e);
      ac[j] = __ldg(a.act + e);
    }
  }
#pragma unroll
  for (int j = 0; j < kVertexPerThread; ++j) {
    const size_t id = id0 + (size_t)j * 256;
    if (id >= a.n) break;
    const int cx = (int)(v[j] & 0xffffu), cy = (int)((v[j] >> 16) & 0x7fffu);
    if (id >= a.first_point) {
      float* p = a.points + 3 * id;
      p[0] = corner_coord(a.geom.spacing[0], a.geom.origin[0], cx + a.coff[0]);
      p[1] = corner_coord(a.geom.spacing[1], a.geom.origin[1], cy + a.coff[1]);
      p[2] = corner_coord(a.geom.spacing[2], a.geom.origin[2], cz[j] + a.coff[2]);
    }
    if (cz[j] >= a.plane_lo && cz[j] <= a.plane_hi) {
      const uint32_t below = (1u << (cx & 31)) - 1u;
      a.perm[co[j] + __popc(ac[j] & below)] = (uint32_t)id;
    }
  }
}

// Raster vertex order (CUB_ORDER_RASTER): vertex id = slot, so the points follow directly from the active
// masks: one thread per corner word, ids cofs[word] + rank are consecutive inside a word.
struct RasterPointArgs {
  const uint32_t* act;
  const uint32_t* cofs;
  int EY, EW, Wc;
  int plane_lo, plane_hi;   // local corner planes to emit (inclusive)
  int coff[3];              // as in VertexArgs
  Geom geom;
  float* points;            // indexed by slot
};

__global__ void __launch_bounds__(256) k_points_raster(const RasterPointArgs a) {
  // grid: x = 32-word segments of a corner row, y = groups of 8 rows (one per warp), z = planes
  const int w = blockIdx.x * 32 + (threadIdx.x & 31), cy = blockIdx.y * 8 + (threadIdx.x >> 5), cz = a.plane_lo + blockIdx.z;
  if (w >= a.Wc || cy >= a.EY) return;
  const size_t e = ((size_t)cz * a.EY + cy) * a.EW + w;
  uint32_t m = __ldg(a.act + e);
  if (!m) return;
  uint32_t id = __ldg(a.cofs + e);
  const float py = corner_coord(a.geom.spacing[1], a.geom.origin[1], cy + a.coff[1]);
  const float pz = corner_coord(a.geom.spacing[2], a.geom.origin[2], cz + a.coff[2]);
  while (m) {
    const int b = __ffs(m) - 1;
    m &= m - 1;
    float* p = a.points + 3 * (size_t)id++;
    p[0] = corner_coord(a.geom.spacing[0], a.geom.origin[0], 32 * w + b + a.coff[0]);
    p[1] = py;
    p[2] = pz;
  }
}

}  // namespace cub
