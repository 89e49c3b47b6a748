// k_vertices.cuh — K3b: one thread per vertex: its position, and the corner -> vertex-id map.
//
// Reference: AddVertex without the projection (txx:257-276): TransformIndexToPhysicalPoint, minus half a
// spacing, mesh->GetPoints()->InsertElement(id, vertex).
//
// K3a (k_sweep.cuh, MODE_ASSIGN) recorded, for every vertex id in the reference's creation order, which
// lattice corner it sits on.  This kernel is perfectly balanced and coalesced: thread id reads 8 bytes,
// writes its 12-byte point, and stores its id at the corner's rank in corner-raster order
//     slot(corner) = cofs[corner word] + popc(act[corner word] & bits below the corner)
// (perm[slot] = id), which is where the face kernel looks it up.  In raster vertex order (the opt-in
// canonical order) ids ARE slots and k_points_raster writes the points straight from the active masks.
#pragma once
#include "cub_common.cuh"

namespace cub {

struct VertexArgs {
  const uint2* vtx;         // [n] packed corner of vertex id (scan-relative id)
  size_t n;                 // ghost vertices + own vertices
  size_t first_point;       // ids below this one belong to the slab underneath: no point is written
  const uint32_t* act;      // entry lattice [Zl+1][EY][EW]
  const uint32_t* cofs;
  int EY, EW;
  int plane_lo, plane_hi;   // local corner planes that faces of this handle reference (inclusive)
  int zg0;                  // global z of local plane 0
  Geom geom;
  float* points;            // indexed by scan-relative vertex id
  uint32_t* perm;           // [active corners of planes plane_lo..plane_hi] -> scan-relative vertex id
};

__global__ void __launch_bounds__(256) k_vertices(const VertexArgs a) {
  const size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= a.n) return;
  const uint2 v = a.vtx[id];
  const int cx = (int)(v.x & 0xffffu), cy = (int)(v.x >> 16), cz = (int)v.y;
  if (id >= a.first_point) {
    float* p = a.points + 3 * id;
    p[0] = corner_coord(a.geom.spacing[0], a.geom.origin[0], cx);
    p[1] = corner_coord(a.geom.spacing[1], a.geom.origin[1], cy);
    p[2] = corner_coord(a.geom.spacing[2], a.geom.origin[2], cz + a.zg0);
  }
  if (cz >= a.plane_lo && cz <= a.plane_hi) {
    const size_t e = ((size_t)cz * a.EY + cy) * a.EW + (cx >> 5);
    const uint32_t below = (1u << (cx & 31)) - 1u;
    a.perm[__ldg(a.cofs + e) + __popc(__ldg(a.act + e) & below)] = (uint32_t)id;
  }
}

// Raster vertex order (CUB_ORDER_RASTER): vertex id = slot, so the points follow directly from the active
// masks: one thread per corner word, ids cofs[word] + rank are consecutive inside a word.
struct RasterPointArgs {
  const uint32_t* act;
  const uint32_t* cofs;
  int EY, EW, Wc;
  int plane_lo, plane_hi;   // local corner planes to emit (inclusive)
  int zg0;
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
  const float py = corner_coord(a.geom.spacing[1], a.geom.origin[1], cy);
  const float pz = corner_coord(a.geom.spacing[2], a.geom.origin[2], cz + a.zg0);
  while (m) {
    const int b = __ffs(m) - 1;
    m &= m - 1;
    float* p = a.points + 3 * (size_t)id++;
    p[0] = corner_coord(a.geom.spacing[0], a.geom.origin[0], 32 * w + b);
    p[1] = py;
    p[2] = pz;
  }
}

}  // namespace cub
