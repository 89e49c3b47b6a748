// k_assign.cuh - K3a: vertex id -> lattice corner, in the reference's creation order.
//
// Reference: the vertex creation loop txx:179-194: voxels in raster order, local corners 0..7, a corner gets
// nextVertexId the first time it is touched.  K2a (k_sweep.cuh) already decided, per voxel word, which voxel
// owns which of its 8 local corners (8 masks, stored where the word owns anything) and K2b turned the counts
// into vofs = first id of the word.  This kernel only walks the masks: one thread per voxel word, voxels in
// bit order, local corners 0..7, writing for vertex id vofs + rank the corner it sits on
// (cx | cy << 16 | oz << 31; k_vertices.cuh recovers the slice from the id).
// The first r1 versions swept the volume a second time to recompute the masks (k_sweep<ASSIGN>, kept behind
// CUB_ASSIGN_SWEEP=1): 0.52 ms against 1.1 GB of extra scratch traffic here.
#pragma once
#include "cub_common.cuh"

namespace cub {

struct AssignArgs {
  const uint32_t* cnt;    // entry lattice: owned corners in the low 10 bits
  const uint32_t* vofs;   // entry lattice: first vertex id of the word
  const uint4* own;       // entry lattice x 2: ownership masks O[0..3], O[4..7]
  int X, Y, Wx, EY, EW;
  int z_begin;            // first local slice of the scan range (blockIdx.z = 0)
  uint32_t* vtx;          // [n vertices] cx | cy << 16 | oz << 31
};

constexpr int kAssignThreads = 128;  // (one row per warp)

__global__ void __launch_bounds__(kAssignThreads) k_assign(const AssignArgs a) {
  // grid: x = 32-word segments of a row, y = groups of kAssignThreads / 32 rows (one row per warp), z = slices of the scan range
  const int w = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * (kAssignThreads / 32) + (threadIdx.x >> 5), z = a.z_begin + blockIdx.z;
  if (w >= a.Wx || y >= a.Y) return;
  const uint32_t e = ((uint32_t)z * (uint32_t)a.EY + (uint32_t)y) * (uint32_t)a.EW + (uint32_t)w;
  if ((__ldg(a.cnt + e) & 0x3ffu) == 0) return;
  uint32_t n = __ldg(a.vofs + e);
  const uint4 lo = __ldcs(a.own + 2 * (size_t)e), hi = __ldcs(a.own + 2 * (size_t)e + 1);
  const uint32_t O[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
  uint32_t U = O[0] | O[1] | O[2] | O[3] | O[4] | O[5] | O[6] | O[7];
  uint32_t* __restrict__ const out = a.vtx;
  const uint32_t xy0 = (uint32_t)(w * 32) | ((uint32_t)y << 16);
  while (U) {
    const int b = __ffs(U) - 1;
    U &= U - 1;
    const uint32_t bit = 1u << b;
    const uint32_t xy = xy0 + (uint32_t)b;
    // local corner l -> (ox, oy, oz) as in txx:236-254; a 32-bit running index (a predicated 64-bit pointer bump
    // costs 6 instructions per store)
    if (O[0] & bit) { out[n] = xy; ++n; }
    if (O[1] & bit) { out[n] = xy + 1u; ++n; }
    if (O[2] & bit) { out[n] = xy + 0x10001u; ++n; }
    if (O[3] & bit) { out[n] = xy + 0x10000u; ++n; }
    if (O[4] & bit) { out[n] = xy + 0x80000000u; ++n; }
    if (O[5] & bit) { out[n] = xy + 0x80000001u; ++n; }
    if (O[6] & bit) { out[n] = xy + 0x80010001u; ++n; }
    if (O[7] & bit) { out[n] = xy + 0x80010000u; ++n; }
  }
}

}  // namespace cub
