// k_vertices.cuh — K3a: vertex creation.  One warp per 32-word segment of a voxel row walks the ownership masks
// K2a stored and emits, for every owned corner, its point and its entry of the corner -> vertex-id map.
//
// Reference: the vertex creation loop txx:179-194 (voxels in raster order, local corners 0..7, a corner gets
// nextVertexId the first time it is touched) and AddVertex without the projection (txx:257-276):
// TransformIndexToPhysicalPoint, minus half a spacing, mesh->GetPoints()->InsertElement(id, vertex).
//
// Round 1 did this in three kernels (k_assign: one thread per word walking its masks and writing 4-byte records
// in id order; k_slice_index; k_vertices: one thread per record -> point + map entry: 0.48 ms, 2.4 GB of DRAM
// traffic for 44 M vertices).  Here the ids of a segment are consecutive (segment base from k_seg_scan + a
// warp scan of the words' counts), so the warp
//   1. walks the masks of its 32 words (voxel bit, then local corner 0..7: the reference's creation order
//      inside the word) into a shared-memory queue of (word, bit, local corner) items - queue position = id, and
//   2. handles one item per lane: the point (consecutive lanes write consecutive ids: coalesced) and
//          perm[slot(corner)] = id,   slot(corner) = cseg[corner row segment] + popc(active corners before it),
//      which is where the face kernel looks the id up.  The active-corner prefix inside the segment is a warp
//      scan of popc(act) of the four corner rows around the voxel row.
// No per-vertex record array, no slice bisection, no gather of dense offset arrays.
//
// In raster vertex order (the opt-in canonical order) ids ARE slots and k_points_raster writes the points
// straight from the active masks.
#pragma once
#include "cbr_common.cuh"
#include "k_segscan.cuh"

namespace cbr {

struct VertexArgs {
  const uint32_t* cnt;      // entry lattice: owned corners in the low 10 bits
  const uint32_t* act;      // entry lattice: active-corner masks
  const uint4* own;         // entry lattice x 2: ownership masks O[0..3], O[4..7] (valid where the word owns a corner)
  const uint4* seg;         // [lattice rows][NS] segment bases {vertices, faces, active corners, -}
  const unsigned long long* info;
  int Wx, Y, EY, EW, NS;
  int z_begin;              // first local slice of the scan range (blockIdx.z = 0)
  int plane_lo, plane_hi;   // local corner planes that faces of this handle reference (inclusive)
  int write_ghost_points;   // also write the points of the vertices that belong to the slab underneath
  int coff[3];              // image index of lattice corner (0, 0, 0): slab offset, region index, minus the pad of image_border_faces
  Geom geom;
  float* points;            // indexed by scan-relative vertex id
  size_t points_cap;        // points the buffer can hold
  uint32_t* perm;           // [active corners of planes plane_lo..plane_hi] -> scan-relative vertex id
  size_t perm_cap;
  unsigned long long* flags;  // info + kInfoFlags
};

constexpr int kVertexThreads = 128;   // one voxel row segment per warp
constexpr int kVertexQueue = 512;     // items per warp and pass

struct VertexSmem {
  uint32_t A[kVertexThreads / 32][32][4];   // active masks of the 4 corner words around the voxel word (index oz*2+oy)
  uint32_t C[kVertexThreads / 32][32][4];   // their slots bases
  uint16_t queue[kVertexThreads / 32][kVertexQueue];  // word << 8 | bit << 3 | local corner
};

__global__ void __launch_bounds__(kVertexThreads) k_vertices(const VertexArgs a) {
  __shared__ VertexSmem sm;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sgm = blockIdx.x, w = sgm * 32 + lane, y = blockIdx.y * (kVertexThreads / 32) + warp, z = a.z_begin + blockIdx.z;
  const bool valid = w < a.Wx && y < a.Y;
  const int plane = a.EY * a.EW;
  const uint32_t e = ((uint32_t)z * (uint32_t)a.EY + (uint32_t)y) * (uint32_t)a.EW + (uint32_t)w;
  // everything is requested up front
  uint32_t nv = 0, A[4] = {0, 0, 0, 0};
  uint4 sb[4] = {};
  if (y < a.Y) {
    const size_t r0 = (size_t)z * a.EY + y;
#pragma unroll
    for (int k = 0; k < 4; ++k) sb[k] = __ldg(a.seg + (r0 + (size_t)(k >> 1) * a.EY + (k & 1)) * a.NS + sgm);
  }
  if (valid) {
    nv = __ldg(a.cnt + e) & 0x3ffu;
#pragma unroll
    for (int k = 0; k < 4; ++k) A[k] = __ldg(a.act + e + (uint32_t)((k >> 1) * plane + (k & 1) * a.EW));
  }
  // warp scans: owned corners (ids) and active corners of the four corner rows (slots), 16-bit fields
  uint32_t s0 = nv, s1 = (uint32_t)__popc(A[0]) | ((uint32_t)__popc(A[1]) << 16), s2 = (uint32_t)__popc(A[2]) | ((uint32_t)__popc(A[3]) << 16);
  const uint32_t m0 = s0, m1 = s1, m2 = s2;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t0 = __shfl_up_sync(0xffffffffu, s0, o), t1 = __shfl_up_sync(0xffffffffu, s1, o), t2 = __shfl_up_sync(0xffffffffu, s2, o);
    if (lane >= o) { s0 += t0; s1 += t1; s2 += t2; }
  }
  const uint32_t total = __shfl_sync(0xffffffffu, s0, 31);
  if (total == 0) return;  // no corner is owned by the 1024 voxels of the segment
  const uint32_t first = s0 - m0;  // position of the lane's first item
  s1 -= m1; s2 -= m2;
  if (nv) {
    uint32_t* cA = sm.A[warp][lane];
    uint32_t* cC = sm.C[warp][lane];
    cA[0] = A[0]; cA[1] = A[1]; cA[2] = A[2]; cA[3] = A[3];
    cC[0] = sb[0].z + (s1 & 0xffffu); cC[1] = sb[1].z + (s1 >> 16); cC[2] = sb[2].z + (s2 & 0xffffu); cC[3] = sb[3].z + (s2 >> 16);
  }
  uint32_t O[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (nv) {
    const uint4 lo = __ldcs(a.own + 2 * (size_t)e), hi = __ldcs(a.own + 2 * (size_t)e + 1);
    O[0] = lo.x; O[1] = lo.y; O[2] = lo.z; O[3] = lo.w; O[4] = hi.x; O[5] = hi.y; O[6] = hi.z; O[7] = hi.w;
  }
  const uint32_t Uall = O[0] | O[1] | O[2] | O[3] | O[4] | O[5] | O[6] | O[7];
  const uint32_t vbase = sb[0].x;                                   // first id of the segment (scan-relative)
  const unsigned long long first_point = a.write_ghost_points ? 0ull : __ldg(a.info + kInfoGhostV);
  uint16_t* q = sm.queue[warp];

  for (uint32_t base = 0; base < total; base += kVertexQueue) {
    // ---- 1. the lanes whose items fall into [base, base + kVertexQueue) walk their masks ------------------
    if (nv && first < base + kVertexQueue && first + nv > base) {
      uint32_t U = Uall, pos = first;
      const uint32_t tag = (uint32_t)lane << 8;
      while (U) {
        const int b = __ffs(U) - 1;
        U &= U - 1;
        // local corners of voxel b that it owns, as a byte (bit l <-> local corner l, txx:179-194 creation order)
        uint32_t m = 0;
#pragma unroll
        for (int l = 0; l < 8; ++l) m |= ((O[l] >> b) & 1u) << l;
        while (m) {
          const int l = __ffs(m) - 1;
          m &= m - 1;
          const uint32_t rel = pos - base;   // (wraps for items before the window)
          if (rel < (uint32_t)kVertexQueue) q[rel] = (uint16_t)(tag | ((uint32_t)b << 3) | (uint32_t)l);
          ++pos;
        }
      }
    }
    __syncwarp();
    // ---- 2. one item per lane -----------------------------------------------------------------------------
    const uint32_t end = min(total - base, (uint32_t)kVertexQueue);
    for (uint32_t s = lane; s < end; s += 32) {
      const uint32_t it = q[s];
      const uint32_t src = it >> 8, b = (it >> 3) & 31u, l = it & 7u;
      // local corner l -> (ox, oy, oz) as in txx:236-254: 0(0,0,0) 1(1,0,0) 2(1,1,0) 3(0,1,0) 4(0,0,1) 5(1,0,1) 6(1,1,1) 7(0,1,1)
      const uint32_t ox = (l ^ (l >> 1)) & 1u, oy = (l >> 1) & 1u, oz = l >> 2;
      const uint32_t k = oz * 2 + oy;
      const uint32_t cb = b + ox;  // corner bit inside corner word `src` (32: bit 0 of the next word)
      const uint32_t Ak = sm.A[warp][src][k], Ck = sm.C[warp][src][k];
      const uint32_t below = cb >= 32u ? 0xffffffffu : ((1u << cb) - 1u);
      const uint32_t slot = Ck + (uint32_t)__popc(Ak & below);
      const unsigned long long id = (unsigned long long)vbase + base + s;  // scan-relative vertex id
      const int cx = (sgm * 32 + (int)src) * 32 + (int)cb, cy = y + (int)oy, cz = z + (int)oz;
      if (id >= first_point) {
        if (id < a.points_cap) {
          float* p = a.points + 3 * id;
          p[0] = corner_coord(a.geom, 0, cx + a.coff[0], cy + a.coff[1], cz + a.coff[2]);
          p[1] = corner_coord(a.geom, 1, cx + a.coff[0], cy + a.coff[1], cz + a.coff[2]);
          p[2] = corner_coord(a.geom, 2, cx + a.coff[0], cy + a.coff[1], cz + a.coff[2]);
        } else {
          atomicOr(a.flags, (unsigned long long)kFlagBufferOverflow);
        }
      }
      if (cz >= a.plane_lo && cz <= a.plane_hi) {
        if (slot < a.perm_cap) a.perm[slot] = (uint32_t)id;
        else atomicOr(a.flags, (unsigned long long)kFlagBufferOverflow);
      }
    }
    __syncwarp();
  }
}

// Raster vertex order (CUB_ORDER_RASTER): vertex id = slot, so the points follow directly from the active
// masks: one thread per corner word, ids cseg[segment] + warp prefix + rank are consecutive inside a word.
struct RasterPointArgs {
  const uint32_t* act;
  const uint4* seg;
  int EY, EW, NS, Wc;
  int plane_lo, plane_hi;   // local corner planes to emit (inclusive)
  int coff[3];              // as in VertexArgs
  Geom geom;
  float* points;            // indexed by slot
  size_t points_cap;
  unsigned long long* flags;
};

__global__ void __launch_bounds__(256) k_points_raster(const RasterPointArgs a) {
  // grid: x = 32-word segments of a corner row, y = groups of 8 rows (one per warp), z = planes
  const int lane = threadIdx.x & 31;
  const int w = blockIdx.x * 32 + lane, cy = blockIdx.y * 8 + (threadIdx.x >> 5), cz = a.plane_lo + blockIdx.z;
  if (cy >= a.EY) return;
  const size_t row = (size_t)cz * a.EY + cy;
  uint32_t m = w < a.Wc ? __ldg(a.act + row * a.EW + w) : 0u;
  uint32_t incl = (uint32_t)__popc(m);
  const uint32_t mine = incl;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (!m) return;
  size_t id = (size_t)__ldg(a.seg + row * a.NS + blockIdx.x).z + (incl - mine);
  while (m) {
    const int b = __ffs(m) - 1;
    m &= m - 1;
    if (id < a.points_cap) {
      float* p = a.points + 3 * id;
      const int cx = 32 * w + b;
      p[0] = corner_coord(a.geom, 0, cx + a.coff[0], cy + a.coff[1], cz + a.coff[2]);
      p[1] = corner_coord(a.geom, 1, cx + a.coff[0], cy + a.coff[1], cz + a.coff[2]);
      p[2] = corner_coord(a.geom, 2, cx + a.coff[0], cy + a.coff[1], cz + a.coff[2]);
    } else {
      atomicOr(a.flags, (unsigned long long)kFlagBufferOverflow);
    }
    ++id;
  }
}

}  // namespace cbr
