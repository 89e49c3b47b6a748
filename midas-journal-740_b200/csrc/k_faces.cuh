// k_faces.cuh — K3c: quad / triangle emission straight into the mesh's cell buffer, one thread per
// 32-voxel word.
//
// Reference: faceHasQuad (txx:164-173), the face -> corner table (txx:197-202, 219-233) and AddQuadFace
// (txx:279-332).  Cell ids follow voxel raster x face index (nextCellId, txx:117): cell index =
// fofs[word] + rank inside the word.  The four vertex ids of a face come from the corner -> id map:
//     slot(corner) = cofs[corner word] + popc(act[corner word] & bits below)      id = perm[slot]
// where the <= 8 corner words around the voxel word are loaded once per word.  No shared memory, no
// synchronisation: words without faces (most of them) exit after seven bitmask loads.
#pragma once
#include "cub_common.cuh"

namespace cub {

enum { kEmitQuads = 0, kEmitTrisFixed = 1, kEmitScratchQuads = 2 };

struct FaceArgs {
  const uint32_t* bits;
  Grid g;
  int EY, EW;
  int z_begin, z_end;          // local voxel slices whose faces are emitted (the handle's own range)
  const uint32_t* fofs;        // entry lattice, exclusive scan of face counts
  const uint32_t* act;
  const uint32_t* cofs;
  const uint32_t* perm;        // slot -> scan-relative vertex id
  uint32_t ghost_f;            // scan offset of the first own face
  unsigned long long id_delta; // scan-relative vertex id -> final id (mod 2^64)
  void* cells;                 // final cells (IdT) or scratch quads (uint32 scan-relative ids)
  int mode;                    // kEmit*
  const void* vol;             // for cell data (may be null)
  void* celldata;
  int pix_bytes;
};

template <typename IdT>
__device__ __forceinline__ void write_cell(const FaceArgs& a, size_t fidx, uint32_t q0, uint32_t q1, uint32_t q2, uint32_t q3) {
  if (a.mode == kEmitScratchQuads) {
    reinterpret_cast<uint4*>(a.cells)[fidx] = make_uint4(q0, q1, q2, q3);
    return;
  }
  const IdT v0 = (IdT)(q0 + a.id_delta), v1 = (IdT)(q1 + a.id_delta), v2 = (IdT)(q2 + a.id_delta), v3 = (IdT)(q3 + a.id_delta);
  IdT* c = reinterpret_cast<IdT*>(a.cells);
  if (a.mode == kEmitQuads) {
    c += fidx * 4;
    if (sizeof(IdT) == 4) {
      *reinterpret_cast<uint4*>(c) = make_uint4((uint32_t)v0, (uint32_t)v1, (uint32_t)v2, (uint32_t)v3);
    } else {
      c[0] = v0; c[1] = v1; c[2] = v2; c[3] = v3;
    }
  } else {
    // unprojected quad: both diagonals are equal, `>=` takes the first split (txx:298-302)
    c += fidx * 6;
    c[0] = v0; c[1] = v1; c[2] = v3;
    c[3] = v1; c[4] = v2; c[5] = v3;
  }
}

__device__ __forceinline__ void write_celldata(const FaceArgs& a, size_t fidx, size_t voxel) {
  const unsigned char* src = reinterpret_cast<const unsigned char*>(a.vol) + voxel * a.pix_bytes;
  const bool two = (a.mode != kEmitQuads);
  unsigned char* dst = reinterpret_cast<unsigned char*>(a.celldata) + (two ? 2 * fidx : fidx) * a.pix_bytes;
  for (int bb = 0; bb < a.pix_bytes; ++bb) {
    const unsigned char v = src[bb];
    dst[bb] = v;
    if (two) dst[a.pix_bytes + bb] = v;
  }
}

template <typename IdT>
__global__ void __launch_bounds__(256) k_faces(const FaceArgs a) {
  const Grid& g = a.g;
  // one thread per voxel word of the padded bitmask layout, own slices only
  const size_t slice_words = (size_t)g.Y * g.Wp;
  const size_t n = slice_words * (size_t)(a.z_end - a.z_begin);
  const size_t gi = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gi >= n) return;
  const int zl = a.z_begin + (int)(gi / slice_words);
  const int rem = (int)(gi - (size_t)(zl - a.z_begin) * slice_words);
  const int y = rem / g.Wp, w = rem - y * g.Wp;
  if (w >= g.Wx) return;

  // ---- face masks of the word (txx:164-173; clamped neighbours: no face on the image border) --------------
  const int zgl = zl + g.zg0;
  const int ym = max(y - 1, 0), yp = min(y + 1, g.Y - 1);
  const int zm = min(max(max(zgl - 1, 0) - g.zg0, 0), g.Zl - 1), zp = max(min(min(zgl + 1, g.Zg - 1) - g.zg0, g.Zl - 1), 0);
  const uint32_t* __restrict__ row = a.bits + (size_t)zl * slice_words + (size_t)y * g.Wp;
  const uint32_t c0 = __ldg(row + w);
  const uint32_t XB = g.X & 31;
  const uint32_t vc = (w == g.Wx - 1 && XB) ? ((1u << XB) - 1u) : ~0u;
  const uint32_t c = c0 & vc;
  if (c == 0) return;  // no inside voxel, no face
  const uint32_t prev = (w == 0) ? (c0 << 31) : __ldg(row + w - 1);
  const uint32_t next = (w == g.Wx - 1) ? (c0 >> 31) : __ldg(row + w + 1);
  uint32_t F[6];
  F[0] = c & ~__funnelshift_l(prev, c0, 1);
  F[1] = c & ~__ldg(a.bits + (size_t)zl * slice_words + (size_t)ym * g.Wp + w);
  F[2] = c & ~__funnelshift_r(c0, next, 1);
  F[3] = c & ~__ldg(a.bits + (size_t)zl * slice_words + (size_t)yp * g.Wp + w);
  F[4] = c & ~__ldg(a.bits + (size_t)zm * slice_words + (size_t)y * g.Wp + w);
  F[5] = c & ~__ldg(a.bits + (size_t)zp * slice_words + (size_t)y * g.Wp + w);
  uint32_t U = F[0] | F[1] | F[2] | F[3] | F[4] | F[5];
  if (U == 0) return;

  // ---- the corner words around this voxel word ------------------------------------------------------------
  const size_t plane = (size_t)a.EY * a.EW;
  const size_t e00 = ((size_t)zl * a.EY + y) * a.EW + w;  // corner word (w, y, z)
  uint32_t A[2][2], C[2][2], Cn[2][2];                     // [oz][oy]: active mask, slot base, slot base of word w+1
#pragma unroll
  for (int oz = 0; oz < 2; ++oz)
#pragma unroll
    for (int oy = 0; oy < 2; ++oy) {
      const size_t e = e00 + oz * plane + (size_t)oy * a.EW;
      A[oz][oy] = __ldg(a.act + e);
      C[oz][oy] = __ldg(a.cofs + e);
      Cn[oz][oy] = __ldg(a.cofs + e + 1);  // EW >= Wc + ... : entry w+1 always exists (padded rows)
    }
  size_t fi = (size_t)(__ldg(a.fofs + e00) - a.ghost_f);

  while (U) {
    const int b = __ffs(U) - 1;
    U &= U - 1;
    const uint32_t bit = 1u << b, below = bit - 1u;
    // vertex ids of the 8 corners of voxel b; local l -> (ox, oy, oz) as in txx:236-254
    uint32_t vid[8];
#pragma unroll
    for (int oz = 0; oz < 2; ++oz)
#pragma unroll
      for (int oy = 0; oy < 2; ++oy) {
        const uint32_t s0 = C[oz][oy] + __popc(A[oz][oy] & below);                        // corner x
        const uint32_t s1 = (b == 31) ? Cn[oz][oy] : s0 + ((A[oz][oy] >> b) & 1u);        // corner x+1
        const int l0 = oz * 4 + (oy ? 3 : 0), l1 = oz * 4 + (oy ? 2 : 1);
        vid[l0] = s0;
        vid[l1] = s1;
      }
    // which corners are needed (vertexHasQuad): local l touches the faces in its three directions
    const bool f0 = F[0] & bit, f1 = F[1] & bit, f2 = F[2] & bit, f3 = F[3] & bit, f4 = F[4] & bit, f5 = F[5] & bit;
    if (a.perm) {
      const bool need[8] = {f0 || f1 || f4, f1 || f2 || f4, f2 || f3 || f4, f0 || f3 || f4,
                            f0 || f1 || f5, f1 || f2 || f5, f2 || f3 || f5, f0 || f3 || f5};
#pragma unroll
      for (int l = 0; l < 8; ++l)
        if (need[l]) vid[l] = __ldg(a.perm + vid[l]);
    }
    const size_t voxel = ((size_t)zl * g.Y + y) * g.X + (size_t)w * 32 + b;
    if (f0) { write_cell<IdT>(a, fi, vid[0], vid[4], vid[7], vid[3]); if (a.celldata) write_celldata(a, fi, voxel); ++fi; }
    if (f1) { write_cell<IdT>(a, fi, vid[0], vid[1], vid[5], vid[4]); if (a.celldata) write_celldata(a, fi, voxel); ++fi; }
    if (f2) { write_cell<IdT>(a, fi, vid[1], vid[2], vid[6], vid[5]); if (a.celldata) write_celldata(a, fi, voxel); ++fi; }
    if (f3) { write_cell<IdT>(a, fi, vid[2], vid[3], vid[7], vid[6]); if (a.celldata) write_celldata(a, fi, voxel); ++fi; }
    if (f4) { write_cell<IdT>(a, fi, vid[0], vid[3], vid[2], vid[1]); if (a.celldata) write_celldata(a, fi, voxel); ++fi; }
    if (f5) { write_cell<IdT>(a, fi, vid[4], vid[5], vid[6], vid[7]); if (a.celldata) write_celldata(a, fi, voxel); ++fi; }
  }
}

// K5: triangle split of projected quads (AddQuadFace txx:286-321): reads the four PROJECTED points
// back, squared diagonal lengths in fp64 from the fp32 points in axis order,
// `>=` tie -> first split.
template <typename IdT>
__global__ void __launch_bounds__(256) k_split_quads(const uint4* __restrict__ quads, const float* __restrict__ points,
                                                     IdT* __restrict__ tris, size_t n_quads,
                                                     unsigned long long id_delta) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_quads) return;
  const uint4 q = quads[i];
  const uint32_t id[4] = {q.x, q.y, q.z, q.w};
  float p[4][3];
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int c = 0; c < 3; ++c) p[k][c] = __ldg(points + 3 * (size_t)id[k] + c);
  double d02 = 0.0, d13 = 0.0;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const double a = __dadd_rn((double)p[0][c], -(double)p[2][c]);
    d02 = __dadd_rn(d02, __dmul_rn(a, a));
    const double b = __dadd_rn((double)p[1][c], -(double)p[3][c]);
    d13 = __dadd_rn(d13, __dmul_rn(b, b));
  }
  IdT v[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) v[k] = (IdT)((unsigned long long)id[k] + id_delta);
  IdT* c = tris + i * 6;
  if (d02 >= d13) {
    c[0] = v[0]; c[1] = v[1]; c[2] = v[3];
    c[3] = v[1]; c[4] = v[2]; c[5] = v[3];
  } else {
    c[0] = v[0]; c[1] = v[1]; c[2] = v[2];
    c[3] = v[0]; c[4] = v[2]; c[5] = v[3];
  }
}

}  // namespace cub
