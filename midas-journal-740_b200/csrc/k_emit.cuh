// k_emit.cuh — K3: vertex and face emission straight into the mesh's point and cell buffers.
//
// Reference: the "Create vertices" / "Create faces" part of the hot loop (txx:179-202), AddVertex
// without the projection (txx:257-276) and AddQuadFace (txx:279-332).
//
// A CTA owns a column of voxels (TX words x TY rows) and sweeps it along z, like the reference
// sweeps slices with its two lookup planes (txx:128-131,155-161) - but the planes here are dense
// shared-memory tables of vertex ids for the (32*TX+1) x (TY+1) lattice corners of the column,
// double-buffered over z, and they are FILLED from the ownership masks, not searched:
//   step (a)  every thread recomputes the masks of one word of slice z (the column plus a halo of
//             one word / one row, because the first-touch owner of a corner can be the -x / -y
//             neighbour), numbers the corners its voxels own (vofs[word] + rank, local order
//             0..7) and writes those ids into the corner planes z and z+1; interior threads also
//             write the vertex positions (the first-touch owner of a corner referenced by a face
//             of slice z always lies in slice z-1 or z, so after (a) both planes are complete);
//   step (b)  interior threads walk the set bits of their face masks and write each quad (or
//             its two triangles) with the four ids read from the planes.
// Output order is the reference's: cell ids follow voxel raster x face index, vertex ids follow
// first touch.  Writes of one row are contiguous in the output arrays.
#pragma once
#include "cub_common.cuh"

namespace cub {

__host__ __device__ constexpr int face_corner(int f, int k) {
  // txx:197-202 / 219-233
  return f == 0 ? (k == 0 ? 0 : k == 1 ? 4 : k == 2 ? 7 : 3)
       : f == 1 ? (k == 0 ? 0 : k == 1 ? 1 : k == 2 ? 5 : 4)
       : f == 2 ? (k == 0 ? 1 : k == 1 ? 2 : k == 2 ? 6 : 5)
       : f == 3 ? (k == 0 ? 2 : k == 1 ? 3 : k == 2 ? 7 : 6)
       : f == 4 ? (k == 0 ? 0 : k == 1 ? 3 : k == 2 ? 2 : 1)
                : (k == 0 ? 4 : k == 1 ? 5 : k == 2 ? 6 : 7);
}

enum { kEmitQuads = 0, kEmitTrisFixed = 1, kEmitScratchQuads = 2 };

struct EmitArgs {
  const uint32_t* bits;
  const uint32_t* vofs;
  const uint32_t* fofs;
  Grid g;
  Geom geom;
  int zs0, zs1;        // local z range whose faces are emitted (the handle's own range)
  int owner_z_min;     // lowest local z inside the scan range (zs0-1, or zs0 at the image bottom)
  int tz;              // slices per CTA sweep
  uint32_t ghost_f;    // scan offset of the first own face
  unsigned long long id_delta;  // (point id base - ghost vertices) mod 2^64: scan offset -> final id
  float* points;       // indexed by scan-relative vertex offset
  void* cells;         // final cells (IdT) or scratch quads (uint32 scan-relative ids)
  int mode;            // kEmit*
  int emit_ghost_points;
  const void* vol;     // for cell data (may be null)
  void* celldata;
  int pix_bytes;
};

template <int TX, int TY, typename IdT>
__global__ void __launch_bounds__((TX + 2) * (TY + 2)) k_emit(const EmitArgs a) {
  constexpr int PX = 32 * TX + 1;
  __shared__ uint32_t plane[2][TY + 1][PX];

  const Grid& g = a.g;
  const int t = threadIdx.x;
  const int ox = t % (TX + 2) - 1, oy = t / (TX + 2) - 1;
  const int w0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
  const int w = w0 + ox, y = y0 + oy;
  const bool exists = (w >= 0) && (w < g.Wx) && (y >= 0) && (y < g.Y);
  const bool interior = exists && ox >= 0 && ox < TX && oy >= 0 && oy < TY;
  const int zs = a.zs0 + blockIdx.z * a.tz;
  const int ze = min(zs + a.tz, a.zs1);
  const int zbeg = max(zs - 1, a.owner_z_min);

  for (int zz = zbeg; zz < ze; ++zz) {
    uint32_t F[6], O[8];
    uint32_t vb = 0, fb = 0;
    const bool emit_faces = interior && zz >= zs;
    if (exists) {
      Nbhd nb;
      load_nbhd(a.bits, g, w, y, zz, nb);
      compute_masks(nb, F, O);
      const size_t wi = row_index(g, y, zz) * (size_t)g.Wp + w;
      vb = __ldg(a.vofs + wi);
      if (emit_faces) fb = __ldg(a.fofs + wi);

      // ---- (a) number the owned corners, fill the planes, write the points -------------------
      const bool emit_pts = interior && (zz >= zs || (a.emit_ghost_points && blockIdx.z == 0));
      uint32_t U = O[0] | O[1] | O[2] | O[3] | O[4] | O[5] | O[6] | O[7];
      uint32_t id = vb;
      while (U) {
        const int b = __ffs(U) - 1;
        U &= U - 1;
        const int x = w * 32 + b;
#pragma unroll
        for (int l = 0; l < 8; ++l) {
          if ((O[l] >> b) & 1u) {
            const int cx = x + corner_ox(l), cy = y + corner_oy(l), cz = zz + corner_oz(l);
            const int cxl = cx - 32 * w0, cyl = cy - y0;
            if (cxl >= 0 && cxl <= 32 * TX && cyl >= 0 && cyl <= TY) plane[cz & 1][cyl][cxl] = id;
            if (emit_pts) {
              float* p = a.points + 3 * (size_t)id;
              p[0] = corner_coord(a.geom.spacing[0], a.geom.origin[0], cx);
              p[1] = corner_coord(a.geom.spacing[1], a.geom.origin[1], cy);
              p[2] = corner_coord(a.geom.spacing[2], a.geom.origin[2], cz + g.zg0);
            }
            ++id;
          }
        }
      }
    }
    __syncthreads();

    // ---- (b) faces of slice zz ------------------------------------------------------------------
    if (emit_faces) {
      uint32_t U = F[0] | F[1] | F[2] | F[3] | F[4] | F[5];
      size_t fi = (size_t)(fb - a.ghost_f);
      while (U) {
        const int b = __ffs(U) - 1;
        U &= U - 1;
        const int xl = (w - w0) * 32 + b, yl = y - y0;
        bool pixel_loaded = false;
        unsigned long long pixel = 0;
#pragma unroll
        for (int f = 0; f < 6; ++f) {
          if ((F[f] >> b) & 1u) {
            uint32_t q[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int l = face_corner(f, k);
              q[k] = plane[(zz + corner_oz(l)) & 1][yl + corner_oy(l)][xl + corner_ox(l)];
            }
            if (a.mode == kEmitScratchQuads) {
              reinterpret_cast<uint4*>(a.cells)[fi] = make_uint4(q[0], q[1], q[2], q[3]);
            } else {
              IdT v[4];
#pragma unroll
              for (int k = 0; k < 4; ++k) v[k] = (IdT)((unsigned long long)q[k] + a.id_delta);
              IdT* c = reinterpret_cast<IdT*>(a.cells);
              if (a.mode == kEmitQuads) {
                c += fi * 4;
                c[0] = v[0]; c[1] = v[1]; c[2] = v[2]; c[3] = v[3];
              } else {
                // unprojected quad: both diagonals are equal, `>=` takes the first split (txx:298-302)
                c += fi * 6;
                c[0] = v[0]; c[1] = v[1]; c[2] = v[3];
                c[3] = v[1]; c[4] = v[2]; c[5] = v[3];
              }
            }
            if (a.celldata) {
              if (!pixel_loaded) {
                const size_t vi = (row_index(g, y, zz) * (size_t)g.X + (size_t)(w * 32 + b)) * a.pix_bytes;
                const unsigned char* src = reinterpret_cast<const unsigned char*>(a.vol) + vi;
                for (int i = 0; i < a.pix_bytes; ++i) pixel |= (unsigned long long)src[i] << (8 * i);
                pixel_loaded = true;
              }
              const bool two = (a.mode != kEmitQuads);
              unsigned char* dst = reinterpret_cast<unsigned char*>(a.celldata) + (two ? 2 * fi : fi) * a.pix_bytes;
              for (int r = 0; r < (two ? 2 : 1); ++r)
                for (int i = 0; i < a.pix_bytes; ++i) dst[r * a.pix_bytes + i] = (unsigned char)(pixel >> (8 * i));
            }
            ++fi;
          }
        }
      }
    }
    __syncthreads();
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
