// k_sweep.cuh — K2a (count) and K3 (emit): corner-centric ownership, shared through shared memory,
// swept along z.
//
// Reference: the "Create vertices" / "Create faces" part of the hot loop (txx:179-202), the two-plane
// vertex lookup it relies on (VertexLookupMap h:273-313, txx:128-131,155-161,186-191), AddVertex without
// the projection (txx:257-276) and AddQuadFace (txx:279-332).
//
// Ownership rule: a lattice corner gets its vertex from the first voxel, in raster
// order, among the 2x2x2 voxels around it that has an active face touching it.  For the eight inside
// bits i0..i7 of that block (block raster order, coordinates clamped to the image) this is the closed form
//     i0 == 0 : the first inside voxel
//     i0 == 1 : 0 if !(i1&i2&i4), else 1 if !(i3&i5), else 2 if !i6, else 3 if !i7, else none
// (equal to "first p with i_p & ~(i_{p^1} & i_{p^2} & i_{p^4})" for all 256 inputs, and for blocks clipped
// by the image border once an out-of-image alias hands its claim to its in-image twin;
// tests/test_ownership_rule.py enumerates both).  It is evaluated for 32 corners at a time, ONCE per
// corner word, by the thread that owns that corner word; the eight one-hot masks it yields are exactly the
// eight "my local corner l is new" masks of the <=4 voxel words around it, so neighbouring threads swap
// them through shared memory instead of re-deriving them (the first version did, at 3.5x the cost).
//
// A CTA is a grid of NTX x NTY threads = corner words (x) x corner rows (y) and walks along z, like the
// reference walks slices with its two lookup planes.  Per z step a thread loads the two new voxel rows
// of its corner word, evaluates the closed form for corner plane z+1, publishes what its -x / -y
// neighbours need, and assembles the 8 ownership masks + 6 face masks of "its" voxel word in slice z.
//   MODE_COUNT : popcounts -> one packed (faces << 16 | vertices) count per voxel word       (K2a)
//   MODE_EMIT  : vertex ids into dense shared-memory corner planes (double-buffered over z), points written
//                by the corner threads, faces compacted per warp and written with ids read from the
//                planes                                                                       (K3)
#pragma once
#include "cub_common.cuh"

namespace cub {

enum { MODE_COUNT = 0, MODE_EMIT = 1 };
enum { kEmitQuads = 0, kEmitTrisFixed = 1, kEmitScratchQuads = 2 };

struct SweepArgs {
  const uint32_t* bits;
  Grid g;
  int Wc;              // corner words per row = ceil((X+1)/32)
  int z_begin, z_end;  // local z range of voxel slices handled by the launch (count: scan range, emit: own range)
  int tz;              // slices per CTA
  // --- count
  uint32_t* counts;    // [Zl][Y][Wp] packed faces<<16 | vertices
  // --- emit
  const uint32_t* vofs;
  const uint32_t* fofs;
  Geom geom;
  int owner_z_min;     // lowest local z inside the scan range (z_begin-1, or z_begin at the image bottom)
  int own_z_top;       // local z of the top plane of the handle's own range (== z_end of the last chunk)
  uint32_t ghost_v, ghost_f;    // scan offsets of the first own vertex / face
  unsigned long long id_delta;  // (point id base - ghost_v) mod 2^64 : scan offset -> final id
  float* points;       // indexed by scan-relative vertex offset
  void* cells;         // final cells (IdT) or scratch quads (uint32 scan-relative ids)
  int mode;            // kEmit*
  int emit_ghost_points;
  const void* vol;     // for cell data (may be null)
  void* celldata;
  int pix_bytes;
};

// one row of the bitmask -> the c (voxel x = corner x) and l (voxel x-1) words of corner word cw,
// edge-replicated on both sides (corner word Wx only exists when X % 32 == 0: a replicate of voxel X-1)
__device__ __forceinline__ void load_cl(const uint32_t* __restrict__ row, int cw, int Wx, uint32_t& c, uint32_t& l) {
  if (cw < Wx) {
    c = __ldg(row + cw);
    const uint32_t carry = (cw == 0) ? (c & 1u) : (__ldg(row + cw - 1) >> 31);
    l = (c << 1) | carry;
  } else {
    const uint32_t carry = __ldg(row + Wx - 1) >> 31;
    c = 0u - carry;
    l = c;
  }
}

// closed-form first-touch owner of 32 corners: in[p] = inside word of block voxel p = qz*4+qy*2+qx
__device__ __forceinline__ void corner_owners(const uint32_t in[8], uint32_t own[8]) {
  uint32_t r = in[0];
#pragma unroll
  for (int p = 1; p < 8; ++p) {
    own[p] = in[p] & ~r;
    r |= in[p];
  }
  const uint32_t a = in[1] & in[2] & in[4];
  own[0] = in[0] & ~a;
  const uint32_t t = in[0] & a;
  const uint32_t b = in[3] & in[5];
  own[1] |= t & ~b;
  const uint32_t t2 = t & b;
  own[2] |= t2 & ~in[6];
  own[3] |= t2 & in[6] & ~in[7];
}

__host__ __device__ constexpr int face_corner(int f, int k) {
  // txx:197-202 / 219-233
  return f == 0 ? (k == 0 ? 0 : k == 1 ? 4 : k == 2 ? 7 : 3)
       : f == 1 ? (k == 0 ? 0 : k == 1 ? 1 : k == 2 ? 5 : 4)
       : f == 2 ? (k == 0 ? 1 : k == 1 ? 2 : k == 2 ? 6 : 5)
       : f == 3 ? (k == 0 ? 2 : k == 1 ? 3 : k == 2 ? 7 : 6)
       : f == 4 ? (k == 0 ? 0 : k == 1 ? 3 : k == 2 ? 2 : 1)
                : (k == 0 ? 4 : k == 1 ? 5 : k == 2 ? 6 : 7);
}

template <int NTX, int NTY, int MODE>
struct SweepSmem {
  static constexpr int NT = NTX * NTY;
  static constexpr int PXW = 32 * (NTX - 1) + 1;  // corner-plane width  (every corner a complete voxel word can touch)
  static constexpr int PY = NTY;                  // corner-plane height
  static_assert(PY * PXW <= 8192, "face queue items keep a 13-bit plane index");
  static constexpr int QCAP = 224;                // per-warp face queue (32-bit items)
  uint32_t ex[2][6][NT];                          // exchange: P0,P1,P4,P5, c word of slice z, bit-0 nibble
  // emit only
  uint32_t plane[MODE == MODE_EMIT ? 2 : 1][MODE == MODE_EMIT ? PY : 1][MODE == MODE_EMIT ? PXW : 1];
  uint4 ftab[2][8];                               // per-step plane offsets of the 4 corners of face f
  float xtab[MODE == MODE_EMIT ? PXW : 1];
  float ytab[MODE == MODE_EMIT ? PY : 1];
  uint32_t queue[MODE == MODE_EMIT ? NT / 32 : 1][MODE == MODE_EMIT ? QCAP : 1];
};

template <int NTX, int NTY, int MODE, typename IdT>
__global__ void __launch_bounds__(NTX* NTY) k_sweep(const SweepArgs a) {
  using S = SweepSmem<NTX, NTY, MODE>;
  constexpr int NT = S::NT;
  constexpr int LO = (MODE == MODE_EMIT) ? 1 : 0;    // low-side halo (owners at x-1 / y-1 of a face's corners)
  constexpr int TXW = NTX - 1 - 2 * LO, TY = NTY - 1 - 2 * LO;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  S& sm = *reinterpret_cast<S*>(smem_raw);

  const Grid& g = a.g;
  const int t = threadIdx.x;
  const int i = t % NTX, j = t / NTX;
  const int lane = t & 31, warp = t >> 5;
  const int w0 = blockIdx.x * TXW, y0 = blockIdx.y * TY;
  const int cw = w0 - LO + i, cy = y0 - LO + j;
  const bool corner_ok = cw >= 0 && cw < a.Wc && cy >= 0 && cy <= g.Y;
  const bool complete = i <= NTX - 2 && j <= NTY - 2;                       // all 8 masks of the voxel word available
  const bool voxel_ok = complete && cw >= 0 && cw < g.Wx && cy >= 0 && cy < g.Y;
  const bool interior = i >= LO && i <= NTX - 2 - LO && j >= LO && j <= NTY - 2 - LO;
  const int zs = a.z_begin + blockIdx.z * a.tz;
  const int ze = min(zs + a.tz, a.z_end);
  // first voxel slice whose masks are assembled: the emit sweep warms up on the slice below its range
  const int z_first = (MODE == MODE_EMIT) ? max(zs - 1, a.owner_z_min) : zs;

  // ---- per-thread constants ---------------------------------------------------------------------------
  const int cwl = min(max(cw, 0), a.Wc - 1);
  const int ya = min(max(cy - 1, 0), g.Y - 1), yb = min(max(cy, 0), g.Y - 1);
  const int XW = g.X >> 5, XB = g.X & 31;
  // corners that exist in this word (cx <= X)
  const uint32_t vcm = !corner_ok ? 0u : (cwl < XW ? ~0u : (cwl == XW ? ((2u << XB) - 1u) : 0u));
  const uint32_t m0 = (cwl == 0) ? 1u : 0u;                                  // the cx == 0 corner
  const uint32_t vc = (cw == g.Wx - 1 && XB) ? ((1u << XB) - 1u) : ~0u;      // valid voxels of the voxel word
  const size_t slice_words = (size_t)g.Y * g.Wp;
  const uint32_t* __restrict__ rowa = a.bits + (size_t)ya * g.Wp;
  const uint32_t* __restrict__ rowb = a.bits + (size_t)yb * g.Wp;

  auto local_slice = [&](int zl) {  // clamp a local slice index to the image (global) and to the buffer
    int zg = zl + g.zg0;
    zg = zg < 0 ? 0 : (zg > g.Zg - 1 ? g.Zg - 1 : zg);
    int z = zg - g.zg0;
    return z < 0 ? 0 : (z > g.Zl - 1 ? g.Zl - 1 : z);
  };

  if (MODE == MODE_EMIT) {
    for (int k = t; k < S::PXW; k += NT) sm.xtab[k] = corner_coord(a.geom.spacing[0], a.geom.origin[0], 32 * (w0 - LO) + k);
    for (int k = t; k < S::PY; k += NT) sm.ytab[k] = corner_coord(a.geom.spacing[1], a.geom.origin[1], y0 - LO + k);
  }

  // rolling voxel words: rows ya (index 0) and yb (index 1); "lo" = slice cz-1, "hi" = slice cz
  uint32_t lo_c[2] = {0, 0}, lo_l[2] = {0, 0}, hi_c[2], hi_l[2];
  uint32_t below_c = 0;  // row yb of slice cz-2 (the -z neighbour of the voxel word being assembled)
  {
    const size_t zo = (size_t)local_slice(z_first - 1) * slice_words;
    load_cl(rowa + zo, cwl, g.Wx, hi_c[0], hi_l[0]);
    load_cl(rowb + zo, cwl, g.Wx, hi_c[1], hi_l[1]);
  }
  // saved from the previous plane: own P6,P7, the +y neighbour's P4,P5, shift-in bits, active masks
  uint32_t s6 = 0, s7 = 0, s4u = 0, s5u = 0, sb6 = 0, sb4 = 0, actL_prev = 0, actU_prev = 0;

  for (int cz = z_first; cz <= ze; ++cz) {
    const int buf = cz & 1;
    // ---- 1. slide the window, load slice cz, closed form for corner plane cz ---------------------------
    below_c = lo_c[1];
    lo_c[0] = hi_c[0]; lo_c[1] = hi_c[1]; lo_l[0] = hi_l[0]; lo_l[1] = hi_l[1];
    {
      const size_t zo = (size_t)local_slice(cz) * slice_words;
      load_cl(rowa + zo, cwl, g.Wx, hi_c[0], hi_l[0]);
      load_cl(rowb + zo, cwl, g.Wx, hi_c[1], hi_l[1]);
    }
    uint32_t own[8];
    {
      const uint32_t in[8] = {lo_l[0], lo_c[0], lo_l[1], lo_c[1], hi_l[0], hi_c[0], hi_l[1], hi_c[1]};
      corner_owners(in, own);
      const int czg = cz + g.zg0;
      if (cy == 0) {  // block row cy-1 is outside the image: aliases hand over to their twins
        own[2] |= own[0]; own[3] |= own[1]; own[6] |= own[4]; own[7] |= own[5];
        own[0] = own[1] = own[4] = own[5] = 0;
      }
      if (czg == 0) {
        own[4] |= own[0]; own[5] |= own[1]; own[6] |= own[2]; own[7] |= own[3];
        own[0] = own[1] = own[2] = own[3] = 0;
      }
#pragma unroll
      for (int p = 0; p < 8; p += 2) {  // cx == 0: voxel x-1 is outside the image
        own[p + 1] |= own[p] & m0;
        own[p] &= ~m0;
      }
#pragma unroll
      for (int p = 0; p < 8; ++p) own[p] &= vcm;
    }
    sm.ex[buf][0][t] = own[0];
    sm.ex[buf][1][t] = own[1];
    sm.ex[buf][2][t] = own[4];
    sm.ex[buf][3][t] = own[5];
    sm.ex[buf][4][t] = lo_c[1];
    sm.ex[buf][5][t] = (own[0] & 1u) | ((own[2] & 1u) << 1) | ((own[4] & 1u) << 2) | ((own[6] & 1u) << 3) | ((lo_c[1] & 1u) << 4);
    if (MODE == MODE_EMIT && t < 6) {
      // plane offsets of the four corners of face t for the voxel slice z = cz-1 (ring slot of planes z, z+1)
      const int sl0 = (cz - 1) & 1, sl1 = cz & 1;
      auto off = [&](int l) {
        return (uint32_t)((((l >> 2) & 1 ? sl1 : sl0) * S::PY + ((0xCC >> l) & 1)) * S::PXW + ((0x66 >> l) & 1));
      };
      sm.ftab[buf][t] = make_uint4(off(kFaceCorners[t][0]), off(kFaceCorners[t][1]), off(kFaceCorners[t][2]),
                                   off(kFaceCorners[t][3]));
    }
    __syncthreads();

    const uint32_t actL = own[0] | own[1] | own[2] | own[3], actU = own[4] | own[5] | own[6] | own[7];
    if (cz > z_first) {
      // ---- 2. masks of voxel word (cw, cy) in slice z = cz-1 --------------------------------------------
      const int z = cz - 1;
      uint32_t p0u = 0, p1u = 0, p4u = 0, p5u = 0, upc = 0, nr = 0, nur = 0;
      if (complete) {
        const int tu = t + NTX;  // thread (i, j+1)
        p0u = sm.ex[buf][0][tu]; p1u = sm.ex[buf][1][tu]; p4u = sm.ex[buf][2][tu]; p5u = sm.ex[buf][3][tu];
        upc = sm.ex[buf][4][tu];
        nr = sm.ex[buf][5][t + 1];    // thread (i+1, j)
        nur = sm.ex[buf][5][tu + 1];  // thread (i+1, j+1)
      }
      uint32_t O[8], F[6];
      O[0] = s7;
      O[1] = (s6 >> 1) | (sb6 << 31);
      O[2] = (s4u >> 1) | (sb4 << 31);
      O[3] = s5u;
      O[4] = own[3];
      O[5] = (own[2] >> 1) | ((nr >> 1) << 31);
      O[6] = (p0u >> 1) | (nur << 31);
      O[7] = p1u;
      const uint32_t c = voxel_ok ? (lo_c[1] & vc) : 0u;
      const uint32_t r = (lo_c[1] >> 1) | ((nr >> 4) << 31);
      F[0] = c & ~lo_l[1];
      F[1] = c & ~lo_c[0];
      F[2] = c & ~r;
      F[3] = c & ~upc;
      F[4] = c & ~below_c;
      F[5] = c & ~hi_c[1];
      if (!voxel_ok) {
#pragma unroll
        for (int l = 0; l < 8; ++l) O[l] = 0;
      }

      if (MODE == MODE_COUNT) {
        if (voxel_ok && z >= zs) {
          uint32_t nv = 0, nf = 0;
#pragma unroll
          for (int l = 0; l < 8; ++l) nv += __popc(O[l]);
#pragma unroll
          for (int f = 0; f < 6; ++f) nf += __popc(F[f]);
          a.counts[(size_t)z * slice_words + (size_t)cy * g.Wp + cw] = (nf << 16) | nv;
        }
      } else {
        // ---- 3. vertex ids of the owned corners -> corner planes z and z+1 -------------------------------
        const size_t wi = (size_t)z * slice_words + (size_t)cy * g.Wp + cw;
        uint32_t U = O[0] | O[1] | O[2] | O[3] | O[4] | O[5] | O[6] | O[7];
        if (U) {
          uint32_t id = __ldg(a.vofs + wi);
          const int s0 = z & 1, s1 = (z + 1) & 1;
          uint32_t* p00 = &sm.plane[s0][j][32 * i];      // corner (x, y, z) of voxel bit 0
          uint32_t* p01 = &sm.plane[s0][j + 1][32 * i];
          uint32_t* p10 = &sm.plane[s1][j][32 * i];
          uint32_t* p11 = &sm.plane[s1][j + 1][32 * i];
          while (U) {
            const int b = __ffs(U) - 1;
            U &= U - 1;
            const uint32_t bit = 1u << b;
            if (O[0] & bit) { p00[b] = id; ++id; }
            if (O[1] & bit) { p00[b + 1] = id; ++id; }
            if (O[2] & bit) { p01[b + 1] = id; ++id; }
            if (O[3] & bit) { p01[b] = id; ++id; }
            if (O[4] & bit) { p10[b] = id; ++id; }
            if (O[5] & bit) { p10[b + 1] = id; ++id; }
            if (O[6] & bit) { p11[b + 1] = id; ++id; }
            if (O[7] & bit) { p11[b] = id; ++id; }
          }
        }
        __syncthreads();

        // ---- 4. points of corner plane z (complete now), written by the corner threads -------------------
        // lower half = owned by slice z-1, upper half = owned by slice z
        if (interior && corner_ok) {
          const bool first_own = (z == a.z_begin), ghost = a.emit_ghost_points != 0;
          uint32_t m = 0;
          if (z >= zs) m = actU_prev | ((z > a.z_begin || ghost) ? actL_prev : 0u);
          else if (ghost && first_own == false && z == a.z_begin - 1) m = actU_prev;  // ghost slice: its own lower corners
          const int s0 = z & 1;
          const float pz = corner_coord(a.geom.spacing[2], a.geom.origin[2], z + g.zg0);
          const float py = sm.ytab[j];
          while (m) {
            const int b = __ffs(m) - 1;
            m &= m - 1;
            const uint32_t id = sm.plane[s0][j][32 * i + b];
            float* p = a.points + 3 * (size_t)id;
            p[0] = sm.xtab[32 * i + b];
            p[1] = py;
            p[2] = pz;
          }
          // the top plane of the handle's range has no later step: its lower half is complete already
          if (cz == ze && ze == a.own_z_top) {
            uint32_t mt = actL;
            const int s1 = cz & 1;
            const float pzt = corner_coord(a.geom.spacing[2], a.geom.origin[2], cz + g.zg0);
            while (mt) {
              const int b = __ffs(mt) - 1;
              mt &= mt - 1;
              const uint32_t id = sm.plane[s1][j][32 * i + b];
              float* p = a.points + 3 * (size_t)id;
              p[0] = sm.xtab[32 * i + b];
              p[1] = py;
              p[2] = pzt;
            }
          }
        }

        // ---- 5. faces of slice z: compact per warp, then one face per lane --------------------------------
        if (z >= zs) {
          const bool emit = interior && voxel_ok;
          uint32_t nf = 0;
          if (emit) {
#pragma unroll
            for (int f = 0; f < 6; ++f) nf += __popc(F[f]);
          }
          uint32_t incl = nf;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
          }
          const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
          if (total) {
            const uint32_t fbase = nf ? (__ldg(a.fofs + wi) - a.ghost_f) : 0u;
            if (total <= (uint32_t)S::QCAP) {
              // item = plane index of the voxel's corner 0 (13 bits) | face << 13 | source lane << 16 | rank << 21
              uint32_t pos = incl - nf, rank = 0;
              uint32_t U2 = nf ? (F[0] | F[1] | F[2] | F[3] | F[4] | F[5]) : 0u;
              uint32_t* q = sm.queue[warp];
              const uint32_t tag = (uint32_t)(j * S::PXW + 32 * i) | ((uint32_t)lane << 16);
              while (U2) {
                const int b = __ffs(U2) - 1;
                U2 &= U2 - 1;
                const uint32_t bit = 1u << b;
                const uint32_t where = tag + (uint32_t)b;
#pragma unroll
                for (int f = 0; f < 6; ++f)
                  if (F[f] & bit) { q[pos++] = where | ((uint32_t)f << 13) | (rank++ << 21); }
              }
              __syncwarp();
              const uint32_t rounds = (total + 31) >> 5;
              for (uint32_t rr = 0; rr < rounds; ++rr) {
                const uint32_t s = rr * 32 + lane;
                const bool live = s < total;
                const uint32_t it = live ? q[s] : 0u;
                const uint32_t fb = __shfl_sync(0xffffffffu, fbase, (it >> 16) & 31u);
                if (!live) continue;
                const uint32_t base = it & 0x1fffu, f = (it >> 13) & 7u;
                const uint4 o = sm.ftab[buf][f];
                const uint32_t* pl = &sm.plane[0][0][0];
                const uint32_t q0 = pl[base + o.x], q1 = pl[base + o.y], q2 = pl[base + o.z], q3 = pl[base + o.w];
                const size_t fidx = (size_t)fb + (it >> 21);
                if (a.mode == kEmitScratchQuads) {
                  reinterpret_cast<uint4*>(a.cells)[fidx] = make_uint4(q0, q1, q2, q3);
                } else {
                  const IdT v0 = (IdT)(q0 + a.id_delta), v1 = (IdT)(q1 + a.id_delta), v2 = (IdT)(q2 + a.id_delta),
                            v3 = (IdT)(q3 + a.id_delta);
                  IdT* cdst = reinterpret_cast<IdT*>(a.cells);
                  if (a.mode == kEmitQuads) {
                    cdst += fidx * 4;
                    if (sizeof(IdT) == 4) {
                      *reinterpret_cast<uint4*>(cdst) = make_uint4((uint32_t)v0, (uint32_t)v1, (uint32_t)v2, (uint32_t)v3);
                    } else {
                      cdst[0] = v0; cdst[1] = v1; cdst[2] = v2; cdst[3] = v3;
                    }
                  } else {
                    // unprojected quad: both diagonals are equal, `>=` takes the first split (txx:298-302)
                    cdst += fidx * 6;
                    cdst[0] = v0; cdst[1] = v1; cdst[2] = v3;
                    cdst[3] = v1; cdst[4] = v2; cdst[5] = v3;
                  }
                }
                if (a.celldata) {
                  const uint32_t rem = base % (uint32_t)S::PXW, jj = base / (uint32_t)S::PXW;
                  const size_t vx = (size_t)(32 * (w0 - LO)) + rem, vy = (size_t)(y0 - LO) + jj;
                  const unsigned char* src = reinterpret_cast<const unsigned char*>(a.vol) +
                                             (((size_t)z * g.Y + vy) * g.X + vx) * a.pix_bytes;
                  const bool two = (a.mode != kEmitQuads);
                  unsigned char* dst = reinterpret_cast<unsigned char*>(a.celldata) + (two ? 2 * fidx : fidx) * a.pix_bytes;
                  for (int bb = 0; bb < a.pix_bytes; ++bb) {
                    const unsigned char v = src[bb];
                    dst[bb] = v;
                    if (two) dst[a.pix_bytes + bb] = v;
                  }
                }
              }
              __syncwarp();
            } else {
              // more faces than the queue holds (noise-like data): every lane writes its own faces
              uint32_t fi = fbase;
              uint32_t U2 = nf ? (F[0] | F[1] | F[2] | F[3] | F[4] | F[5]) : 0u;
              const uint32_t* pl = &sm.plane[0][0][0];
              while (U2) {
                const int b = __ffs(U2) - 1;
                U2 &= U2 - 1;
                const uint32_t bit = 1u << b;
                const uint32_t base = (uint32_t)(j * S::PXW + 32 * i + b);
                for (int f = 0; f < 6; ++f) {
                  if (!(F[f] & bit)) continue;
                  const uint4 o = sm.ftab[buf][f];
                  const uint32_t q0 = pl[base + o.x], q1 = pl[base + o.y], q2 = pl[base + o.z], q3 = pl[base + o.w];
                  const size_t fidx = fi++;
                  if (a.mode == kEmitScratchQuads) {
                    reinterpret_cast<uint4*>(a.cells)[fidx] = make_uint4(q0, q1, q2, q3);
                  } else {
                    const IdT v0 = (IdT)(q0 + a.id_delta), v1 = (IdT)(q1 + a.id_delta), v2 = (IdT)(q2 + a.id_delta),
                              v3 = (IdT)(q3 + a.id_delta);
                    IdT* cdst = reinterpret_cast<IdT*>(a.cells);
                    if (a.mode == kEmitQuads) {
                      cdst += fidx * 4;
                      cdst[0] = v0; cdst[1] = v1; cdst[2] = v2; cdst[3] = v3;
                    } else {
                      cdst += fidx * 6;
                      cdst[0] = v0; cdst[1] = v1; cdst[2] = v3;
                      cdst[3] = v1; cdst[4] = v2; cdst[5] = v3;
                    }
                  }
                  if (a.celldata) {
                    const size_t vx = (size_t)cw * 32 + b;
                    const unsigned char* src = reinterpret_cast<const unsigned char*>(a.vol) +
                                               (((size_t)z * g.Y + cy) * g.X + vx) * a.pix_bytes;
                    const bool two = (a.mode != kEmitQuads);
                    unsigned char* dst = reinterpret_cast<unsigned char*>(a.celldata) + (two ? 2 * fidx : fidx) * a.pix_bytes;
                    for (int bb = 0; bb < a.pix_bytes; ++bb) {
                      const unsigned char v = src[bb];
                      dst[bb] = v;
                      if (two) dst[a.pix_bytes + bb] = v;
                    }
                  }
                }
              }
            }
          }
        }
      }

      // ---- 6. what the next plane needs from this one ------------------------------------------------------
      s4u = p4u; s5u = p5u;
      sb6 = (nr >> 3) & 1u;
      sb4 = (nur >> 2) & 1u;
    } else if (complete) {
      // first plane of the sweep: nothing to assemble yet, only remember the upper halves
      const int tu = t + NTX;
      s4u = sm.ex[buf][2][tu]; s5u = sm.ex[buf][3][tu];
      sb6 = (sm.ex[buf][5][t + 1] >> 3) & 1u;
      sb4 = (sm.ex[buf][5][tu + 1] >> 2) & 1u;
    }
    s6 = own[6]; s7 = own[7];
    actL_prev = actL; actU_prev = actU;
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
