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
// A CTA is a grid of NTX x NTY threads; thread (i, j) owns corner word i of R consecutive corner rows and
// walks along z, like the reference walks slices with its two lookup planes.  Per z step a thread loads
// the R+1 voxel rows under its corner rows for the new slice, evaluates the closed form for corner plane
// z+1, publishes what its -x / -y neighbours need (6 words per thread, whatever R), and assembles the 8
// ownership masks + 6 face masks of its R voxel words in slice z.
//   MODE_COUNT : popcounts -> one packed (faces << 16 | vertices) count per voxel word       (K2a)
//   MODE_EMIT  : vertex ids into dense shared-memory corner planes (double-buffered over z), points written
//                by the corner threads, faces compacted per warp and written, one face per lane, with ids
//                read from the planes                                                         (K3)
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
  uint32_t ghost_f;    // scan offset of the first own face
  unsigned long long id_delta;  // (point id base - ghost vertices) mod 2^64 : scan offset -> final id
  float* points;       // indexed by scan-relative vertex offset
  void* cells;         // final cells (IdT) or scratch quads (uint32 scan-relative ids)
  int mode;            // kEmit*
  int emit_ghost_points;
  const void* vol;     // for cell data (may be null)
  void* celldata;
  int pix_bytes;
};

template <int NTX_, int NTY_, int R_, int MODE_>
struct SweepCfg {
  static constexpr int NTX = NTX_, NTY = NTY_, R = R_, MODE = MODE_;
  static constexpr int NT = NTX * NTY;               // live threads
  static constexpr int NTP = (NT + 31) / 32 * 32;    // launched threads
  static constexpr int NW = NTP / 32;
  static constexpr int LO = (MODE == MODE_EMIT) ? 1 : 0;  // low-side halo: a face's corner can belong to x-1 / y-1
  static constexpr int CR = NTY * R;                 // corner rows per CTA
  static constexpr int TXW = NTX - 1 - 2 * LO;       // voxel words (x) whose results the CTA produces
  static constexpr int TY = CR - 1 - 2 * LO;         // voxel rows  (y)
  static constexpr int PXW = 32 * (NTX - 1) + 1;     // corner plane: every corner a complete voxel word can touch
  static constexpr int PY = CR;
  static constexpr int QCAP = 96 + 64 * R;           // per-warp face queue (32-bit items)
  static_assert(R >= 1 && R <= 4, "nibble word holds 5 bits per row, face items 7 bits of (lane, row)");
  static_assert(MODE != MODE_EMIT || PY * PXW <= 16384, "face queue items keep a 14-bit plane index");
};

template <typename C>
struct SweepSmem {
  uint32_t ex[2][6][C::NT];  // exchange: P0,P1,P4,P5 of the thread's lowest corner row, its c word, bit-0 nibbles
  uint32_t plane[C::MODE == MODE_EMIT ? 2 : 1][C::MODE == MODE_EMIT ? C::PY : 1][C::MODE == MODE_EMIT ? C::PXW : 1];
  uint4 ftab[2][8];          // per-step plane offsets of the 4 corners of face f
  float xtab[C::MODE == MODE_EMIT ? C::PXW : 1];
  float ytab[C::MODE == MODE_EMIT ? C::PY : 1];
  uint32_t queue[C::MODE == MODE_EMIT ? C::NW : 1][C::MODE == MODE_EMIT ? C::QCAP : 1];
  uint32_t fbase[C::MODE == MODE_EMIT ? C::NW : 1][C::MODE == MODE_EMIT ? 32 * C::R : 1];
};

// closed-form first-touch owner of 32 corners: in[p] = inside word of block voxel p = qz*4+qy*2+qx;
// vm = corners of this word that exist
__device__ __forceinline__ void corner_owners(const uint32_t in[8], uint32_t vm, uint32_t own[8]) {
  uint32_t r = in[0];
#pragma unroll
  for (int p = 1; p < 8; ++p) {
    own[p] = in[p] & ~r & vm;
    r |= in[p];
  }
  const uint32_t a = in[1] & in[2] & in[4];
  own[0] = in[0] & ~a & vm;
  const uint32_t t = in[0] & a & vm;
  const uint32_t b = in[3] & in[5];
  own[1] |= t & ~b;
  const uint32_t t2 = t & b;
  own[2] |= t2 & ~in[6];
  own[3] |= t2 & in[6] & ~in[7];
}

template <typename C, typename IdT>
__device__ __forceinline__ void write_cell(const SweepArgs& a, size_t fidx, uint32_t q0, uint32_t q1, uint32_t q2, uint32_t q3) {
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

__device__ __forceinline__ void write_celldata(const SweepArgs& a, size_t fidx, size_t voxel) {
  const unsigned char* src = reinterpret_cast<const unsigned char*>(a.vol) + voxel * a.pix_bytes;
  const bool two = (a.mode != kEmitQuads);
  unsigned char* dst = reinterpret_cast<unsigned char*>(a.celldata) + (two ? 2 * fidx : fidx) * a.pix_bytes;
  for (int bb = 0; bb < a.pix_bytes; ++bb) {
    const unsigned char v = src[bb];
    dst[bb] = v;
    if (two) dst[a.pix_bytes + bb] = v;
  }
}

template <typename C, typename IdT>
__global__ void __launch_bounds__(C::NTP) k_sweep(const SweepArgs a) {
  using S = SweepSmem<C>;
  constexpr int NTX = C::NTX, NTY = C::NTY, R = C::R, MODE = C::MODE, NT = C::NT, LO = C::LO, CR = C::CR;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  S& sm = *reinterpret_cast<S*>(smem_raw);

  const Grid& g = a.g;
  const int t = threadIdx.x;
  const bool alive = t < NT;
  const int tt = alive ? t : 0;
  const int i = tt % NTX, j = tt / NTX;
  const int lane = t & 31, warp = t >> 5;
  const int w0 = blockIdx.x * C::TXW, y0 = blockIdx.y * C::TY;
  const int cw = w0 - LO + i;
  const int rr0 = j * R;                // first corner row of the thread, CTA-relative
  const int cy0 = y0 - LO + rr0;        // ... and in the image
  const int zs = a.z_begin + blockIdx.z * a.tz;
  const int ze = min(zs + a.tz, a.z_end);
  // first voxel slice whose masks are assembled: the emit sweep warms up on the slice below its range
  const int z_first = (MODE == MODE_EMIT) ? max(zs - 1, a.owner_z_min) : zs;

  // ---- per-thread constants ---------------------------------------------------------------------------
  const int XW = g.X >> 5, XB = g.X & 31;
  const bool col_ok = alive && cw >= 0 && cw < a.Wc;
  const uint32_t vmx = !col_ok ? 0u : (cw < XW ? ~0u : ((2u << XB) - 1u));       // corners cx <= X of this word
  const uint32_t vc = (cw == g.Wx - 1 && XB) ? ((1u << XB) - 1u) : ~0u;           // voxels x < X of this word
  const bool first = cw <= 0, synth = cw >= g.Wx;  // no voxel word on the left / this corner word is the x = X replicate
  const int cwl = min(max(cw, 0), g.Wx - 1);
  const bool xcomplete = alive && i <= NTX - 2;    // the +x neighbour thread exists
  const bool word_ok = xcomplete && cw >= 0 && cw < g.Wx;
  const bool xinterior = i >= LO && i <= NTX - 2 - LO;
  // the corner column one past the interior belongs to this CTA when there is no tile to the right
  const bool xpoints = col_ok && (xinterior || (LO && i == NTX - 1 - LO && cw >= g.Wx));
  const size_t slice_words = (size_t)g.Y * g.Wp;
  int rowoff[R + 1];
#pragma unroll
  for (int k = 0; k <= R; ++k) rowoff[k] = min(max(cy0 - 1 + k, 0), g.Y - 1) * g.Wp + cwl;
  const int dprev = (first || synth) ? 0 : 1;

  // local slice indices are clamped to the image (globally) and, for memory safety, to the local buffer
  const int zlo = max(0, -g.zg0), zhi = min(g.Zl - 1, g.Zg - 1 - g.zg0);
  const int wi0 = cy0 * g.Wp + cw;      // word index of voxel row 0 of this thread inside a slice (valid rows only)
  // per-row predicates, one bit per k: voxel row complete & inside the image / emits faces / corner row emits points
  uint32_t rowok_bits = 0, face_bits = 0, ypoint_bits = 0;
#pragma unroll
  for (int k = 0; k < R; ++k) {
    const int rr = rr0 + k, cy = cy0 + k;
    const bool ok = word_ok && rr <= CR - 2 && cy >= 0 && cy < g.Y;
    const bool yint = rr >= LO && rr <= CR - 2 - LO;
    if (ok) rowok_bits |= 1u << k;
    if (ok && xinterior && yint) face_bits |= 1u << k;
    // rows one past the interior belong to this CTA when there is no tile above
    if (xpoints && cy >= 0 && cy <= g.Y && (yint || (LO && rr == CR - 1 - LO && cy >= g.Y))) ypoint_bits |= 1u << k;
  }
  const int exo = alive ? t : 0;
  // raw words (this corner word and the one to its left) of the R+1 window rows of one slice; fetched one
  // step ahead of their use so that the L2/HBM latency overlaps the previous step
  auto fetch_slice = [&](int zl, uint32_t (&cwv)[R + 1], uint32_t (&pwv)[R + 1]) {
    const uint32_t* __restrict__ sl = a.bits + (size_t)min(max(zl, zlo), zhi) * slice_words;
#pragma unroll
    for (int k = 0; k <= R; ++k) {
      cwv[k] = __ldg(sl + rowoff[k]);
      pwv[k] = __ldg(sl + rowoff[k] - dprev);
    }
  };
  // -> c (voxel x = corner x) and l (voxel x-1) words, edge-replicated in x
  auto decode_slice = [&](const uint32_t (&cwv)[R + 1], const uint32_t (&pwv)[R + 1], uint32_t (&c)[R + 1], uint32_t (&l)[R + 1]) {
#pragma unroll
    for (int k = 0; k <= R; ++k) {
      const uint32_t rep = 0u - (cwv[k] >> 31);
      const uint32_t pw = first ? (cwv[k] << 31) : pwv[k];
      c[k] = synth ? rep : cwv[k];
      l[k] = synth ? rep : __funnelshift_l(pw, cwv[k], 1);
    }
  };

  if (MODE == MODE_EMIT) {
    for (int k = t; k < C::PXW; k += C::NTP) sm.xtab[k] = corner_coord(a.geom.spacing[0], a.geom.origin[0], 32 * (w0 - LO) + k);
    for (int k = t; k < C::PY; k += C::NTP) sm.ytab[k] = corner_coord(a.geom.spacing[1], a.geom.origin[1], y0 - LO + k);
  }

  // rolling voxel words of the window rows: "lo" = slice cz-1, "hi" = slice cz
  uint32_t lo_c[R + 1], lo_l[R + 1], hi_c[R + 1], hi_l[R + 1];
  uint32_t below_c[R];  // voxel rows of slice cz-2 (the -z neighbours of the voxel words being assembled)
#pragma unroll
  for (int k = 0; k <= R; ++k) lo_c[k] = lo_l[k] = 0;
  uint32_t nx_c[R + 1], nx_p[R + 1];  // prefetched raw words of the next slice
  fetch_slice(z_first - 1, nx_c, nx_p);
  decode_slice(nx_c, nx_p, hi_c, hi_l);
  fetch_slice(z_first, nx_c, nx_p);
  uint32_t vnx[R], fnx[R];            // prefetched scan offsets of the voxel words of the next assembled slice
#pragma unroll
  for (int k = 0; k < R; ++k) vnx[k] = fnx[k] = 0;
  // kept from the previous plane: own P4..P7 per row, the +y neighbour's P4,P5, the neighbours' nibbles, active masks
  uint32_t sv[R][4], su4 = 0, su5 = 0, nr_prev = 0, nur_prev = 0, actL_prev[R], actU_prev[R];
#pragma unroll
  for (int k = 0; k < R; ++k) {
    sv[k][0] = sv[k][1] = sv[k][2] = sv[k][3] = 0;
    actL_prev[k] = actU_prev[k] = 0;
  }

  for (int cz = z_first; cz <= ze; ++cz) {
    const int buf = cz & 1;
    // ---- 1. slide the window, load slice cz, closed form for corner plane cz ---------------------------
#pragma unroll
    for (int k = 0; k < R; ++k) below_c[k] = lo_c[k + 1];
#pragma unroll
    for (int k = 0; k <= R; ++k) { lo_c[k] = hi_c[k]; lo_l[k] = hi_l[k]; }
    decode_slice(nx_c, nx_p, hi_c, hi_l);
    if (cz < ze) fetch_slice(cz + 1, nx_c, nx_p);
    uint32_t vpre[R], fpre[R];  // scan offsets of this thread's voxel words in slice cz-1 (fetched one step early)
    if (MODE == MODE_EMIT) {
#pragma unroll
      for (int k = 0; k < R; ++k) { vpre[k] = vnx[k]; fpre[k] = fnx[k]; }
      if (cz < ze) {  // the next step assembles slice cz
        const uint32_t* __restrict__ vz = a.vofs + (size_t)cz * slice_words;
        const uint32_t* __restrict__ fz = a.fofs + (size_t)cz * slice_words;
#pragma unroll
        for (int k = 0; k < R; ++k) {
          if ((rowok_bits >> k) & 1u) {
            vnx[k] = __ldg(vz + wi0 + k * g.Wp);
            fnx[k] = __ldg(fz + wi0 + k * g.Wp);
          }
        }
      }
    }
    const int czg = cz + g.zg0;
    uint32_t own[R][8];
    uint32_t nib = 0;
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const int cy = cy0 + k;
      const uint32_t in[8] = {lo_l[k], lo_c[k], lo_l[k + 1], lo_c[k + 1], hi_l[k], hi_c[k], hi_l[k + 1], hi_c[k + 1]};
      corner_owners(in, (cy >= 0 && cy <= g.Y) ? vmx : 0u, own[k]);
      if (cy == 0) {  // block row cy-1 is outside the image: aliases hand over to their twins
        own[k][2] |= own[k][0]; own[k][3] |= own[k][1]; own[k][6] |= own[k][4]; own[k][7] |= own[k][5];
        own[k][0] = own[k][1] = own[k][4] = own[k][5] = 0;
      }
      if (czg == 0) {
        own[k][4] |= own[k][0]; own[k][5] |= own[k][1]; own[k][6] |= own[k][2]; own[k][7] |= own[k][3];
        own[k][0] = own[k][1] = own[k][2] = own[k][3] = 0;
      }
      if (cw == 0) {  // cx == 0: voxel x-1 is outside the image
#pragma unroll
        for (int p = 0; p < 8; p += 2) {
          own[k][p + 1] |= own[k][p] & 1u;
          own[k][p] &= ~1u;
        }
      }
      nib |= ((own[k][0] & 1u) | ((own[k][2] & 1u) << 1) | ((own[k][4] & 1u) << 2) | ((own[k][6] & 1u) << 3) |
              ((lo_c[k + 1] & 1u) << 4)) << (5 * k);
    }
    uint32_t* __restrict__ exw = &sm.ex[buf][0][exo];
    if (alive) {
      exw[0 * NT] = own[0][0];
      exw[1 * NT] = own[0][1];
      exw[2 * NT] = own[0][4];
      exw[3 * NT] = own[0][5];
      exw[4 * NT] = lo_c[1];
      exw[5 * NT] = nib;
    }
    if (MODE == MODE_EMIT && t < 6) {
      // plane offsets of the four corners of face t for the voxel slice z = cz-1 (planes z and z+1)
      const int sl0 = (cz - 1) & 1, sl1 = cz & 1;
      auto off = [&](int l) {
        return (uint32_t)((((l >> 2) & 1 ? sl1 : sl0) * C::PY + ((0xCC >> l) & 1)) * C::PXW + ((0x66 >> l) & 1));
      };
      sm.ftab[buf][t] = make_uint4(off(kFaceCorners[t][0]), off(kFaceCorners[t][1]), off(kFaceCorners[t][2]),
                                   off(kFaceCorners[t][3]));
    }
    __syncthreads();

    uint32_t actL[R], actU[R];
#pragma unroll
    for (int k = 0; k < R; ++k) {
      actL[k] = own[k][0] | own[k][1] | own[k][2] | own[k][3];
      actU[k] = own[k][4] | own[k][5] | own[k][6] | own[k][7];
    }
    // what the neighbours published for this plane
    uint32_t p0u = 0, p1u = 0, p4u = 0, p5u = 0, upc = 0, nr = 0, nur = 0;
    if (xcomplete) {
      nr = exw[5 * NT + 1];                       // thread (i+1, j)
      if (j <= NTY - 2) {
        const uint32_t* __restrict__ exu = exw + NTX;  // thread (i, j+1)
        p0u = exu[0 * NT]; p1u = exu[1 * NT]; p4u = exu[2 * NT]; p5u = exu[3 * NT];
        upc = exu[4 * NT];
        nur = exu[5 * NT + 1];                    // thread (i+1, j+1)
      }
    }

    if (cz > z_first) {
      const int z = cz - 1;  // the voxel slice being assembled
      // ownership masks of voxel row k (corner rows k and k+1 of planes z [kept] and z+1 [fresh])
      auto assemble = [&](int k, uint32_t (&O)[8]) {
        const bool top = (k == R - 1);
        const int ku = top ? 0 : k + 1;
        const uint32_t s4 = top ? su4 : sv[ku][0], s5 = top ? su5 : sv[ku][1];
        const uint32_t f0 = top ? p0u : own[ku][0], f1 = top ? p1u : own[ku][1];
        const uint32_t nup = top ? nur : (nr >> (5 * ku)), nup_prev = top ? nur_prev : (nr_prev >> (5 * ku));
        O[0] = sv[k][3];
        O[1] = __funnelshift_r(sv[k][2], nr_prev >> (5 * k + 3), 1);
        O[2] = __funnelshift_r(s4, nup_prev >> 2, 1);
        O[3] = s5;
        O[4] = own[k][3];
        O[5] = __funnelshift_r(own[k][2], nr >> (5 * k + 1), 1);
        O[6] = __funnelshift_r(f0, nup, 1);
        O[7] = f1;
      };
      auto row_ok = [&](int k) {  // voxel row k of this thread exists and all its masks are available
        return ((rowok_bits >> k) & 1u) != 0;
      };
      auto faces_of = [&](int k, uint32_t (&F)[6]) {
        const uint32_t c = lo_c[k + 1] & vc;
        F[0] = c & ~lo_l[k + 1];
        F[1] = c & ~lo_c[k];
        F[2] = c & ~__funnelshift_r(lo_c[k + 1], nr >> (5 * k + 4), 1);
        F[3] = c & ~((k == R - 1) ? upc : lo_c[(k == R - 1) ? 0 : k + 2]);
        F[4] = c & ~below_c[k];
        F[5] = c & ~hi_c[k + 1];
      };

      if (MODE == MODE_COUNT) {
#pragma unroll
        for (int k = 0; k < R; ++k) {
          if (row_ok(k)) {
            uint32_t O[8], F[6];
            assemble(k, O);
            faces_of(k, F);
            uint32_t nv = 0, nf = 0;
#pragma unroll
            for (int l = 0; l < 8; ++l) nv += __popc(O[l]);
#pragma unroll
            for (int f = 0; f < 6; ++f) nf += __popc(F[f]);
            (a.counts + (size_t)z * slice_words)[wi0 + k * g.Wp] = (nf << 16) | nv;
          }
        }
      } else {
        // ---- 3. vertex ids of the owned corners -> corner planes z and z+1 -------------------------------
        const int s0 = z & 1, s1 = (z + 1) & 1;
#pragma unroll
        for (int k = 0; k < R; ++k) {
          if (!row_ok(k)) continue;
          uint32_t O[8];
          assemble(k, O);
          uint32_t U = O[0] | O[1] | O[2] | O[3] | O[4] | O[5] | O[6] | O[7];
          if (U) {
            uint32_t id = vpre[k];
            uint32_t* p00 = &sm.plane[s0][rr0 + k][32 * i];  // corner (x, y, z) of voxel bit 0
            uint32_t* p01 = p00 + C::PXW;
            uint32_t* p10 = &sm.plane[s1][rr0 + k][32 * i];
            uint32_t* p11 = p10 + C::PXW;
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
        }
        __syncthreads();

        // ---- 4. points of corner plane z (complete now), written by the corner threads -------------------
        // lower half = owned by slice z-1, upper half = owned by slice z
        if (ypoint_bits) {
          const bool ghost = a.emit_ghost_points != 0;
          const bool want_upper = (z >= zs) || (ghost && z == a.z_begin - 1);
          const bool want_lower = (z >= zs) && (z > a.z_begin || ghost);
          const bool top_plane = (cz == ze && ze == a.z_end);  // no later step: its lower half is complete already
          const float pz = corner_coord(a.geom.spacing[2], a.geom.origin[2], z + g.zg0);
          const float pzt = corner_coord(a.geom.spacing[2], a.geom.origin[2], cz + g.zg0);
#pragma unroll
          for (int k = 0; k < R; ++k) {
            const int rr = rr0 + k;
            if (!((ypoint_bits >> k) & 1u)) continue;
            const float py = sm.ytab[rr];
            uint32_t m = (want_upper ? actU_prev[k] : 0u) | (want_lower ? actL_prev[k] : 0u);
            while (m) {
              const int b = __ffs(m) - 1;
              m &= m - 1;
              const uint32_t id = sm.plane[s0][rr][32 * i + b];
              float* p = a.points + 3 * (size_t)id;
              p[0] = sm.xtab[32 * i + b];
              p[1] = py;
              p[2] = pz;
            }
            if (top_plane) {
              uint32_t mt = actL[k];
              while (mt) {
                const int b = __ffs(mt) - 1;
                mt &= mt - 1;
                const uint32_t id = sm.plane[s1][rr][32 * i + b];
                float* p = a.points + 3 * (size_t)id;
                p[0] = sm.xtab[32 * i + b];
                p[1] = py;
                p[2] = pzt;
              }
            }
          }
        }

        // ---- 5. faces of slice z: compact per warp, then one face per lane --------------------------------
        if (z >= zs) {
          uint32_t nf = 0;
          uint32_t nfk[R];
#pragma unroll
          for (int k = 0; k < R; ++k) {
            nfk[k] = 0;
            if ((face_bits >> k) & 1u) {
              uint32_t F[6];
              faces_of(k, F);
#pragma unroll
              for (int f = 0; f < 6; ++f) nfk[k] += __popc(F[f]);
            }
            nf += nfk[k];
          }
          uint32_t incl = nf;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
          }
          const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
          if (total) {
            const uint32_t* pl = &sm.plane[0][0][0];
            const bool queued = total <= (uint32_t)C::QCAP;
            uint32_t pos = incl - nf;
            uint32_t* q = sm.queue[warp];
#pragma unroll
            for (int k = 0; k < R; ++k) {
              if (!nfk[k]) continue;
              uint32_t F[6];
              faces_of(k, F);
              const uint32_t fb = fpre[k] - a.ghost_f;
              uint32_t U2 = F[0] | F[1] | F[2] | F[3] | F[4] | F[5];
              const uint32_t where0 = (uint32_t)((rr0 + k) * C::PXW + 32 * i);
              if (queued) {
                // item = plane index of the voxel's corner 0 (14 bits) | face << 14 | (lane, row) << 17 | rank << 24
                sm.fbase[warp][lane * R + k] = fb;
                const uint32_t tag = where0 | ((uint32_t)(lane * R + k) << 17);
                uint32_t rank = 0;
                while (U2) {
                  const int b = __ffs(U2) - 1;
                  U2 &= U2 - 1;
                  const uint32_t bit = 1u << b;
                  const uint32_t where = tag + (uint32_t)b;
#pragma unroll
                  for (int f = 0; f < 6; ++f)
                    if (F[f] & bit) { q[pos++] = where | ((uint32_t)f << 14) | (rank++ << 24); }
                }
              } else {
                // more faces than the queue holds (noise-like data): every lane writes its own faces
                size_t fi = fb;
                while (U2) {
                  const int b = __ffs(U2) - 1;
                  U2 &= U2 - 1;
                  const uint32_t bit = 1u << b;
                  const uint32_t base = where0 + (uint32_t)b;
                  for (int f = 0; f < 6; ++f) {
                    if (!(F[f] & bit)) continue;
                    const uint4 o = sm.ftab[buf][f];
                    write_cell<C, IdT>(a, fi, pl[base + o.x], pl[base + o.y], pl[base + o.z], pl[base + o.w]);
                    if (a.celldata) write_celldata(a, fi, ((size_t)z * g.Y + (cy0 + k)) * g.X + (size_t)cw * 32 + b);
                    ++fi;
                  }
                }
              }
            }
            if (queued) {
              __syncwarp();
              for (uint32_t s = lane; s < total; s += 32) {
                const uint32_t it = q[s];
                const uint32_t base = it & 0x3fffu, f = (it >> 14) & 7u;
                const size_t fidx = (size_t)sm.fbase[warp][(it >> 17) & 127u] + (it >> 24);
                const uint4 o = sm.ftab[buf][f];
                write_cell<C, IdT>(a, fidx, pl[base + o.x], pl[base + o.y], pl[base + o.z], pl[base + o.w]);
                if (a.celldata) {
                  const uint32_t rem = base % (uint32_t)C::PXW, jj = base / (uint32_t)C::PXW;
                  write_celldata(a, fidx, ((size_t)z * g.Y + (size_t)(y0 - LO) + jj) * g.X + (size_t)(32 * (w0 - LO)) + rem);
                }
              }
              __syncwarp();
            }
          }
        }
      }
    }

    // ---- 6. what the next plane needs from this one --------------------------------------------------------
#pragma unroll
    for (int k = 0; k < R; ++k) {
      sv[k][0] = own[k][4]; sv[k][1] = own[k][5]; sv[k][2] = own[k][6]; sv[k][3] = own[k][7];
      actL_prev[k] = actL[k];
      actU_prev[k] = actU[k];
    }
    su4 = p4u; su5 = p5u;
    nr_prev = nr; nur_prev = nur;
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
