// k_count_scan.cuh — K2b: single-pass exclusive scan of the per-entry (vertex, face, active-corner) counts.
//
// Replaces, for id assignment, the per-slice vertex lookup of the reference
// (VertexLookupMap h:273-313, used at txx:186-191): ids follow from an exclusive prefix sum, in
// voxel-raster order, of "corners first touched by this voxel" (vertex ids, nextVertexId
// txx:116,189-190) and "faces of this voxel" (cell ids, nextCellId txx:117,197-202).  The third
// quantity, active corners per corner word in corner-raster order, indexes the corner -> id map.
//
// K2a (k_sweep.cuh, MODE_COUNT) leaves one packed count per entry of the [Zl+1][EY][EW] lattice
// (owned corners | faces << 10 | active corners << 20).  This kernel turns them into the three
// exclusive-offset arrays vofs / fofs / cofs: each thread takes 16 consecutive entries (four 16-byte
// loads), a block-wide scan combines 4096 entries at a time (the three counts travel packed in one
// 64-bit word, 21 bits each), and the (large) tiles are chained with decoupled look-back (flag+value packed in one
// 64-bit descriptor per tile and quantity; tile numbers are handed out by an atomic ticket so a tile only
// ever waits for tiles that started before it).  HBM-bound: 4 bytes in, 12 bytes out per entry.
#pragma once
#include "cub_common.cuh"

namespace cub {

constexpr int kScanThreads = 256;
constexpr int kScanPerThread = 4;   // one 16-byte load per thread: a warp reads / writes 512 contiguous bytes
constexpr int kScanTile = kScanThreads * kScanPerThread;

constexpr uint64_t kFlagShift = 62;
constexpr uint64_t kFlagAggregate = 1ull << kFlagShift;
constexpr uint64_t kFlagPrefix = 2ull << kFlagShift;
constexpr uint64_t kValueMask = (1ull << kFlagShift) - 1ull;

struct ScanArgs {
  const uint32_t* cnt;
  uint32_t* vofs;
  uint32_t* fofs;
  uint32_t* cofs;
  size_t e_begin, n;             // scanned entries [e_begin, e_begin + n): whole planes of the lattice
  unsigned plane_entries;        // EY * EW
  unsigned plane_lo;             // active corners are counted from this local plane on (ghost plane below: not)
  unsigned long long* status;    // [3][n_tiles] descriptors
  unsigned n_tiles;
  size_t tile;                   // entries per tile (a multiple of kScanTile)
  unsigned int* ticket;
  unsigned long long* totals;    // [0] vertices, [1] faces, [2] active corners of the scanned range
};

__device__ __forceinline__ unsigned long long ld_relaxed(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// warp 0 only: sum of the aggregates of all tiles before `tile`
__device__ __forceinline__ unsigned long long lookback(const unsigned long long* status, int tile, int lane) {
  unsigned long long exclusive = 0;
  int pos = tile - 1;
  while (true) {
    const int idx = pos - lane;
    unsigned long long d = kFlagPrefix;  // virtual tiles before tile 0: inclusive prefix 0
    if (idx >= 0) {
      d = ld_relaxed(status + idx);
      while ((d >> kFlagShift) == 0) d = ld_relaxed(status + idx);
    }
    const unsigned has_prefix = __ballot_sync(0xffffffffu, (d >> kFlagShift) == 2);
    const int first = has_prefix ? (__ffs(has_prefix) - 1) : 32;
    unsigned long long v = (lane <= first) ? (d & kValueMask) : 0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    exclusive += v;
    if (has_prefix) break;
    pos -= 32;
  }
  return exclusive;
}

// Two passes over a LARGE tile per CTA (args.tile entries, a multiple of kScanTile; a few hundred tiles in all):
//   pass 1  sums the tile (loads only) -> publishes the aggregate -> look-back gives the tile's exclusive prefix
//   pass 2  re-reads the tile (L2), block-scans it 4096 entries at a time and writes the three offset arrays.
// The look-back walks over every tile that is in flight without a finished prefix, so its cost is per TILE, not
// per byte: with 4096-entry tiles (r1-e) it dominated (0.25 ms for 0.6 GB); one large tile per resident CTA pays
// it once.
__global__ void __launch_bounds__(kScanThreads) k_count_scan(const ScanArgs a) {
  __shared__ unsigned int s_tile;
  __shared__ unsigned long long s_warp[kScanThreads / 32];
  __shared__ unsigned long long s_excl[3];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  if (threadIdx.x == 0) s_tile = atomicAdd(a.ticket, 1u);
  __syncthreads();
  const int tile = (int)s_tile;
  if (tile >= (int)a.n_tiles) return;
  const size_t t0 = (size_t)tile * a.tile;                       // first entry of the tile, within the range
  const size_t t1 = t0 + a.tile < a.n ? t0 + a.tile : a.n;

  // three 21-bit fields in one 64-bit word: vertices | faces << 21 | active corners << 42
  auto widen = [](uint32_t p) {
    return (unsigned long long)(p & 0x3ffu) | ((unsigned long long)((p >> 10) & 0x3ffu) << 21) |
           ((unsigned long long)(p >> 20) << 42);
  };
  // 16 consecutive entries of this thread in the 4096-entry block starting at g (masked + zero padded)
  auto load16 = [&](size_t g, uint32_t (&c)[kScanPerThread]) {
    const size_t g0 = g + (size_t)threadIdx.x * kScanPerThread;
    const size_t e = a.e_begin + g0;  // absolute entry index, a multiple of 4 (EW is)
#pragma unroll
    for (int v4 = 0; v4 < kScanPerThread / 4; ++v4) {
      uint4 q = make_uint4(0, 0, 0, 0);
      if (g0 + 4 * v4 < t1) {
        q = __ldg(reinterpret_cast<const uint4*>(a.cnt + e) + v4);
        // a 4-entry group never straddles planes; active corners below plane_lo belong to the slab underneath
        const bool ghost = (unsigned)((e + 4 * v4) / a.plane_entries) < a.plane_lo;
        const uint32_t keep = ghost ? 0xfffffu : 0xffffffffu;
        q.x &= keep; q.y &= keep; q.z &= keep; q.w &= keep;
      }
      c[4 * v4 + 0] = q.x; c[4 * v4 + 1] = q.y; c[4 * v4 + 2] = q.z; c[4 * v4 + 3] = q.w;
    }
  };

  // ---- pass 1: tile aggregate (fields can exceed 21 bits over a large tile: three separate sums) ------------
  unsigned long long sum[3] = {0, 0, 0};
  for (size_t g = t0; g < t1; g += kScanTile) {
    uint32_t c[kScanPerThread];
    load16(g, c);
    unsigned long long m = 0;
#pragma unroll
    for (int j = 0; j < kScanPerThread; ++j) m += widen(c[j]);
    sum[0] += m & 0x1fffffull; sum[1] += (m >> 21) & 0x1fffffull; sum[2] += m >> 42;
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum[k] += __shfl_xor_sync(0xffffffffu, sum[k], o);
  }
  __shared__ unsigned long long s_part[3][kScanThreads / 32];
  if (lane == 0) { s_part[0][warp] = sum[0]; s_part[1][warp] = sum[1]; s_part[2][warp] = sum[2]; }
  __syncthreads();
  if (warp < 3) {
    // warps 0..2 chain one quantity each, concurrently
    const int k = warp;
    unsigned long long aggk = 0;
#pragma unroll
    for (int i = 0; i < kScanThreads / 32; ++i) aggk += s_part[k][i];
    unsigned long long* st = a.status + (size_t)k * a.n_tiles;
    if (lane == 0) st_relaxed(st + tile, (tile == 0 ? kFlagPrefix : kFlagAggregate) | aggk);
    unsigned long long ex = 0;
    if (tile > 0) {
      ex = lookback(st, tile, lane);
      if (lane == 0) st_relaxed(st + tile, kFlagPrefix | (ex + aggk));
    }
    if (lane == 0) {
      s_excl[k] = ex;
      if (tile == (int)a.n_tiles - 1) a.totals[k] = ex + aggk;  // last tile: grand totals
    }
  }
  __syncthreads();
  unsigned long long run_v = s_excl[0], run_f = s_excl[1], run_k = s_excl[2];

  // ---- pass 2: offsets ------------------------------------------------------------------------------------
  for (size_t g = t0; g < t1; g += kScanTile) {
    uint32_t c[kScanPerThread];
    load16(g, c);
    unsigned long long mine = 0;
#pragma unroll
    for (int j = 0; j < kScanPerThread; ++j) mine += widen(c[j]);
    unsigned long long incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    __syncthreads();  // s_warp of the previous block has been read
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    unsigned long long warp_excl = 0, block_total = 0;
#pragma unroll
    for (int i = 0; i < kScanThreads / 32; ++i) {
      const unsigned long long t = s_warp[i];
      if (i < warp) warp_excl += t;
      block_total += t;
    }
    const unsigned long long excl = warp_excl + (incl - mine);
    uint32_t v = (uint32_t)(run_v + (excl & 0x1fffffull));
    uint32_t f = (uint32_t)(run_f + ((excl >> 21) & 0x1fffffull));
    uint32_t k = (uint32_t)(run_k + (excl >> 42));
    const size_t g0 = g + (size_t)threadIdx.x * kScanPerThread;
    const size_t e = a.e_begin + g0;
#pragma unroll
    for (int v4 = 0; v4 < kScanPerThread / 4; ++v4) {
      if (g0 + 4 * v4 < t1) {
        uint4 ov, of, ok;
#define CUB_STEP(field, j)                                   \
        ov.field = v; of.field = f; ok.field = k;            \
        v += c[4 * v4 + j] & 0x3ffu; f += (c[4 * v4 + j] >> 10) & 0x3ffu; k += c[4 * v4 + j] >> 20;
        CUB_STEP(x, 0) CUB_STEP(y, 1) CUB_STEP(z, 2) CUB_STEP(w, 3)
#undef CUB_STEP
        reinterpret_cast<uint4*>(a.vofs + e)[v4] = ov;
        reinterpret_cast<uint4*>(a.fofs + e)[v4] = of;
        reinterpret_cast<uint4*>(a.cofs + e)[v4] = ok;
      }
    }
    run_v += block_total & 0x1fffffull; run_f += (block_total >> 21) & 0x1fffffull; run_k += block_total >> 42;
  }
}

// reads the exclusive offsets at the first own entry (slab runs: what the ghost slice below contributed) and
// the number of active corners of the bottom plane of the own range (owned by the slab underneath)
__global__ void k_gather_marks(const uint32_t* __restrict__ vofs, const uint32_t* __restrict__ fofs,
                               const uint32_t* __restrict__ cofs, size_t mark0, size_t mark_c,
                               unsigned long long* out /* [3..5] */) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    out[3] = (mark0 != (size_t)-1) ? vofs[mark0] : 0ull;
    out[4] = (mark0 != (size_t)-1) ? fofs[mark0] : 0ull;
    out[5] = (mark_c != (size_t)-1) ? cofs[mark_c] : 0ull;
  }
}

}  // namespace cub
