// k_count_scan.cuh — K2b: single-pass exclusive scan of the per-word (vertex, face) counts.
//
// Replaces, for id assignment, the per-slice vertex lookup of the reference
// (VertexLookupMap h:273-313, used at txx:186-191): ids follow from an exclusive prefix sum, in
// voxel-raster order, of "corners first touched by this voxel" (vertex ids, nextVertexId
// txx:116,189-190) and "faces of this voxel" (cell ids, nextCellId txx:117,197-202).
//
// K2a (k_sweep.cuh, MODE_COUNT) leaves one packed count (faces << 16 | vertices) per 32-voxel word.
// This kernel turns them into the two exclusive-offset arrays vofs / fofs: each thread takes 16
// consecutive words (four 16-byte loads), a block-wide scan combines the 4096 words of a tile, and tiles
// are chained with decoupled look-back (flag+value packed in one 64-bit descriptor per tile and
// quantity; tile numbers are handed out by an atomic ticket so a tile only ever waits for tiles that
// started before it).  HBM-bound: N/8 bytes in, 2*N/8 bytes out.
#pragma once
#include "cub_common.cuh"

namespace cub {

constexpr int kScanThreads = 256;
constexpr int kScanWordsPerThread = 16;
constexpr int kScanTileWords = kScanThreads * kScanWordsPerThread;

constexpr uint64_t kFlagShift = 62;
constexpr uint64_t kFlagAggregate = 1ull << kFlagShift;
constexpr uint64_t kFlagPrefix = 2ull << kFlagShift;
constexpr uint64_t kValueMask = (1ull << kFlagShift) - 1ull;

struct ScanState {
  unsigned long long* status_v;  // [n_tiles] descriptors, vertices
  unsigned long long* status_f;  // [n_tiles] descriptors, faces
  unsigned int* ticket;          // tile ticket counter
  unsigned long long* totals;    // [0] = vertices, [1] = faces in the scan range
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

// scan range: words [word_begin, word_begin + n_words) of the padded [Zl][Y][Wp] layout (whole slices)
__global__ void __launch_bounds__(kScanThreads)
    k_count_scan(const uint32_t* __restrict__ counts, uint32_t* __restrict__ vofs, uint32_t* __restrict__ fofs, Grid g,
                 size_t word_begin, size_t n_words, ScanState st) {
  __shared__ unsigned int s_tile;
  __shared__ unsigned long long s_warp[kScanThreads / 32];
  __shared__ unsigned long long s_excl_v, s_excl_f;

  if (threadIdx.x == 0) s_tile = atomicAdd(st.ticket, 1u);
  __syncthreads();
  const int tile = (int)s_tile;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  const size_t gw = (size_t)tile * kScanTileWords + (size_t)threadIdx.x * kScanWordsPerThread;  // within the range
  uint32_t cv[kScanWordsPerThread], cf[kScanWordsPerThread];
#pragma unroll
  for (int j = 0; j < kScanWordsPerThread; ++j) cv[j] = cf[j] = 0;

  if (gw < n_words) {
    const size_t aw = word_begin + gw;  // absolute padded word index (a multiple of 4; rows are Wp = 4k words)
    const int w0 = (int)(aw % (size_t)g.Wp);
#pragma unroll
    for (int v4 = 0; v4 < kScanWordsPerThread / 4; ++v4) {
      if (gw + 4 * v4 < n_words) {
        const uint4 q = __ldg(reinterpret_cast<const uint4*>(counts + aw) + v4);
        const uint32_t c[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int w = (w0 + 4 * v4 + j) % g.Wp;
          if (w < g.Wx) {  // pad words of a row are never written by K2a
            cv[4 * v4 + j] = c[j] & 0xffffu;
            cf[4 * v4 + j] = c[j] >> 16;
          }
        }
      }
    }
  }

  // thread totals packed as (faces << 32 | vertices); a tile holds < 2^20 of either
  unsigned long long mine = 0;
#pragma unroll
  for (int j = 0; j < kScanWordsPerThread; ++j) mine += ((unsigned long long)cf[j] << 32) | cv[j];
  unsigned long long incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  unsigned long long warp_excl = 0, block_total = 0;
#pragma unroll
  for (int i = 0; i < kScanThreads / 32; ++i) {
    const unsigned long long t = s_warp[i];
    if (i < warp) warp_excl += t;
    block_total += t;
  }
  const unsigned long long agg_v = block_total & 0xffffffffull, agg_f = block_total >> 32;

  if (warp == 0) {
    if (lane == 0) {
      st_relaxed(st.status_v + tile, (tile == 0 ? kFlagPrefix : kFlagAggregate) | agg_v);
      st_relaxed(st.status_f + tile, (tile == 0 ? kFlagPrefix : kFlagAggregate) | agg_f);
    }
    unsigned long long ev = 0, ef = 0;
    if (tile > 0) {
      ev = lookback(st.status_v, tile, lane);
      ef = lookback(st.status_f, tile, lane);
      if (lane == 0) {
        st_relaxed(st.status_v + tile, kFlagPrefix | (ev + agg_v));
        st_relaxed(st.status_f + tile, kFlagPrefix | (ef + agg_f));
      }
    }
    if (lane == 0) {
      s_excl_v = ev;
      s_excl_f = ef;
      if ((size_t)(tile + 1) * kScanTileWords >= n_words) {  // last tile: grand totals
        st.totals[0] = ev + agg_v;
        st.totals[1] = ef + agg_f;
      }
    }
  }
  __syncthreads();

  if (gw < n_words) {
    const unsigned long long excl = warp_excl + (incl - mine);
    uint32_t v = (uint32_t)(s_excl_v + (excl & 0xffffffffull));
    uint32_t f = (uint32_t)(s_excl_f + (excl >> 32));
    const size_t aw = word_begin + gw;
#pragma unroll
    for (int v4 = 0; v4 < kScanWordsPerThread / 4; ++v4) {
      if (gw + 4 * v4 < n_words) {
        uint4 ov, of;
        ov.x = v; of.x = f; v += cv[4 * v4 + 0]; f += cf[4 * v4 + 0];
        ov.y = v; of.y = f; v += cv[4 * v4 + 1]; f += cf[4 * v4 + 1];
        ov.z = v; of.z = f; v += cv[4 * v4 + 2]; f += cf[4 * v4 + 2];
        ov.w = v; of.w = f; v += cv[4 * v4 + 3]; f += cf[4 * v4 + 3];
        reinterpret_cast<uint4*>(vofs + aw)[v4] = ov;
        reinterpret_cast<uint4*>(fofs + aw)[v4] = of;
      }
    }
  }
}

// reads the exclusive offsets at up to 2 word positions (slab own-range boundaries)
__global__ void k_gather_marks(const uint32_t* __restrict__ vofs, const uint32_t* __restrict__ fofs, size_t mark0,
                               size_t mark1, unsigned long long* out /* [2..5] */) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    out[2] = (mark0 != (size_t)-1) ? vofs[mark0] : 0ull;
    out[3] = (mark0 != (size_t)-1) ? fofs[mark0] : 0ull;
    out[4] = (mark1 != (size_t)-1) ? vofs[mark1] : ~0ull;
    out[5] = (mark1 != (size_t)-1) ? fofs[mark1] : ~0ull;
  }
}

}  // namespace cub
