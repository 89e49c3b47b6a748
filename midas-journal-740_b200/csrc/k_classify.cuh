// k_classify.cuh — K1: one pass over the volume -> 1 bit / voxel inside mask.
//
// Reference: the suitability test `center < m_IsoSurfaceValue -> skip` (txx:139-141) and the
// neighbour test `GetPixel(offset) < m_IsoSurfaceValue` (txx:167) compare pixels with the iso
// value in the pixel type; both are the same predicate, so it is evaluated ONCE per voxel here
// and every later kernel works on the bitmask (1/32 of the bytes for float input).
//
// HBM-bound streaming read: algorithmic bytes = X*Y*Zl*sizeof(T) read (+ X*Y*Zl/8 written).
// Each warp owns a task of kWordsPerTask consecutive words of one row; every lane issues
// kWordsPerTask independent coalesced loads (lane b of load k reads voxel 32*(w0+k)+b: one full
// 128-byte line per warp instruction for 4-byte pixels, immediate offsets from one base pointer)
// before the first ballot, so a resident SM keeps warps * kWordsPerTask * 128 B in flight.  A
// ballot of the predicate IS the output word - no shuffles; lane 0 stores the task's words with
// 16-byte stores.  Budget: ~4.5 issued instructions per 128 B of input (r1 profile: the first
// version spent 21 and was issue-bound at 74 % of the HBM roofline).
// Lanes past the end of a ragged row re-read the row's last voxel, which produces the "replicate
// bit X-1" padding the later kernels rely on (generic path only).
#pragma once
#include "cub_common.cuh"

namespace cub {

constexpr int kWordsPerTask = 16;

template <typename T>
__device__ __forceinline__ T load_stream(const T* p) {
  return __ldcs(p);  // read-once data: evict-first
}

// FULL: X % 32 == 0 and Wx % kWordsPerTask == 0 (no clamping, no partial tasks)
template <typename T, bool FULL>
__global__ void __launch_bounds__(256) k_classify(const T* __restrict__ vol, uint32_t* __restrict__ bits, Grid g,
                                                  T iso, unsigned n_tasks, unsigned groups_per_row) {
  const int lane = threadIdx.x & 31;
  const unsigned warp0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned n_warps = (gridDim.x * blockDim.x) >> 5;
  for (unsigned task = warp0; task < n_tasks; task += n_warps) {
    const unsigned row = task / groups_per_row;
    const int w0 = (int)(task - row * groups_per_row) * kWordsPerTask;
    const T* __restrict__ src = vol + (size_t)row * g.X + (size_t)w0 * 32 + lane;
    T v[kWordsPerTask];
    if (FULL) {
#pragma unroll
      for (int k = 0; k < kWordsPerTask; ++k) v[k] = load_stream(src + k * 32);
    } else {
      const int xmax = g.X - 1 - (w0 * 32 + lane);  // largest valid offset from src (may be negative)
#pragma unroll
      for (int k = 0; k < kWordsPerTask; ++k) {
        const int off = k * 32 < xmax ? k * 32 : xmax;
        v[k] = load_stream(src + off);
      }
    }
    uint32_t word[kWordsPerTask];
#pragma unroll
    for (int k = 0; k < kWordsPerTask; ++k) word[k] = __ballot_sync(0xffffffffu, !(v[k] < iso));
    if (lane == 0) {
      uint32_t* dst = bits + (size_t)row * g.Wp + w0;
      if (FULL) {
#pragma unroll
        for (int k = 0; k < kWordsPerTask; k += 4)
          *reinterpret_cast<uint4*>(dst + k) = make_uint4(word[k], word[k + 1], word[k + 2], word[k + 3]);
      } else {
#pragma unroll
        for (int k = 0; k < kWordsPerTask; ++k)
          if (w0 + k < g.Wx) dst[k] = word[k];
      }
    }
  }
}

}  // namespace cub
