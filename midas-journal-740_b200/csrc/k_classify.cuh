// k_classify.cuh — K1: one pass over the volume -> 1 bit / voxel inside mask.
//
// Reference: the suitability test `center < m_IsoSurfaceValue -> skip` (txx:139-141) and the
// neighbour test `GetPixel(offset) < m_IsoSurfaceValue` (txx:167) compare pixels with the iso
// value in the pixel type; both are the same predicate, so it is evaluated ONCE per voxel here
// and every later kernel works on the bitmask (1/32 of the bytes for float input).
//
// HBM-bound streaming read: algorithmic bytes = X*Y*Zl*sizeof(T) read + X*Y*Zl/8 written.
// Each warp owns a task of kWordsPerTask consecutive words of one row; every lane issues
// kWordsPerTask independent coalesced loads (lane b of load k reads voxel 32*(w0+k)+b, i.e. a
// full 128-byte line per warp instruction for 4-byte pixels) before the first ballot, so a
// resident SM keeps (warps * kWordsPerTask * 128 B) in flight.  A ballot of the predicate IS
// the output word - no shuffles.  Lanes past the end of the row re-read the row's last voxel,
// which produces the "replicate bit X-1" padding the later kernels rely on.
#pragma once
#include "cub_common.cuh"

namespace cub {

constexpr int kWordsPerTask = 8;

template <typename T>
__device__ __forceinline__ T load_stream(const T* p) {
  return __ldcs(p);  // read-once data: evict-first
}

template <typename T>
__global__ void __launch_bounds__(256) k_classify(const T* __restrict__ vol, uint32_t* __restrict__ bits, Grid g,
                                                  T iso, long long n_tasks, int groups_per_row) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long task = warp0; task < n_tasks; task += n_warps) {
    const long long row = task / groups_per_row;
    const int w0 = (int)(task - row * groups_per_row) * kWordsPerTask;
    const T* __restrict__ src = vol + (size_t)row * g.X;
    T v[kWordsPerTask];
#pragma unroll
    for (int k = 0; k < kWordsPerTask; ++k) {
      int x = (w0 + k) * 32 + lane;
      x = x < g.X ? x : g.X - 1;
      v[k] = load_stream(src + x);
    }
    uint32_t mine = 0;
#pragma unroll
    for (int k = 0; k < kWordsPerTask; ++k) {
      const uint32_t word = __ballot_sync(0xffffffffu, !(v[k] < iso));
      if (lane == k) mine = word;
    }
    if (lane < kWordsPerTask && w0 + lane < g.Wx) bits[(size_t)row * g.Wp + w0 + lane] = mine;
  }
}

}  // namespace cub
