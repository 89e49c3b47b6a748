// k_project.cuh — K4: ProjectVertexToIsoSurface, one thread per vertex.
//
// Reference: txx:440-474 (default branch) on top of ComputeGradientImage (txx:479-498).  The
// reference materialises a 12 B/voxel fp32 gradient image for the whole volume and interpolates
// it; here the central differences (GradientImageFilter semantics) are
// evaluated on the fly at the <=8 lattice nodes a vertex touches, with exactly the arithmetic of
// the oracle: fp32 differences, fp64 trilinear weights in ITK 3.x neighbour order with the
// zero-overlap skip and the `totalOverlap == 1` early exit (Appendix A.4), fp32 normal,
// fp64 norm / step / sign, fp32 vertex.  The library is compiled with -fmad=false so that no
// multiply-add is contracted; double division and sqrt are IEEE in CUDA.  The result is
// bit-identical to oracle/cuberille_oracle.cpp::project_vertex (tests/test_gpu_parity.py).
//
// Divergence from the reference, shared with the oracle: out-of-image neighbour indices are
// clamped (the reference reads out of bounds) and a zero gradient stops the vertex where it is
// (the reference divides by zero).
#pragma once
#include "cub_common.cuh"

namespace cub {

struct ProjArgs {
  const void* vol;
  Grid g;
  Geom geom;
  double iso;    // (double)(T)iso
  double thr;    // m_ProjectVertexSurfaceDistanceThreshold
  double step0;  // m_ProjectVertexStepLength (after the auto rule txx:82-85)
  double relax;  // m_ProjectVertexStepLengthRelaxationFactor
  unsigned max_steps;
  float* points;
  size_t n_points;
};

template <typename T>
struct VolView {
  const T* __restrict__ d;
  int X, Y, Zl, zg0, Zg;
  __device__ __forceinline__ T at(int x, int y, int zg) const {  // image-clamped (zg is a GLOBAL z index)
    x = x < 0 ? 0 : (x > X - 1 ? X - 1 : x);
    y = y < 0 ? 0 : (y > Y - 1 ? Y - 1 : y);
    zg = zg < 0 ? 0 : (zg > Zg - 1 ? Zg - 1 : zg);
    int z = zg - zg0;
    z = z < 0 ? 0 : (z > Zl - 1 ? Zl - 1 : z);  // memory safety only (halo contract of cub_set_slab)
    return __ldg(d + ((size_t)z * Y + y) * X + x);
  }
};

template <typename T> struct is_fp { static constexpr bool value = false; };
template <> struct is_fp<float> { static constexpr bool value = true; };
template <> struct is_fp<double> { static constexpr bool value = true; };

// GradientImageFilter at one node: sum = 0; sum += (-c)*I[-1]; sum += 0*I[0]; sum += c*I[+1]  (fp32)
template <typename T>
__device__ __forceinline__ void gradient_at(const VolView<T>& v, const float c[3], int x, int y, int z, float g[3]) {
  // the 0*I[0] term only matters for non-finite float pixels; integer pixels skip the load
  const float mid = is_fp<T>::value ? (float)v.at(x, y, z) : 0.0f;
  {
    float s = 0.0f;
    s += (-c[0]) * (float)v.at(x - 1, y, z);
    s += 0.0f * mid;
    s += c[0] * (float)v.at(x + 1, y, z);
    g[0] = s;
  }
  {
    float s = 0.0f;
    s += (-c[1]) * (float)v.at(x, y - 1, z);
    s += 0.0f * mid;
    s += c[1] * (float)v.at(x, y + 1, z);
    g[1] = s;
  }
  {
    float s = 0.0f;
    s += (-c[2]) * (float)v.at(x, y, z - 1);
    s += 0.0f * mid;
    s += c[2] * (float)v.at(x, y, z + 1);
    g[2] = s;
  }
}

__device__ __forceinline__ int clampi(long long v, int hi) { return v < 0 ? 0 : (v > hi ? hi : (int)v); }

template <typename T>
__global__ void __launch_bounds__(128) k_project(const ProjArgs a) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n_points) return;
  VolView<T> v{static_cast<const T*>(a.vol), a.g.X, a.g.Y, a.g.Zl, a.g.zg0, a.g.Zg};
  float gc[3];
  double inv_sp[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    inv_sp[k] = 1.0 / a.geom.spacing[k];
    gc[k] = (float)(0.5 * inv_sp[k]);
  }
  float vert[3] = {a.points[3 * i], a.points[3 * i + 1], a.points[3 * i + 2]};

  bool done = false;
  double step = a.step0;
  unsigned numberOfSteps = 0;
  while (!done) {
    // continuous index, base index and distances (shared by both interpolators)
    long long base[3];
    double dist[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double ci = ((double)vert[k] - a.geom.origin[k]) * inv_sp[k];
      const double f = floor(ci);
      base[k] = (long long)f;
      dist[k] = ci - f;
    }
    double gd[3] = {0.0, 0.0, 0.0};
    double value = 0.0, total = 0.0;
    bool open = true;
#pragma unroll
    for (int counter = 0; counter < 8; ++counter) {
      if (open) {
        double overlap = 1.0;
        overlap *= (counter & 1) ? dist[0] : 1.0 - dist[0];
        overlap *= (counter & 2) ? dist[1] : 1.0 - dist[1];
        overlap *= (counter & 4) ? dist[2] : 1.0 - dist[2];
        if (overlap != 0.0) {
          const long long nx = base[0] + ((counter & 1) ? 1 : 0);
          const long long ny = base[1] + ((counter & 2) ? 1 : 0);
          const long long nz = base[2] + ((counter & 4) ? 1 : 0);
          const int cx = clampi(nx, v.X - 1), cy = clampi(ny, v.Y - 1), cz = clampi(nz, v.Zg - 1);
          float g[3];
          gradient_at(v, gc, cx, cy, cz, g);
          gd[0] += overlap * (double)g[0];
          gd[1] += overlap * (double)g[1];
          gd[2] += overlap * (double)g[2];
          value += overlap * (double)v.at(cx, cy, cz);
          total += overlap;
        }
        if (total == 1.0) open = false;
      }
    }
    // normal = (CovariantVector<float,3>) gradient; normal.Normalize()        txx:451-452
    float normal[3] = {(float)gd[0], (float)gd[1], (float)gd[2]};
    double sq = 0.0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double c = (double)normal[k];
      sq += c * c;
    }
    const double norm = sqrt(sq);
    if (norm == 0.0) break;
#pragma unroll
    for (int k = 0; k < 3; ++k) normal[k] = (float)((double)normal[k] / norm);

    done |= fabs(value - a.iso) < a.thr;  // txx:456
    if (done) break;
    const double sign = (value < a.iso) ? +1.0 : -1.0;  // txx:463
#pragma unroll
    for (int k = 0; k < 3; ++k) vert[k] = (float)((double)vert[k] + ((double)normal[k] * sign) * step);  // txx:466
    step *= a.relax;                                  // txx:468
    done |= numberOfSteps++ > a.max_steps;            // txx:469
  }
  a.points[3 * i] = vert[0];
  a.points[3 * i + 1] = vert[1];
  a.points[3 * i + 2] = vert[2];
}

}  // namespace cub
