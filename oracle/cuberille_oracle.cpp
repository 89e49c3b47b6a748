// cuberille_oracle.cpp — CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
//
// A plain C++17 restatement, over flat buffers and without ITK, of the hot path
// of itk::CuberilleImageToMeshFilter::GenerateData()
//   reference: Source/itkCuberilleImageToMeshFilter.txx:59-498 ("txx") and
//              Source/itkCuberilleImageToMeshFilter.h:243-313 ("h", lookup map).
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs may load this library, and only as the checker / CPU baseline.
// The product (libcuberille_cuda.so) never links, loads or calls it.
//
// Parity pinning: the reference's own tests pin ONLY point and cell counts
// (Testing/CuberilleTest01.cxx:193-204; 19 rows of Testing/CMakeLists.txt:10-331).
// tests/test_oracle_kat.py checks this file against all 19 rows.  Connectivity,
// vertex order and positions are "parity unpinned" by the reference (it ships no
// golden mesh and ITK is absent here, so the filter itself cannot be run):
// for those this file follows the reference source line by line (mode LITERAL)
// and the ITK semantics written down in SURVEY.md Appendix A.
//
// The ITK arithmetic restated here (ITK is a third-party dependency that is not
// vendored: FIND_PACKAGE(ITK) CMakeLists.txt:7-8, version unpinned, era 3.18-3.20):
//   * ConstShapedNeighborhoodIterator + ZeroFluxNeumannBoundaryCondition: an
//     out-of-image neighbour reads the index-clamped pixel (txx:98-100,167).
//   * Image::TransformIndexToPhysicalPoint into Point<float,3> (txx:266).
//   * LinearInterpolateImageFunction / VectorLinearInterpolateImageFunction
//     ::Evaluate, ITK 3.x N-d form (txx:451,455).
//   * GradientImageFilter: fp32 central differences, edge replicate (txx:486-495).
//   * CovariantVector<float,3>::Normalize (txx:452).
//   * Point::SquaredEuclideanDistanceTo (txx:298).
//   * Oriented images (a non-identity direction matrix D; no test of the reference has one):
//     itk::Image::TransformIndexToPhysicalPoint with M = D*diag(spacing) into a Point<float> (every
//     accumulation rounds to float), TransformPhysicalPointToContinuousIndex with M^-1 (fp64, row sums from 0),
//     GradientImageFilter's UseImageDirection rotation of every gradient pixel
//     (TransformLocalVectorToPhysicalVector: fp64 row sums of D * fp32 gradient, rounded to fp32).  The
//     reference's half-spacing shift stays axis-aligned in PHYSICAL space (txx:268-270), as written.
//     M^-1 is the cofactor inverse here (ITK takes vnl's SVD inverse: equal for axis flips / permutations,
//     equal to rounding otherwise).
//
// Build: see oracle/Makefile (g++ -O2 -ffp-contract=off; no -ffast-math, no -march).

#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <map>
#include <utility>
#include <vector>

extern "C" {

enum { ORC_U8 = 0, ORC_I8 = 1, ORC_U16 = 2, ORC_I16 = 3, ORC_U32 = 4, ORC_I32 = 5, ORC_F32 = 6, ORC_F64 = 7 };

// mode of the vertex lookup
enum {
  ORC_LITERAL = 0,     // two std::map planes swapped on the first inside voxel of a new z (txx:128-131,155-161,186-191)
  ORC_CLOSED_FORM = 1  // owner = adjacent (voxel, local) pair with minimal raster*8+local key
};

struct orc_params {
  double iso_value;
  int32_t generate_triangles;
  int32_t project_vertices;
  int32_t save_pixel_as_cell_data;
  int32_t mode;
  double surface_distance_threshold;
  double step_length;  // <0: auto (txx:82-85)
  double step_relaxation;
  uint32_t max_steps;
  uint32_t image_border_faces;  // 0: the reference (clamped neighbours: no face on the image border);
                                // 1: a neighbour outside the image is outside the surface (closed mesh, txx:133 TODO)
  int64_t region_index[3];      // image index of the buffer's first voxel (ImageRegion::GetIndex; the tests of the
                                // reference only have 0): TransformIndexToPhysicalPoint sees index + region_index
  double direction[9];          // row-major direction cosines; all zero or the identity: a non-oriented image
  int32_t projection_method;    // 0: the default branch (txx:440-474); 1: USE_ADVANCED_PROJECTION (txx:340-397);
                                // 2: USE_LINESEARCH_PROJECTION (txx:398-438) - compile-time alternates of the reference (h:22-23)
  int32_t reserved;
};

struct orc_mesh {
  std::vector<float> points;     // xyz
  std::vector<uint64_t> cells;   // 3 or 4 ids per cell
  std::vector<uint8_t> celldata; // raw pixels of the input dtype
  int verts_per_cell = 4;
  int pixel_bytes = 1;
  double step_length_used = 0.0;
  // number of points / cells that exist when the raster loop enters slice z (size nz + 1): what a z-slab
  // decomposition must reproduce with its exclusive scan of per-slab counts
  std::vector<uint64_t> points_before_slice, cells_before_slice;
};

}  // extern "C"

namespace {

struct Geometry {
  int64_t nx, ny, nz;
  double spacing[3];
  double origin[3];
  int64_t i0[3];  // image index of buffer voxel (0, 0, 0)
  bool oriented;  // a non-identity direction matrix
  double dir[9], m[9], minv[9];  // direction D, M = D*diag(spacing), M^-1 (row-major)
};

template <typename T>
struct Volume {
  const T* data;
  Geometry g;
  // ZeroFluxNeumannBoundaryCondition: index clamp
  inline T at_clamped(int64_t x, int64_t y, int64_t z) const {
    x = x < 0 ? 0 : (x >= g.nx ? g.nx - 1 : x);
    y = y < 0 ? 0 : (y >= g.ny ? g.ny - 1 : y);
    z = z < 0 ? 0 : (z >= g.nz ? g.nz - 1 : z);
    return data[(z * g.ny + y) * g.nx + x];
  }
};

// GradientImageFilter<Image,float,float> at one voxel:
// per axis sum = 0; sum += (-c)*I[-1]; sum += 0*I[0]; sum += c*I[+1]; all fp32,
// c = (float)(0.5/spacing), neighbours edge-replicated.
template <typename T>
inline void gradient_at(const Volume<T>& v, int64_t x, int64_t y, int64_t z, float g[3]) {
  const int64_t d[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (int a = 0; a < 3; ++a) {
    const float c = (float)(0.5 * (1.0 / v.g.spacing[a]));
    const float lo = (float)v.at_clamped(x - d[a][0], y - d[a][1], z - d[a][2]);
    const float mid = (float)v.at_clamped(x, y, z);
    const float hi = (float)v.at_clamped(x + d[a][0], y + d[a][1], z + d[a][2]);
    float sum = 0.0f;
    sum += (-c) * lo;
    sum += 0.0f * mid;
    sum += c * hi;
    g[a] = sum;
  }
  if (v.g.oriented) {  // GradientImageFilter::m_UseImageDirection: TransformLocalVectorToPhysicalVector
    float r[3];
    for (int i = 0; i < 3; ++i) {
      double sum = 0.0;
      for (int j = 0; j < 3; ++j) sum += v.g.dir[3 * i + j] * (double)g[j];
      r[i] = (float)sum;
    }
    g[0] = r[0]; g[1] = r[1]; g[2] = r[2];
  }
}

// Continuous index of a physical point (fp64): (p - origin) * (1/spacing)
// (ImageBase::TransformPhysicalPointToContinuousIndex with a diagonal
//  physical-to-index matrix; identity direction only).
inline void cont_index(const Geometry& g, const double p[3], double ci[3]) {
  if (!g.oriented) {
    for (int a = 0; a < 3; ++a) ci[a] = (p[a] - g.origin[a]) * (1.0 / g.spacing[a]);
    return;
  }
  // cvector = point - origin; cvector = m_PhysicalPointToIndex * cvector  (itk::Matrix * Vector: row sums from 0)
  double c[3];
  for (int a = 0; a < 3; ++a) c[a] = p[a] - g.origin[a];
  for (int i = 0; i < 3; ++i) {
    double sum = 0.0;
    for (int j = 0; j < 3; ++j) sum += g.minv[3 * i + j] * c[j];
    ci[i] = sum;
  }
}

struct InterpSetup {
  int64_t base[3];
  double dist[3];
};

inline InterpSetup interp_setup(const Geometry& g, const double p[3]) {
  InterpSetup s;
  double ci[3];
  cont_index(g, p, ci);
  for (int a = 0; a < 3; ++a) {
    const double f = std::floor(ci[a]);
    s.base[a] = (int64_t)f - g.i0[a];  // buffer-relative (the continuous index is an IMAGE index)
    s.dist[a] = ci[a] - f;
  }
  return s;
}

// LinearInterpolateImageFunction::EvaluateAtContinuousIndex, ITK 3.x N-d form
// (SURVEY Appendix A.4): neighbours in counter order 0..7 (bit0->x, bit1->y,
// bit2->z), overlap = ((1*wx)*wy)*wz, zero-overlap neighbours skipped, stop once
// the accumulated overlap == 1.0.  Out-of-image neighbour indices are clamped
// (the 3.x code reads out of bounds there: documented divergence).
template <typename T>
inline double interp_scalar(const Volume<T>& v, const double p[3]) {
  const InterpSetup s = interp_setup(v.g, p);
  double value = 0.0, total = 0.0;
  for (unsigned counter = 0; counter < 8; ++counter) {
    double overlap = 1.0;
    unsigned upper = counter;
    int64_t n[3];
    for (int a = 0; a < 3; ++a) {
      if (upper & 1) {
        n[a] = s.base[a] + 1;
        overlap *= s.dist[a];
      } else {
        n[a] = s.base[a];
        overlap *= 1.0 - s.dist[a];
      }
      upper >>= 1;
    }
    if (overlap) {
      value += overlap * (double)v.at_clamped(n[0], n[1], n[2]);
      total += overlap;
    }
    if (total == 1.0) break;
  }
  return value;
}

// VectorLinearInterpolateImageFunction on the gradient image: same weights,
// per component, fp64 accumulation of the fp32 gradient pixels.
template <typename T>
inline void interp_gradient(const Volume<T>& v, const double p[3], double out[3]) {
  const InterpSetup s = interp_setup(v.g, p);
  out[0] = out[1] = out[2] = 0.0;
  double total = 0.0;
  for (unsigned counter = 0; counter < 8; ++counter) {
    double overlap = 1.0;
    unsigned upper = counter;
    int64_t n[3];
    for (int a = 0; a < 3; ++a) {
      if (upper & 1) {
        n[a] = s.base[a] + 1;
        overlap *= s.dist[a];
      } else {
        n[a] = s.base[a];
        overlap *= 1.0 - s.dist[a];
      }
      upper >>= 1;
    }
    if (overlap) {
      // the gradient image has the input's geometry; clamp like the scalar case
      int64_t cx = n[0] < 0 ? 0 : (n[0] >= v.g.nx ? v.g.nx - 1 : n[0]);
      int64_t cy = n[1] < 0 ? 0 : (n[1] >= v.g.ny ? v.g.ny - 1 : n[1]);
      int64_t cz = n[2] < 0 ? 0 : (n[2] >= v.g.nz ? v.g.nz - 1 : n[2]);
      float g[3];
      gradient_at(v, cx, cy, cz, g);
      for (int k = 0; k < 3; ++k) out[k] += overlap * (double)g[k];
      total += overlap;
    }
    if (total == 1.0) break;
  }
}

// ProjectVertexToIsoSurface, default branch (txx:440-474).
template <typename T>
inline void project_vertex(const Volume<T>& v, const orc_params& P, double step0, float vertex[3]) {
  bool done = false;
  double sign = 1.0;
  double step = step0;                       // txx:443
  unsigned numberOfSteps = 0;
  const double iso = (double)(T)P.iso_value; // m_IsoSurfaceValue is an InputPixelType
  while (!done) {
    // normal = m_GradientInterpolator->Evaluate(vertex); normal.Normalize();   txx:451-452
    const double p[3] = {(double)vertex[0], (double)vertex[1], (double)vertex[2]};
    double gd[3];
    interp_gradient(v, p, gd);
    float normal[3] = {(float)gd[0], (float)gd[1], (float)gd[2]};  // CovariantVector<float,3>
    double sq = 0.0;
    for (int k = 0; k < 3; ++k) {
      const double c = (double)normal[k];
      sq += c * c;
    }
    const double norm = std::sqrt(sq);
    if (norm == 0.0) {
      // The reference divides by zero here (NaN vertex).  Policy of the new
      // build: leave the vertex where it is.
      break;
    }
    for (int k = 0; k < 3; ++k) normal[k] = (float)((double)normal[k] / norm);

    const double value = interp_scalar(v, p);  // txx:455
    done |= std::fabs(value - iso) < P.surface_distance_threshold;  // txx:456
    if (done) break;

    sign = (value < iso) ? +1.0 : -1.0;  // txx:463
    for (int k = 0; k < 3; ++k) {
      // vertex[i] += ( normal[i] * sign * step );  float += double     txx:466
      vertex[k] = (float)((double)vertex[k] + ((double)normal[k] * sign) * step);
    }
    step *= P.step_relaxation;                      // txx:468
    done |= numberOfSteps++ > P.max_steps;          // txx:469
  }
}

// normal = m_GradientInterpolator->Evaluate(vertex); normal.Normalize();  (txx:351-352, 411-412, 451-452)
// false: zero gradient (the reference would divide by zero)
template <typename T>
inline bool unit_normal(const Volume<T>& v, const float vertex[3], float normal[3]) {
  const double p[3] = {(double)vertex[0], (double)vertex[1], (double)vertex[2]};
  double gd[3];
  interp_gradient(v, p, gd);
  for (int k = 0; k < 3; ++k) normal[k] = (float)gd[k];  // CovariantVector<float,3>
  double sq = 0.0;
  for (int k = 0; k < 3; ++k) {
    const double c = (double)normal[k];
    sq += c * c;
  }
  const double norm = std::sqrt(sq);
  if (norm == 0.0) return false;
  for (int k = 0; k < 3; ++k) normal[k] = (float)((double)normal[k] / norm);
  return true;
}

template <typename T>
inline double value_at(const Volume<T>& v, const float pt[3]) {
  const double p[3] = {(double)pt[0], (double)pt[1], (double)pt[2]};
  return interp_scalar(v, p);
}

// ProjectVertexToIsoSurface, USE_ADVANCED_PROJECTION branch (txx:340-397): step both ways along the normal, keep the
// side whose value is closer to the iso value, stop on the threshold, the step count or five changes of side.
template <typename T>
inline void project_vertex_advanced(const Volume<T>& v, const orc_params& P, double step0, float vertex[3]) {
  bool done = false;
  double step = step0;
  unsigned numberOfSteps = 0;
  const double iso = (double)(T)P.iso_value;
  unsigned swaps = 0;
  int previousi = -1;
  while (!done) {
    float normal[3];
    if (!unit_normal(v, vertex, normal)) break;  // (same policy as the default branch)
    float temp[2][3];
    for (int i = 0; i < 3; ++i) {  // txx:355-359: float = float + (float * double * double)
      temp[0][i] = (float)((double)vertex[i] + ((double)normal[i] * +1.0) * step);
      temp[1][i] = (float)((double)vertex[i] + ((double)normal[i] * -1.0) * step);
    }
    step *= P.step_relaxation;  // txx:360
    const double diff[2] = {std::fabs(value_at(v, temp[0]) - iso), std::fabs(value_at(v, temp[1]) - iso)};  // txx:363-366
    const int i = (diff[0] <= diff[1]) ? 0 : 1;  // txx:367
    if (previousi < 0) previousi = i;
    swaps += (unsigned)(previousi != i);  // txx:369 (previousi is never updated in the reference)
    for (int k = 0; k < 3; ++k) vertex[k] = temp[i][k];
    done |= diff[i] < P.surface_distance_threshold;  // txx:373
    if (done) break;
    done |= numberOfSteps++ > P.max_steps;  // txx:380
    if (done) break;
    done |= (swaps >= 5);  // txx:387
  }
}

// ProjectVertexToIsoSurface, USE_LINESEARCH_PROJECTION branch (txx:398-438): one normal, max_steps/2 - 1 samples on
// either side within one step length, keep the sample whose value is closest to the iso value.
template <typename T>
inline void project_vertex_linesearch(const Volume<T>& v, const orc_params& P, double step0, float vertex[3]) {
  const double iso = (double)(T)P.iso_value;
  float normal[3];
  if (!unit_normal(v, vertex, normal)) return;
  float best[3] = {vertex[0], vertex[1], vertex[2]};  // (bestVertex is uninitialised in the reference if no sample improves on 10000)
  double bestMetric = 10000;
  const unsigned half = P.max_steps / 2;  // txx:418, integer division
  for (int side = 0; side < 2; ++side) {
    const double sign = side == 0 ? -1.0 : 1.0;  // txx:415
    for (unsigned j = 1; j < half; ++j) {
      const double d = (double)j / ((double)P.max_steps / 2.0);  // txx:421
      float temp[3];
      for (int i = 0; i < 3; ++i) temp[i] = (float)((double)vertex[i] + (((double)normal[i] * sign) * step0) * d);  // txx:424
      const double metric = std::fabs(value_at(v, temp) - iso);  // txx:427-428
      if (metric < bestMetric) {
        bestMetric = metric;
        best[0] = temp[0]; best[1] = temp[1]; best[2] = temp[2];
      }
    }
  }
  vertex[0] = best[0]; vertex[1] = best[1]; vertex[2] = best[2];
}

template <typename T>
inline void project_any(const Volume<T>& v, const orc_params& P, double step0, float vertex[3]) {
  if (P.projection_method == 1) project_vertex_advanced(v, P, step0, vertex);
  else if (P.projection_method == 2) project_vertex_linesearch(v, P, step0, vertex);
  else project_vertex(v, P, step0, vertex);
}

// AddVertex without the projection (txx:265-270; SURVEY Appendix A.2, ITK 3.x form).
inline void corner_position(const Geometry& g, int64_t cx, int64_t cy, int64_t cz, float out[3]) {
  const int64_t idx[3] = {cx, cy, cz};
  for (int a = 0; a < 3; ++a) {
    float p;
    if (!g.oriented) {
      p = (float)(g.spacing[a] * (double)(idx[a] + g.i0[a]) + g.origin[a]);  // TransformIndexToPhysicalPoint -> Point<float>
    } else {
      // point[i] = m_Origin[i]; for j: point[i] += m_IndexToPhysicalPoint[i][j] * index[j];   (Point<float>)
      p = (float)g.origin[a];
      for (int j = 0; j < 3; ++j) p = (float)((double)p + g.m[3 * a + j] * (double)(idx[j] + g.i0[j]));
    }
    p = (float)((double)p - g.spacing[a] / 2.0);                     // vertex[a] -= spacing[a]/2.0
    out[a] = p;
  }
}

// txx:219-233
const int kFaceCorners[6][4] = {{0, 4, 7, 3}, {0, 1, 5, 4}, {1, 2, 6, 5}, {2, 3, 7, 6}, {0, 3, 2, 1}, {4, 5, 6, 7}};
// txx:236-254
const int kCornerOffset[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0}, {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}};
// txx:121-127
const int kFaceOffset[6][3] = {{-1, 0, 0}, {0, -1, 0}, {+1, 0, 0}, {0, +1, 0}, {0, 0, -1}, {0, 0, +1}};

template <typename T>
struct Runner {
  Volume<T> vol;
  orc_params P;
  orc_mesh* mesh;
  T iso;
  double step0;

  void mark_slice() {
    mesh->points_before_slice.push_back(mesh->points.size() / 3);
    mesh->cells_before_slice.push_back(mesh->cells.size() / (size_t)mesh->verts_per_cell);
  }

  void add_vertex(int64_t cx, int64_t cy, int64_t cz) {  // txx:257-276
    float v[3];
    corner_position(vol.g, cx, cy, cz, v);
    if (P.project_vertices) project_any(vol, P, step0, v);
    mesh->points.insert(mesh->points.end(), v, v + 3);
  }

  double sqdist(uint64_t a, uint64_t b) const {  // Point::SquaredEuclideanDistanceTo (Appendix A.5)
    double s = 0.0;
    for (int k = 0; k < 3; ++k) {
      const double d = (double)mesh->points[3 * a + k] - (double)mesh->points[3 * b + k];
      s += d * d;
    }
    return s;
  }

  void push_celldata(T pixel) {
    const uint8_t* b = reinterpret_cast<const uint8_t*>(&pixel);
    mesh->celldata.insert(mesh->celldata.end(), b, b + sizeof(T));
  }

  void add_quad_face(const uint64_t f[4], T pixel) {  // txx:279-332
    if (P.generate_triangles) {
      uint64_t t1[3], t2[3];
      if (sqdist(f[0], f[2]) >= sqdist(f[1], f[3])) {  // txx:298
        t1[0] = f[0]; t1[1] = f[1]; t1[2] = f[3];
        t2[0] = f[1]; t2[1] = f[2]; t2[2] = f[3];
      } else {
        t1[0] = f[0]; t1[1] = f[1]; t1[2] = f[2];
        t2[0] = f[0]; t2[1] = f[2]; t2[2] = f[3];
      }
      mesh->cells.insert(mesh->cells.end(), t1, t1 + 3);
      mesh->cells.insert(mesh->cells.end(), t2, t2 + 3);
      if (P.save_pixel_as_cell_data) { push_celldata(pixel); push_celldata(pixel); }
    } else {
      mesh->cells.insert(mesh->cells.end(), f, f + 4);
      if (P.save_pixel_as_cell_data) push_celldata(pixel);
    }
  }

  // faceHasQuad / vertexHasQuad of one inside voxel (txx:164-173)
  inline int classify(int64_t x, int64_t y, int64_t z, bool faceHasQuad[6], bool vertexHasQuad[8]) const {
    int numFaces = 0;
    for (int i = 0; i < 6; ++i) faceHasQuad[i] = false;
    for (int i = 0; i < 8; ++i) vertexHasQuad[i] = false;
    for (int i = 0; i < 6; ++i) {
      const int64_t nx = x + kFaceOffset[i][0], ny = y + kFaceOffset[i][1], nz = z + kFaceOffset[i][2];
      if (P.image_border_faces && (nx < 0 || ny < 0 || nz < 0 || nx >= vol.g.nx || ny >= vol.g.ny || nz >= vol.g.nz))
        faceHasQuad[i] = true;  // opt-in: what padding the image with one outside layer would give
      else
        faceHasQuad[i] = vol.at_clamped(nx, ny, nz) < iso;
      if (faceHasQuad[i]) {
        ++numFaces;
        for (int k = 0; k < 4; ++k) vertexHasQuad[kFaceCorners[i][k]] = true;  // SetVerticesFromFace
      }
    }
    return numFaces;
  }

  // ---- mode LITERAL: the loop of txx:136-206 with its two lookup planes -----
  void run_literal() {
    typedef std::map<std::pair<uint64_t, uint64_t>, uint64_t> Plane;  // key (y, x): VertexLookupNode order h:258-265
    Plane lookup[2];
    unsigned look0 = 1, look1 = 0;
    int64_t lastZ = -1;
    uint64_t nextVertexId = 0;
    bool faceHasQuad[6], vertexHasQuad[8];
    uint64_t v[8], f[4];
    const Geometry& g = vol.g;
    for (int64_t z = 0; z < g.nz; ++z) {
      mark_slice();
      for (int64_t y = 0; y < g.ny; ++y)
        for (int64_t x = 0; x < g.nx; ++x) {
          const T center = vol.data[(z * g.ny + y) * g.nx + x];
          if (center < iso) continue;  // txx:139-141
          if (z != lastZ) {            // txx:155-161
            unsigned t = look0; look0 = look1; look1 = t;
            lookup[look1].clear();
            lastZ = z;
          }
          const int numFaces = classify(x, y, z, faceHasQuad, vertexHasQuad);
          if (numFaces > 0) {
            for (int i = 0; i < 8; ++i) {
              if (!vertexHasQuad[i]) continue;
              const int64_t cx = x + kCornerOffset[i][0], cy = y + kCornerOffset[i][1], cz = z + kCornerOffset[i][2];
              Plane& plane = lookup[(i < 4) ? look0 : look1];  // txx:185
              const std::pair<uint64_t, uint64_t> key((uint64_t)cy, (uint64_t)cx);
              Plane::iterator it = plane.find(key);
              if (it != plane.end()) {
                v[i] = it->second;
              } else {
                v[i] = nextVertexId;
                add_vertex(cx, cy, cz);
                ++nextVertexId;
                plane.insert(Plane::value_type(key, v[i]));
              }
            }
            for (int i = 0; i < 6; ++i) {  // txx:197-202
              if (!faceHasQuad[i]) continue;
              for (int k = 0; k < 4; ++k) f[k] = v[kFaceCorners[i][k]];
              add_quad_face(f, center);
            }
          }
        }
    }
    mark_slice();
  }

  // ---- mode CLOSED_FORM: ownership by minimal raster*8+local key -----------
  // A dense corner->id table; vertices are created when their owner (the first
  // voxel in raster order that activates the corner) is visited, in local order.
  void run_closed_form() {
    const Geometry& g = vol.g;
    const int64_t cnx = g.nx + 1, cny = g.ny + 1;
    std::vector<int64_t> cornerId((size_t)(cnx * cny * (g.nz + 1)), -1);
    uint64_t nextVertexId = 0;
    bool faceHasQuad[6], vertexHasQuad[8];
    uint64_t f[4];
    for (int64_t z = 0; z < g.nz; ++z) {
      mark_slice();
      for (int64_t y = 0; y < g.ny; ++y)
        for (int64_t x = 0; x < g.nx; ++x) {
          const T center = vol.data[(z * g.ny + y) * g.nx + x];
          if (center < iso) continue;
          const int numFaces = classify(x, y, z, faceHasQuad, vertexHasQuad);
          if (numFaces == 0) continue;
          int64_t cid[8];
          for (int i = 0; i < 8; ++i) {
            if (!vertexHasQuad[i]) continue;
            const int64_t cx = x + kCornerOffset[i][0], cy = y + kCornerOffset[i][1], cz = z + kCornerOffset[i][2];
            int64_t& slot = cornerId[(size_t)((cz * cny + cy) * cnx + cx)];
            if (slot < 0) {
              slot = (int64_t)nextVertexId++;
              add_vertex(cx, cy, cz);
            }
            cid[i] = slot;
          }
          for (int i = 0; i < 6; ++i) {
            if (!faceHasQuad[i]) continue;
            for (int k = 0; k < 4; ++k) f[k] = (uint64_t)cid[kFaceCorners[i][k]];
            add_quad_face(f, center);
          }
        }
    }
    mark_slice();
  }
};

template <typename T>
orc_mesh* run_typed(const void* data, const Geometry& g, const orc_params& P) {
  orc_mesh* m = new orc_mesh;
  m->verts_per_cell = P.generate_triangles ? 3 : 4;
  m->pixel_bytes = (int)sizeof(T);
  Runner<T> r;
  r.vol.data = static_cast<const T*>(data);
  r.vol.g = g;
  r.P = P;
  r.mesh = m;
  r.iso = (T)P.iso_value;
  // txx:75-85: auto step length = max spacing * 0.25
  double maxSpacing = g.spacing[0];
  for (int a = 1; a < 3; ++a) maxSpacing = g.spacing[a] > maxSpacing ? g.spacing[a] : maxSpacing;
  r.step0 = P.step_length < 0.0 ? maxSpacing * 0.25 : P.step_length;
  m->step_length_used = r.step0;
  if (P.mode == ORC_CLOSED_FORM) r.run_closed_form(); else r.run_literal();
  return m;
}

template <typename T>
void classify_typed(const void* data, const Geometry& g, double isod, uint32_t* out, uint64_t wpr) {
  const T* d = static_cast<const T*>(data);
  const T iso = (T)isod;
  for (int64_t z = 0; z < g.nz; ++z)
    for (int64_t y = 0; y < g.ny; ++y) {
      uint32_t* row = out + (size_t)((z * g.ny + y) * (int64_t)wpr);
      for (uint64_t w = 0; w < wpr; ++w) row[w] = 0;
      for (int64_t x = 0; x < g.nx; ++x) {
        const bool inside = !(d[(z * g.ny + y) * g.nx + x] < iso);
        if (inside) row[x >> 5] |= 1u << (x & 31);
      }
    }
}

template <typename T>
void project_typed(const void* data, const Geometry& g, const orc_params& P, float* pts, uint64_t n) {
  Volume<T> v;
  v.data = static_cast<const T*>(data);
  v.g = g;
  double maxSpacing = g.spacing[0];
  for (int a = 1; a < 3; ++a) maxSpacing = g.spacing[a] > maxSpacing ? g.spacing[a] : maxSpacing;
  const double step0 = P.step_length < 0.0 ? maxSpacing * 0.25 : P.step_length;
  for (uint64_t i = 0; i < n; ++i) project_any(v, P, step0, pts + 3 * i);
}

template <typename T>
void sample_typed(const void* data, const Geometry& g, const double* pts, uint64_t n, double* val, double* grad) {
  Volume<T> v;
  v.data = static_cast<const T*>(data);
  v.g = g;
  for (uint64_t i = 0; i < n; ++i) {
    val[i] = interp_scalar(v, pts + 3 * i);
    interp_gradient(v, pts + 3 * i, grad + 3 * i);
  }
}

Geometry make_geometry(const uint64_t dims[3], const double spacing[3], const double origin[3], const int64_t* i0 = nullptr,
                       const double* direction = nullptr) {
  Geometry g;
  for (int a = 0; a < 3; ++a) g.i0[a] = i0 ? i0[a] : 0;
  g.nx = (int64_t)dims[0]; g.ny = (int64_t)dims[1]; g.nz = (int64_t)dims[2];
  for (int a = 0; a < 3; ++a) { g.spacing[a] = spacing ? spacing[a] : 1.0; g.origin[a] = origin ? origin[a] : 0.0; }
  static const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  bool zero = true, ident = true;
  for (int k = 0; k < 9; ++k) {
    const double d = direction ? direction[k] : 0.0;
    zero = zero && d == 0.0;
    ident = ident && d == I[k];
  }
  g.oriented = direction && !zero && !ident;
  for (int k = 0; k < 9; ++k) {
    g.dir[k] = g.oriented ? direction[k] : I[k];
    g.m[k] = g.dir[k] * g.spacing[k % 3];   // D * diag(spacing)
  }
  // cofactor inverse of M
  const double* m = g.m;
  const double c00 = m[4] * m[8] - m[5] * m[7], c01 = m[5] * m[6] - m[3] * m[8], c02 = m[3] * m[7] - m[4] * m[6];
  const double det = m[0] * c00 + m[1] * c01 + m[2] * c02;
  g.minv[0] = c00 / det; g.minv[1] = (m[2] * m[7] - m[1] * m[8]) / det; g.minv[2] = (m[1] * m[5] - m[2] * m[4]) / det;
  g.minv[3] = c01 / det; g.minv[4] = (m[0] * m[8] - m[2] * m[6]) / det; g.minv[5] = (m[2] * m[3] - m[0] * m[5]) / det;
  g.minv[6] = c02 / det; g.minv[7] = (m[1] * m[6] - m[0] * m[7]) / det; g.minv[8] = (m[0] * m[4] - m[1] * m[3]) / det;
  return g;
}

#define ORC_DISPATCH(dtype, CALL)                         \
  switch (dtype) {                                        \
    case ORC_U8:  { typedef uint8_t  T; CALL; } break;    \
    case ORC_I8:  { typedef int8_t   T; CALL; } break;    \
    case ORC_U16: { typedef uint16_t T; CALL; } break;    \
    case ORC_I16: { typedef int16_t  T; CALL; } break;    \
    case ORC_U32: { typedef uint32_t T; CALL; } break;    \
    case ORC_I32: { typedef int32_t  T; CALL; } break;    \
    case ORC_F32: { typedef float    T; CALL; } break;    \
    case ORC_F64: { typedef double   T; CALL; } break;    \
    default: break;                                       \
  }

}  // namespace

extern "C" {

orc_mesh* orc_cuberille(const void* data, int dtype, const uint64_t dims[3], const double spacing[3],
                        const double origin[3], const orc_params* P) {
  const Geometry g = make_geometry(dims, spacing, origin, P->region_index, P->direction);
  orc_mesh* m = nullptr;
  ORC_DISPATCH(dtype, m = run_typed<T>(data, g, *P));
  return m;
}

uint64_t orc_num_points(const orc_mesh* m) { return m->points.size() / 3; }
uint64_t orc_num_cells(const orc_mesh* m) { return m->cells.size() / (uint64_t)m->verts_per_cell; }
int orc_verts_per_cell(const orc_mesh* m) { return m->verts_per_cell; }
const float* orc_points(const orc_mesh* m) { return m->points.data(); }
const uint64_t* orc_cells(const orc_mesh* m) { return m->cells.data(); }
const void* orc_cell_data(const orc_mesh* m) { return m->celldata.data(); }
uint64_t orc_cell_data_bytes(const orc_mesh* m) { return m->celldata.size(); }
double orc_step_length_used(const orc_mesh* m) { return m->step_length_used; }
// points / cells that exist when the loop enters slice z, z = 0..nz (nz + 1 values each)
const uint64_t* orc_points_before_slice(const orc_mesh* m) { return m->points_before_slice.data(); }
const uint64_t* orc_cells_before_slice(const orc_mesh* m) { return m->cells_before_slice.data(); }
void orc_free(orc_mesh* m) { delete m; }

// inside bitmask, 1 bit per voxel, rows padded to words_per_row 32-bit words,
// bits beyond nx are zero (the CUDA library's padding bits are compared masked).
int orc_classify(const void* data, int dtype, const uint64_t dims[3], double iso, uint32_t* out, uint64_t words_per_row) {
  const Geometry g = make_geometry(dims, nullptr, nullptr);
  if (words_per_row * 32 < dims[0]) return 1;
  ORC_DISPATCH(dtype, classify_typed<T>(data, g, iso, out, words_per_row));
  return 0;
}

// ProjectVertexToIsoSurface on caller points, in place.
int orc_project_points(const void* data, int dtype, const uint64_t dims[3], const double spacing[3],
                       const double origin[3], const orc_params* P, float* pts, uint64_t n) {
  const Geometry g = make_geometry(dims, spacing, origin, P->region_index, P->direction);
  ORC_DISPATCH(dtype, project_typed<T>(data, g, *P, pts, n));
  return 0;
}

// interpolated value and gradient at fp64 physical points (unit-level checks)
int orc_sample(const void* data, int dtype, const uint64_t dims[3], const double spacing[3],
               const double origin[3], const double* pts, uint64_t n, double* val, double* grad) {
  const Geometry g = make_geometry(dims, spacing, origin);
  ORC_DISPATCH(dtype, sample_typed<T>(data, g, pts, n, val, grad));
  return 0;
}

}  // extern "C"
