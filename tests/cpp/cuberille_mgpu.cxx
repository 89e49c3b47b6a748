// cuberille_mgpu.cxx — the multi-GPU host in C++ (north_star: "host code stays C++ and calls CUDA through a thin
// C-ABI layer"): one thread per GPU, z-slabs of ONE image, nothing but include/cuberille_c.h.
//
//   cuberille_mgpu <n_gpus> <size> [period] [triangles] [project]
//
// Every rank generates its slab of a size^3 gyroid (own range + halo) on its GPU, runs the hot path without a host
// round trip (cub_count_async -> cub_comm_exchange_counts -> cub_emit_async), and the meshes are gathered with
// cub_comm_gather_mesh.  Rank 0 then runs the WHOLE image through a second handle and compares the gathered mesh
// with it byte for byte: the concatenation of the slabs must be the single-GPU mesh (which tests/test_gpu_parity.py
// pins to the oracle).  Prints timing of the step (max over ranks) and "mgpu ok" / exits 1.
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "cuberille_c.h"

namespace {

struct Barrier {
  std::mutex m; std::condition_variable cv; int n, count = 0, phase = 0;
  explicit Barrier(int n_) : n(n_) {}
  void wait() {
    std::unique_lock<std::mutex> l(m);
    const int p = phase;
    if (++count == n) { count = 0; ++phase; cv.notify_all(); }
    else cv.wait(l, [&] { return phase != p; });
  }
};

#define CK(h, call)                                                                     \
  do {                                                                                  \
    int rc__ = (call);                                                                  \
    if (rc__ != CUB_OK) {                                                               \
      std::fprintf(stderr, "rank %d: %s -> %d: %s\n", rank, #call, rc__, cub_last_error(h)); \
      std::exit(1);                                                                     \
    }                                                                                   \
  } while (0)

struct Shared {
  unsigned char id[128];
  std::vector<double> step_ms;
  std::vector<uint64_t> counts;
  bool ok = true;
};

void run_rank(int rank, int world, uint64_t S, double period, int tri, int proj, Barrier* bar, Shared* sh) {
  cub_handle h = nullptr;
  if (cub_create(rank, nullptr, &h) != CUB_OK) { std::fprintf(stderr, "rank %d: no usable CUDA device\n", rank); std::exit(1); }
  cub_params p;
  cub_default_params(&p);
  p.iso_value = 0.0; p.generate_triangles = tri; p.project_vertices = proj; p.surface_distance_threshold = 0.01;
  uint64_t below = 2, above = 1;
  CK(h, cub_projection_halo(&p, nullptr, &below, &above));
  const uint64_t z0 = S * rank / world, z1 = S * (rank + 1) / world;
  const uint64_t lo = z0 > below ? z0 - below : 0, hi = z1 + above < S ? z1 + above : S;
  const uint64_t dims[3] = {S, S, hi - lo}, image[3] = {S, S, S};
  CK(h, cub_generate_volume(h, CUB_GEN_GYROID, dims, image, lo, period, 1.0, 1234));
  CK(h, cub_set_slab(h, S, lo, z0, z1));
  if (rank == 0 && cub_comm_unique_id(sh->id) != CUB_OK) { std::fprintf(stderr, "NCCL is not available\n"); std::exit(1); }
  bar->wait();
  cub_comm c = nullptr;
  CK(h, cub_comm_create(h, sh->id, world, rank, &c));
  uint64_t np = 0, nc = 0;
  for (int it = 0; it < 6; ++it) {  // the first pass sizes the buffers; the later ones are fully asynchronous
    bar->wait();
    const auto t0 = std::chrono::steady_clock::now();
    CK(h, cub_count_async(h, &p));
    CK(h, cub_comm_exchange_counts(c));
    CK(h, cub_emit_async(h, 4));
    CK(h, cub_finish(h, &np, &nc));
    sh->step_ms[rank] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  }
  std::vector<uint64_t> counts(2 * world);
  CK(h, cub_comm_counts(c, counts.data()));
  uint64_t tot_p = 0, tot_c = 0;
  for (int r = 0; r < world; ++r) { tot_p += counts[2 * r]; tot_c += counts[2 * r + 1] * (tri ? 2 : 1); }
  const int vpc = tri ? 3 : 4;
  void *d_pts = nullptr, *d_cells = nullptr;
  CK(h, cub_device_alloc(h, tot_p * 12, &d_pts));
  CK(h, cub_device_alloc(h, tot_c * vpc * 4, &d_cells));
  CK(h, cub_comm_gather_mesh(c, static_cast<float*>(d_pts), d_cells, nullptr));
  CK(h, cub_synchronize(h));
  bar->wait();
  if (rank == 0) {
    double ms = 0;
    for (double v : sh->step_ms) ms = v > ms ? v : ms;
    std::printf("%d GPUs, gyroid %llu^3: %llu points, %llu cells; step (host clock incl. the final sync, max over ranks) %.3f ms\n",
                world, (unsigned long long)S, (unsigned long long)tot_p, (unsigned long long)tot_c, ms);
    // the whole image on one GPU
    std::vector<float> gp(tot_p * 3);
    std::vector<uint32_t> gc(tot_c * vpc);
    CK(h, cub_device_copy(h, gp.data(), d_pts, tot_p * 12, CUB_MEM_HOST, CUB_MEM_DEVICE));
    CK(h, cub_device_copy(h, gc.data(), d_cells, tot_c * vpc * 4, CUB_MEM_HOST, CUB_MEM_DEVICE));
    cub_handle w = nullptr;
    if (cub_create(0, nullptr, &w) != CUB_OK) std::exit(1);
    CK(w, cub_generate_volume(w, CUB_GEN_GYROID, image, image, 0, period, 1.0, 1234));
    uint64_t wp = 0, wc = 0;
    CK(w, cub_run(w, &p, 4, &wp, &wc));
    std::vector<float> sp(wp * 3);
    std::vector<uint32_t> sc(wc * vpc);
    CK(w, cub_fetch(w, sp.data(), sc.data(), nullptr, CUB_MEM_HOST));
    const bool same = wp == tot_p && wc == tot_c && std::memcmp(sp.data(), gp.data(), sp.size() * 4) == 0 &&
                      std::memcmp(sc.data(), gc.data(), sc.size() * 4) == 0;
    std::printf("single GPU: %llu points, %llu cells -> gathered mesh %s\n", (unsigned long long)wp, (unsigned long long)wc,
                same ? "identical (bytes)" : "DIFFERS");
    sh->ok = same;
    cub_destroy(w);
  }
  CK(h, cub_device_free(h, d_pts));
  CK(h, cub_device_free(h, d_cells));
  cub_comm_destroy(c);
  cub_destroy(h);
}

}  // namespace

int main(int argc, char** argv) {
  if (argc < 3) { std::printf("USAGE: %s n_gpus size [period] [triangles] [project]\n", argv[0]); return 2; }
  const int world = std::atoi(argv[1]);
  const uint64_t S = std::strtoull(argv[2], nullptr, 10);
  const double period = argc > 3 ? std::atof(argv[3]) : 32.0;
  const int tri = argc > 4 ? std::atoi(argv[4]) : 0, proj = argc > 5 ? std::atoi(argv[5]) : 0;
  if (world < 1 || S < (uint64_t)world) return 2;
  Barrier bar(world);
  Shared sh;
  sh.step_ms.assign(world, 0.0);
  std::vector<std::thread> th;
  for (int r = 0; r < world; ++r) th.emplace_back(run_rank, r, world, S, period, tri, proj, &bar, &sh);
  for (auto& t : th) t.join();
  if (!sh.ok) return 1;
  std::printf("mgpu ok\n");
  return 0;
}
