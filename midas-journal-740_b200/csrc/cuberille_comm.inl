// cuberille_comm.inl — multi-GPU part of the C-ABI (included by cuberille_capi.cu): z-slabs over NCCL.
//
// The reference has no counterpart (GenerateData is one raster loop on one core, txx:136-206); SURVEY section 8e:
// the image is cut into contiguous z-slabs, one handle per GPU.  The only exchange of the data path is an
// all-gather of two integers per rank - the (points, quads) the own ranges produce - whose exclusive prefix gives
// the global id bases (NCCL has no exscan).  It is queued on a side stream, device to device: the gathered counts
// never visit the host, the prefix is taken by a one-thread kernel into the handle's info block, and the face
// kernel (the only consumer of the id base) waits for it through an event.  Optionally the meshes are gathered
// with an "allgatherv": grouped ncclSend / ncclRecv with the true counts, every rank's part landing at its id base
// of the destination buffers (no padding, no staging copy).
//
// NCCL is bound at run time (dlopen of libnccl.so.2: in a process that already uses NCCL - torch - this resolves to
// the same library), so single-GPU users of libcuberille_cuda.so do not need it at all.
#include <dlfcn.h>

namespace {

// the handful of NCCL entry points used here, with the types of nccl.h (2.x ABI)
typedef void* nccl_comm_t;
typedef struct { char internal[128]; } nccl_unique_id;
enum { kNcclUint8 = 1, kNcclUint64 = 5, kNcclFloat32 = 7 };
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(nccl_unique_id*) = nullptr;
  int (*CommInitRank)(nccl_comm_t*, int, nccl_unique_id, int) = nullptr;
  int (*CommDestroy)(nccl_comm_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, nccl_comm_t, cudaStream_t) = nullptr;
  int (*Send)(const void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  int (*GetVersion)(int*) = nullptr;
  std::string why;
};

NcclApi* nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return &api;
  tried = true;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (api.lib) break;
  }
  if (!api.lib) { api.why = std::string("libnccl.so.2 not found: ") + (dlerror() ? dlerror() : ""); return &api; }
#define CBR_SYM(field, name)                                                     \
  api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.lib, name));       \
  if (!api.field) { api.why = std::string("NCCL symbol missing: ") + name; api.lib = nullptr; return &api; }
  CBR_SYM(GetUniqueId, "ncclGetUniqueId")
  CBR_SYM(CommInitRank, "ncclCommInitRank")
  CBR_SYM(CommDestroy, "ncclCommDestroy")
  CBR_SYM(AllGather, "ncclAllGather")
  CBR_SYM(Send, "ncclSend")
  CBR_SYM(Recv, "ncclRecv")
  CBR_SYM(GroupStart, "ncclGroupStart")
  CBR_SYM(GroupEnd, "ncclGroupEnd")
  CBR_SYM(GetErrorString, "ncclGetErrorString")
  CBR_SYM(GetVersion, "ncclGetVersion")
#undef CBR_SYM
  return &api;
}

}  // namespace

struct cub_comm_s {
  cub_handle h = nullptr;
  nccl_comm_t comm = nullptr;
  int world = 1, rank = 0;
  cudaStream_t side = nullptr;            // the exchange runs here, beside the handle's vertex stage
  cudaEvent_t ev_counted = nullptr, ev_bases = nullptr;
  unsigned long long* d_gathered = nullptr;   // [2 * world] (points, quads) of every rank
  unsigned long long* h_gathered = nullptr;   // pinned
  bool exchanged = false, host_valid = false;
};

namespace {

#define NCCL_TRY(h, api, call)                                                                                  \
  do {                                                                                                          \
    int r__ = (call);                                                                                           \
    if (r__ != 0) return fail(h, CUB_ERR_CUDA, "%s: %s", #call, (api)->GetErrorString ? (api)->GetErrorString(r__) : "NCCL error"); \
  } while (0)

int comm_host_counts(cub_comm c) {
  cub_handle h = c->h;
  if (!c->exchanged) return fail(h, CUB_ERR_INVALID, "no count exchange has been queued on this communicator");
  if (c->host_valid) return CUB_OK;
  CU_TRY(h, cudaSetDevice(h->device));
  CU_TRY(h, cudaMemcpyAsync(c->h_gathered, c->d_gathered, 2 * (size_t)c->world * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->side));
  CU_TRY(h, cudaStreamSynchronize(c->side));
  c->host_valid = true;
  return CUB_OK;
}

}  // namespace

extern "C" {

int cub_comm_unique_id(unsigned char id[128]) {
  if (!id) return CUB_ERR_INVALID;
  NcclApi* api = nccl_api();
  if (!api->lib) return CUB_ERR_UNSUPPORTED;
  nccl_unique_id u;
  if (api->GetUniqueId(&u) != 0) return CUB_ERR_CUDA;
  memcpy(id, u.internal, 128);
  return CUB_OK;
}

int cub_comm_create(cub_handle h, const unsigned char id[128], int world, int rank, cub_comm* out) {
  if (!h || !out) return CUB_ERR_INVALID;
  *out = nullptr;
  if (!id || world < 1 || rank < 0 || rank >= world) return fail(h, CUB_ERR_INVALID, "bad communicator arguments");
  NcclApi* api = nccl_api();
  if (!api->lib) return fail(h, CUB_ERR_UNSUPPORTED, "NCCL is not available: %s", api->why.c_str());
  CU_TRY(h, cudaSetDevice(h->device));
  cub_comm c = new (std::nothrow) cub_comm_s;
  if (!c) return CUB_ERR_NOMEM;
  c->h = h; c->world = world; c->rank = rank;
  nccl_unique_id u;
  memcpy(u.internal, id, 128);
  int rc = api->CommInitRank(&c->comm, world, u, rank);
  if (rc != 0) { delete c; return fail(h, CUB_ERR_CUDA, "ncclCommInitRank: %s", api->GetErrorString(rc)); }
  bool ok = cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking) == cudaSuccess &&
            cudaEventCreateWithFlags(&c->ev_counted, cudaEventDisableTiming) == cudaSuccess &&
            cudaEventCreateWithFlags(&c->ev_bases, cudaEventDisableTiming) == cudaSuccess &&
            cudaMalloc(&c->d_gathered, 2 * (size_t)world * sizeof(unsigned long long)) == cudaSuccess &&
            cudaMallocHost(&c->h_gathered, 2 * (size_t)world * sizeof(unsigned long long)) == cudaSuccess;
  if (!ok) { cub_comm_destroy(c); return fail(h, CUB_ERR_CUDA, "communicator resources"); }
  *out = c;
  return CUB_OK;
}

int cub_comm_destroy(cub_comm c) {
  if (!c) return CUB_OK;
  cudaSetDevice(c->h->device);
  if (c->side) cudaStreamSynchronize(c->side);
  if (c->h->wait_before_faces == c->ev_bases) c->h->wait_before_faces = nullptr;
  NcclApi* api = nccl_api();
  if (c->comm && api->lib) api->CommDestroy(c->comm);
  if (c->d_gathered) cudaFree(c->d_gathered);
  if (c->h_gathered) cudaFreeHost(c->h_gathered);
  if (c->ev_counted) cudaEventDestroy(c->ev_counted);
  if (c->ev_bases) cudaEventDestroy(c->ev_bases);
  if (c->side) cudaStreamDestroy(c->side);
  delete c;
  return CUB_OK;
}

int cub_comm_exchange_counts(cub_comm c) {
  if (!c) return CUB_ERR_INVALID;
  cub_handle h = c->h;
  if (!h->count_queued) return fail(h, CUB_ERR_INVALID, "cub_comm_exchange_counts before cub_count / cub_count_async");
  NcclApi* api = nccl_api();
  CU_TRY(h, cudaSetDevice(h->device));
  // side stream: after the count kernels, beside whatever the caller queues next on the handle's stream
  CU_TRY(h, cudaEventRecord(c->ev_counted, h->stream));
  CU_TRY(h, cudaStreamWaitEvent(c->side, c->ev_counted, 0));
  NCCL_TRY(h, api, api->AllGather(h->d_info + kInfoPoints, c->d_gathered, 2, kNcclUint64, c->comm, c->side));
  k_bases_from_gathered<<<1, 32, 0, c->side>>>(h->d_info, c->d_gathered, c->rank, h->params.generate_triangles ? 2 : 1);
  h->launches++;
  CU_TRY(h, cudaGetLastError());
  CU_TRY(h, cudaEventRecord(c->ev_bases, c->side));
  h->wait_before_faces = c->ev_bases;  // the face kernel (and any read of the bases) waits for the exchange
  c->exchanged = true;
  c->host_valid = false;
  return CUB_OK;
}

int cub_comm_counts(cub_comm c, uint64_t* counts) {
  if (!c || !counts) return CUB_ERR_INVALID;
  CUB_TRY(comm_host_counts(c));
  for (int i = 0; i < 2 * c->world; ++i) counts[i] = c->h_gathered[i];
  return CUB_OK;
}

int cub_comm_gather_mesh(cub_comm c, float* points, void* cells, void* cell_data) {
  if (!c) return CUB_ERR_INVALID;
  cub_handle h = c->h;
  if (!h->emitted) return fail(h, CUB_ERR_INVALID, "cub_comm_gather_mesh before cub_emit");
  NcclApi* api = nccl_api();
  CU_TRY(h, cudaSetDevice(h->device));
  CUB_TRY(verify_emit(h));
  CUB_TRY(comm_host_counts(c));
  if (cell_data && !h->params.save_pixel_as_cell_data) return fail(h, CUB_ERR_INVALID, "cell data was not requested");
  const size_t cell_bytes = (size_t)h->verts_per_cell * h->id_bytes;
  const size_t cpq = h->params.generate_triangles ? 2 : 1;
  // every rank's part goes to its id base in the destination buffers: an all-gather with the true counts
  // (NCCL has no v-variant: one grouped send / receive per pair, straight from / into the final place)
  std::vector<size_t> pb(c->world + 1, 0), cb(c->world + 1, 0);
  for (int r = 0; r < c->world; ++r) {
    pb[r + 1] = pb[r] + (size_t)c->h_gathered[2 * r];
    cb[r + 1] = cb[r] + (size_t)c->h_gathered[2 * r + 1] * cpq;
  }
  const float* my_points = h->points.p + 3 * (size_t)h->ghost_v;
  const size_t my_np = (size_t)h->n_points, my_nc = (size_t)h->n_cells;
  NCCL_TRY(h, api, api->GroupStart());
  for (int r = 0; r < c->world; ++r) {
    if (r == c->rank) continue;
    const size_t np = pb[r + 1] - pb[r], nc = cb[r + 1] - cb[r];
    if (points) {
      if (my_np) NCCL_TRY(h, api, api->Send(my_points, my_np * 3, kNcclFloat32, r, c->comm, h->stream));
      if (np) NCCL_TRY(h, api, api->Recv(points + 3 * pb[r], np * 3, kNcclFloat32, r, c->comm, h->stream));
    }
    if (cells) {
      if (my_nc) NCCL_TRY(h, api, api->Send(h->cells.p, my_nc * cell_bytes, kNcclUint8, r, c->comm, h->stream));
      if (nc) NCCL_TRY(h, api, api->Recv(static_cast<unsigned char*>(cells) + cb[r] * cell_bytes, nc * cell_bytes, kNcclUint8, r, c->comm, h->stream));
    }
    if (cell_data) {
      if (my_nc) NCCL_TRY(h, api, api->Send(h->celldata.p, my_nc * h->pix_bytes, kNcclUint8, r, c->comm, h->stream));
      if (nc) NCCL_TRY(h, api, api->Recv(static_cast<unsigned char*>(cell_data) + cb[r] * h->pix_bytes, nc * h->pix_bytes, kNcclUint8, r, c->comm, h->stream));
    }
  }
  NCCL_TRY(h, api, api->GroupEnd());
  // the own part: device to device
  if (points && my_np) CU_TRY(h, cudaMemcpyAsync(points + 3 * pb[c->rank], my_points, my_np * 12, cudaMemcpyDeviceToDevice, h->stream));
  if (cells && my_nc)
    CU_TRY(h, cudaMemcpyAsync(static_cast<unsigned char*>(cells) + cb[c->rank] * cell_bytes, h->cells.p, my_nc * cell_bytes, cudaMemcpyDeviceToDevice, h->stream));
  if (cell_data && my_nc)
    CU_TRY(h, cudaMemcpyAsync(static_cast<unsigned char*>(cell_data) + cb[c->rank] * h->pix_bytes, h->celldata.p, my_nc * h->pix_bytes, cudaMemcpyDeviceToDevice, h->stream));
  return CUB_OK;
}

}  // extern "C"
