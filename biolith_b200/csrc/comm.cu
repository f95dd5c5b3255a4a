// Site-sharded evaluation (BASELINE.json configs[4]).
//
// Sites are split contiguously over ranks; every rank holds all C chains' parameters.  Per
// evaluation each rank produces raw fp64 sums [C][1+D] over ITS sites (likelihood kernels with
// EvalParams::allreduce = 1), the sums are added across ranks, and a finalize kernel adds the priors
// once and writes logp / grad -- identical bits on every rank, so the ranks' NUTS state machines stay
// in lock-step without any further exchange.  Two exchange modes:
//   mode 1 (NCCL):  ncclAllReduce(sum, fp64, C*(1+D)) on the eval stream (NVLink 5 / NVSwitch).  NCCL
//                   is bound at run time (dlopen libnccl.so.2) so the .so has no link-time dependency
//                   and shares the copy torch already loaded when the host uses torch.distributed.
//   mode 2 (P2P):   one fused kernel over CUDA-IPC peer memory: each rank publishes its sums + an epoch
//                   flag in its own exchange buffer, then reads every peer's buffer with system-scope
//                   loads in rank order (deterministic), adds the priors and writes the outputs: no
//                   NCCL launch on the critical path (the message is ~22 KB: pure latency).
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <new>

#include "engine.cuh"
#include "handle.h"

namespace bl {

struct NcclApi {
  void* lib = nullptr;
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclAllReduce) AllReduce = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
};

static NcclApi* nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.lib) break;
    }
    if (api.lib) {
      api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.lib, "ncclGetUniqueId");
      api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.lib, "ncclCommInitRank");
      api.AllReduce = (decltype(api.AllReduce))dlsym(api.lib, "ncclAllReduce");
      api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.lib, "ncclCommDestroy");
      api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.lib, "ncclGetErrorString");
      if (!api.GetUniqueId || !api.CommInitRank || !api.AllReduce || !api.CommDestroy) api.lib = nullptr;
    }
  }
  return api.lib ? &api : nullptr;
}

constexpr int kMaxPeers = 16;

struct Comm {
  int mode = 0, rank = 0, world = 1;
  ncclComm_t nccl = nullptr;
  // P2P exchange: [2 epochs parity][cap] doubles + flags, local buffer first then peers' mappings
  double* xbuf[kMaxPeers] = {};          // xbuf[r] = rank r's exchange buffer as mapped in this process
  unsigned int* xflag[kMaxPeers] = {};   // xflag[r][chunk] = last epoch published by rank r
  size_t cap = 0;                        // doubles per parity buffer
  unsigned int epoch = 0;
  int* d_error = nullptr;                // set by the kernel when a peer did not show up in time
};

constexpr size_t kFlagSlots = 64;

struct P2PParams {
  const double* xbuf[kMaxPeers];
  const unsigned int* xflag[kMaxPeers];
  double* my_buf;
  unsigned int* my_flag;
  int rank, world;
  unsigned int epoch;
  size_t cap;
  int* error;
};

// Fused exchange + finalize: grid = ceil(C*NQ / 256) blocks; every block first waits for all peers.
template <typename T>
__global__ void p2p_exchange_finalize_kernel(EvalParams p, P2PParams x) {
  const int n = p.C * p.NQ;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const size_t off = (size_t)(x.epoch & 1u) * x.cap;
  // 1. publish my sums (they were produced by the likelihood kernel earlier on this stream)
  if (idx < n) x.my_buf[off + idx] = p.sums[idx];
  __threadfence_system();
  __syncthreads();
  __shared__ int s_ok;
  if (threadIdx.x == 0) {
    // one arrival counter per rank: blocks of this rank count up to gridDim.x for this epoch
    const unsigned int done = atomicAdd(&x.my_flag[1], 1u) + 1u;
    if (done == gridDim.x) {
      x.my_flag[1] = 0;
      __threadfence_system();
      *(volatile unsigned int*)&x.my_flag[0] = x.epoch;  // all of my sums for this epoch are visible
    }
    // 2. wait until every peer has published this epoch (bounded spin: never hang the GPU)
    int ok = 1;
    const long long t0 = clock64();
    for (int r = 0; r < x.world && ok; ++r) {
      if (r == x.rank) continue;
      const volatile unsigned int* f = x.xflag[r];
      while ((int)(*f - x.epoch) < 0) {
        if (clock64() - t0 > 8000000000LL) { ok = 0; break; }  // ~4 s at 2 GHz
        __nanosleep(200);
      }
    }
    if (!ok) atomicExch(x.error, 1);
    s_ok = ok;
  }
  __syncthreads();
  __threadfence_system();
  if (!s_ok || idx >= n) return;
  // 3. deterministic rank-order sum, then priors
  double total = 0.0;
  for (int r = 0; r < x.world; ++r) {
    const double* src = (r == x.rank) ? (const double*)x.my_buf : x.xbuf[r];
    total += __ldcv(src + off + idx);
  }
  finalize_chain<T>(p, idx / p.NQ, idx % p.NQ, total, false);
}

int comm_post_eval(bl_dataset* ds, EvalParams& p, cudaStream_t st) {
  Comm* cm = reinterpret_cast<Comm*>(ds->comm);
  const int n = p.C * p.NQ;
  const int threads = 256, blocks = (n + threads - 1) / threads;
  const bool f32 = ds->desc.dtype == BL_F32;
  if (cm->mode == 1) {
    NcclApi* api = nccl_api();
    ncclResult_t r = api->AllReduce(p.sums, p.sums, (size_t)n, ncclDouble, ncclSum, cm->nccl, st);
    if (r != ncclSuccess) return fail(BL_ERR_NCCL, "ncclAllReduce: %s", api->GetErrorString ? api->GetErrorString(r) : "?");
    if (f32) finalize_kernel<float><<<blocks, threads, 0, st>>>(p);
    else finalize_kernel<double><<<blocks, threads, 0, st>>>(p);
  } else {
    if ((size_t)n > cm->cap) return fail(BL_ERR_INVALID, "P2P exchange buffer too small (%d > %zu)", n, cm->cap);
    P2PParams x{};
    for (int r = 0; r < cm->world; ++r) { x.xbuf[r] = cm->xbuf[r]; x.xflag[r] = cm->xflag[r]; }
    x.my_buf = cm->xbuf[cm->rank];
    x.my_flag = cm->xflag[cm->rank];
    x.rank = cm->rank; x.world = cm->world;
    x.epoch = ++cm->epoch;
    x.cap = cm->cap;
    x.error = cm->d_error;
    if (f32) p2p_exchange_finalize_kernel<float><<<blocks, threads, 0, st>>>(p, x);
    else p2p_exchange_finalize_kernel<double><<<blocks, threads, 0, st>>>(p, x);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(BL_ERR_CUDA, "exchange/finalize launch: %s", cudaGetErrorString(e));
  g_launches.fetch_add(cm->mode == 1 ? 2 : 1);
  return BL_OK;
}

void comm_destroy(bl_dataset* ds) {
  Comm* cm = reinterpret_cast<Comm*>(ds->comm);
  if (!cm) return;
  if (cm->nccl) { NcclApi* api = nccl_api(); if (api) api->CommDestroy(cm->nccl); }
  for (int r = 0; r < cm->world; ++r) {
    if (cm->mode != 2 || !cm->xbuf[r]) continue;
    if (r == cm->rank) cudaFree(cm->xbuf[r]);
    else cudaIpcCloseMemHandle(cm->xbuf[r]);
  }
  cudaFree(cm->d_error);
  delete cm;
  ds->comm = nullptr;
}

}  // namespace bl

using namespace bl;

#define CU_TRY(expr)                                                                              \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess) return fail(BL_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(_e));     \
  } while (0)

extern "C" {

int bl_comm_unique_id(void* id_out, size_t bytes) {
  if (!id_out || bytes < sizeof(ncclUniqueId)) return fail(BL_ERR_INVALID, "need %zu bytes", sizeof(ncclUniqueId));
  NcclApi* api = nccl_api();
  if (!api) return fail(BL_ERR_NCCL, "libnccl.so.2 could not be loaded");
  ncclUniqueId id;
  ncclResult_t r = api->GetUniqueId(&id);
  if (r != ncclSuccess) return fail(BL_ERR_NCCL, "ncclGetUniqueId failed (%d)", (int)r);
  memcpy(id_out, &id, sizeof(id));
  return BL_OK;
}

int bl_dataset_attach_nccl(bl_dataset* ds, const void* unique_id, size_t bytes, int32_t rank, int32_t world) {
  if (!ds || !unique_id || bytes < sizeof(ncclUniqueId)) return fail(BL_ERR_INVALID, "bad argument");
  if (world < 1 || rank < 0 || rank >= world) return fail(BL_ERR_INVALID, "rank %d / world %d", rank, world);
  NcclApi* api = nccl_api();
  if (!api) return fail(BL_ERR_NCCL, "libnccl.so.2 could not be loaded");
  CU_TRY(cudaSetDevice(ds->desc.device));
  comm_destroy(ds);
  Comm* cm = new (std::nothrow) Comm();
  if (!cm) return fail(BL_ERR_NOMEM, "host allocation failed");
  cm->mode = 1; cm->rank = rank; cm->world = world;
  ncclUniqueId id;
  memcpy(&id, unique_id, sizeof(id));
  ncclResult_t r = api->CommInitRank(&cm->nccl, world, id, rank);
  if (r != ncclSuccess) {
    delete cm;
    return fail(BL_ERR_NCCL, "ncclCommInitRank: %s", api->GetErrorString ? api->GetErrorString(r) : "?");
  }
  ds->comm = cm;
  return BL_OK;
}

/* P2P mode, step 1: allocate this rank's exchange buffer and export its CUDA-IPC handle (64 bytes). */
int bl_dataset_p2p_export(bl_dataset* ds, int32_t rank, int32_t world, int32_t max_chains, void* handle_out,
                          size_t bytes) {
  if (!ds || !handle_out || bytes < sizeof(cudaIpcMemHandle_t)) return fail(BL_ERR_INVALID, "bad argument");
  if (world < 1 || world > kMaxPeers || rank < 0 || rank >= world) return fail(BL_ERR_INVALID, "rank/world");
  CU_TRY(cudaSetDevice(ds->desc.device));
  comm_destroy(ds);
  Comm* cm = new (std::nothrow) Comm();
  if (!cm) return fail(BL_ERR_NOMEM, "host allocation failed");
  cm->mode = 2; cm->rank = rank; cm->world = world;
  cm->cap = (size_t)max_chains * (1 + ds->D);
  const size_t bytes_buf = 2 * cm->cap * sizeof(double) + kFlagSlots * sizeof(unsigned int);
  void* buf = nullptr;
  cudaError_t e = cudaMalloc(&buf, bytes_buf);
  if (e == cudaSuccess) e = cudaMemset(buf, 0, bytes_buf);
  if (e == cudaSuccess) e = cudaMalloc(&cm->d_error, sizeof(int));
  if (e == cudaSuccess) e = cudaMemset(cm->d_error, 0, sizeof(int));
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, buf);
  if (e != cudaSuccess) { cudaFree(buf); delete cm; return fail(BL_ERR_CUDA, "p2p export: %s", cudaGetErrorString(e)); }
  cm->xbuf[rank] = (double*)buf;
  cm->xflag[rank] = (unsigned int*)((char*)buf + 2 * cm->cap * sizeof(double));
  memcpy(handle_out, &h, sizeof(h));
  ds->comm = cm;
  return BL_OK;
}

/* P2P mode, step 2: handles = world x 64 bytes gathered from all ranks (own entry ignored). */
int bl_dataset_p2p_attach(bl_dataset* ds, const void* handles, size_t bytes_each) {
  if (!ds || !ds->comm || !handles || bytes_each < sizeof(cudaIpcMemHandle_t)) return fail(BL_ERR_INVALID, "bad argument");
  Comm* cm = reinterpret_cast<Comm*>(ds->comm);
  if (cm->mode != 2) return fail(BL_ERR_INVALID, "call bl_dataset_p2p_export first");
  CU_TRY(cudaSetDevice(ds->desc.device));
  for (int r = 0; r < cm->world; ++r) {
    if (r == cm->rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char*)handles + (size_t)r * bytes_each, sizeof(h));
    void* ptr = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return fail(BL_ERR_CUDA, "cudaIpcOpenMemHandle(rank %d): %s", r, cudaGetErrorString(e));
    cm->xbuf[r] = (double*)ptr;
    cm->xflag[r] = (unsigned int*)((char*)ptr + 2 * cm->cap * sizeof(double));
  }
  return BL_OK;
}

int bl_dataset_comm_error(bl_dataset* ds, int32_t* error) {
  if (!ds || !error) return fail(BL_ERR_INVALID, "bad argument");
  *error = 0;
  Comm* cm = reinterpret_cast<Comm*>(ds->comm);
  if (cm && cm->d_error) CU_TRY(cudaMemcpy(error, cm->d_error, sizeof(int), cudaMemcpyDeviceToHost));
  return BL_OK;
}

int bl_dataset_detach_comm(bl_dataset* ds) {
  if (!ds) return BL_OK;
  cudaSetDevice(ds->desc.device);
  comm_destroy(ds);
  return BL_OK;
}

}  // extern "C"
