// The opaque dataset handle behind bl_dataset* (host side).
#pragma once
#include <atomic>
#include <map>
#include <vector>

#include "engine.cuh"

namespace bl {
struct Plan {
  Geometry g;
  int occupancy;
  uint32_t rn_scratch_off;
  bool rn_global;  // occu_rn A_k scratch lives in global memory
  int nch;           // site-parallel engine: chains interleaved per pass over a warp-tile
  int chain_variant; // BL_CHAIN_VARIANT as read when the plan was made
  int chain_bt;      // threads per block of the lane = chain variant (occu: 128 or 256)
  int chain_kernel;  // 0: site-parallel engine; 1: occu lane=chain kernel (K1c); 2 / 3 / 4: occu_rn / occu_cop / occu_cs lane=chain
                     // kernels; 5: occu lane=chain kernel over signed records (K1d, occu_signed.cu); 6: occu_rn lane=chain
                     // kernel in probability space over sorted records (K2d, occu_rn2.cu)
};
}  // namespace bl

struct bl_dataset {
  bl_desc desc{};
  bl::Layout L{};
  int D = 0, n_extras = 0, DS = 0;
  int num_sms = 0;
  size_t smem_limit = 0;
  void* packed = nullptr;
  size_t packed_bytes = 0;
  void* packed_signed = nullptr;  // occu fp32 without extras: AoS signed site records for K1d (occu_signed.cu)
  size_t packed_signed_bytes = 0;
  void* packed_rn2 = nullptr;     // occu_rn fp32 without extras: visits sorted detections-first for K2d (occu_rn2.cu)
  size_t packed_rn2_bytes = 0;
  double cop_const = 0.0;
  int64_t n_masked = 0;
  // fp64 block partials [nsplit][C][NQ], per-chunk tickets, raw sums for the collective path
  double* partial = nullptr;
  size_t partial_cap = 0;
  unsigned int* counters = nullptr;
  size_t counters_cap = 0;
  double* sums = nullptr;
  size_t sums_cap = 0;
  void* rn_scratch = nullptr;  // occu_rn global A_k scratch (only when it does not fit in smem)
  size_t rn_scratch_cap = 0;
  std::map<int, bl::Plan> plans;
  // n_species > 1 (multi.cu): one likelihood-only child handle per species + the gather / combine workspace
  std::vector<bl_dataset*> species;
  void *ms_theta = nullptr, *ms_lp = nullptr, *ms_grad = nullptr;
  double* ms_lp64 = nullptr;
  int ms_cap = 0;
  bool re = false;            // occu with site / observation random effects (occu_re.cu): its own kernel and layout
  bool force_engine = false;  // BL_FLAG_STRICT_MATH (shapes K1d does not cover), random effects: the site-parallel engine
  bool strict_chain = false;  // BL_FLAG_STRICT_MATH on an occu shape K1d covers: its libm instantiations for C >= 32
  // bl_eval_host staging
  void *d_theta = nullptr, *d_out = nullptr, *h_theta = nullptr, *h_out = nullptr;
  int host_cap = 0;
  cudaStream_t own_stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  // site-sharded collective (comm.cu)
  void* comm = nullptr;
};

namespace bl {
int fail(int code, const char* fmt, ...);
int eval_device(bl_dataset* ds, const void* theta, int C, void* logp, void* grad, cudaStream_t st, int allreduce,
                double* logp64);
extern std::atomic<int64_t> g_launches;
}  // namespace bl
