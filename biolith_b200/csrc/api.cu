// C ABI of libbiolith_b200.so (declared in include/biolith_b200.h).  Plain pointers and sizes only.
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <new>
#include <vector>

#include "engine.cuh"
#include "handle.h"

namespace bl {

thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

// kernels defined in the per-model translation units
cudaError_t launch_pack(int data_dtype, int dtype, const void* y, const void* X, const void* W, const void* T,
                        void* out, const Layout& L, int model, int K, int* err_flag, unsigned long long* n_masked,
                        cudaStream_t st);
cudaError_t launch_export_mask(int dtype, const void* packed, uint8_t* mask, const Layout& L, cudaStream_t st);
cudaError_t launch_occu(const EvalParams& p, int dtype, dim3 grid, size_t smem, cudaStream_t stream, int* occ);
cudaError_t launch_occu_rn(const EvalParams& p, int dtype, dim3 grid, size_t smem, cudaStream_t stream, int* occ);
cudaError_t launch_occu_cop(const EvalParams& p, int dtype, dim3 grid, size_t smem, cudaStream_t stream, int* occ);
cudaError_t launch_nmixture(const EvalParams& p, int dtype, dim3 grid, size_t smem, cudaStream_t stream, int* occ);
cudaError_t launch_nmixture_summary(const EvalParams& p, int dtype, float* out, cudaStream_t st);
cudaError_t launch_occu_cs(const EvalParams& p, int dtype, dim3 grid, size_t smem, cudaStream_t stream, int* occ);
cudaError_t launch_occu_cs_summary(const EvalParams& p, int dtype, float* out, cudaStream_t st);
int occu_cs_has_specialisation(int ks, int ko);
int occu_has_specialisation(int ks, int ko, bool fp);
cudaError_t launch_occu_summary(const EvalParams& p, int dtype, float* out, cudaStream_t st);
cudaError_t launch_occu_rn_summary(const EvalParams& p, int dtype, float* out, cudaStream_t st);
cudaError_t launch_occu_cop_summary(const EvalParams& p, int dtype, float* out, cudaStream_t st);
int occu_derived_slots(uint32_t flags);
int occu_rn_derived_slots(uint32_t flags);
int occu_cop_derived_slots(uint32_t flags);
size_t occu_rn_extra_smem(const Layout& L, int K, int elem);
bool occu_chain_supported(int dtype, int ks, int ko, uint32_t flags);
cudaError_t launch_occu_chain(const EvalParams& p, dim3 grid, size_t smem, cudaStream_t st, int* occ);
size_t occu_chain_smem(const Layout& L, int nstage, int block_threads);
int occu_chain_block_threads(int ks, int ko, int C);
int occu_chain_variant();
bool occu_signed_supported(int dtype, int ks, int ko, uint32_t flags);
size_t occu_signed_bytes(const Layout& L);
size_t occu_signed_smem(const Layout& L, int nstage, int block_threads);
int occu_signed_block_threads(int C);
int64_t occu_signed_block_tiles(const Layout& L);
cudaError_t launch_repack_signed(const void* packed, void* out, const Layout& L, cudaStream_t st);
cudaError_t launch_occu_signed(const EvalParams& p, dim3 grid, size_t smem, cudaStream_t st, int* occ);
bool occu_small_supported(int dtype, int ks, int ko, uint32_t flags);
int occu_small_max_chains();
int occu_small_block_threads(const Layout& L, int CB, size_t smem_budget);
size_t occu_small_smem(const Layout& L, int nstage, int block_threads);
cudaError_t launch_occu_small(const EvalParams& p, dim3 grid, size_t smem, cudaStream_t st, int* occ);
bool occu_rn2_supported(int dtype, int ks, int ko, int J, int K, uint32_t flags);
size_t occu_rn2_bytes(const Layout& L);
size_t occu_rn2_smem(const Layout& L, int nstage, int K, int bt);
int occu_rn2_block_threads(const Layout& L, int C, int K, size_t smem_limit);
cudaError_t launch_repack_rn2(const void* packed, void* out, const Layout& L, cudaStream_t st);
cudaError_t launch_occu_rn2(const EvalParams& p, dim3 grid, size_t smem, cudaStream_t st, int* occ);
bool occu_rn_chain_supported(int dtype, int ks, int ko, uint32_t flags);
int occu_rn_chain_block_threads(int C);
size_t occu_rn_chain_smem(const Layout& L, int nstage, int K, int D, int bt);
cudaError_t launch_occu_rn_chain(const EvalParams& p, dim3 grid, size_t smem, cudaStream_t st, int* occ);
bool occu_cop_chain_supported(int dtype, int ks, int ko, uint32_t flags);
int occu_cop_chain_block_threads(int C);
size_t occu_cop_chain_smem(const Layout& L, int nstage, int bt);
cudaError_t launch_occu_cop_chain(const EvalParams& p, dim3 grid, size_t smem, cudaStream_t st, int* occ);

bool occu_cs_chain_supported(int dtype, int ks, int ko, uint32_t flags);
int occu_cs_chain_block_threads(int C);
size_t occu_cs_chain_smem(const Layout& L, int nstage, int bt);
cudaError_t launch_occu_cs_chain(const EvalParams& p, dim3 grid, size_t smem, cudaStream_t st, int* occ);

int64_t occu_re_theta_dim(const Layout& L, int64_t S, uint32_t flags);
int occu_re_n_small(const Layout& L, int64_t S, uint32_t flags);
cudaError_t launch_occu_re(const EvalParams& p, int dtype, int64_t S, int grid_x, double sd_scale_site,
                           double sd_scale_obs, cudaStream_t st);
cudaError_t launch_obs_loglik(const EvalParams& p, int dtype, int n_draws, float* lppd, float* var, cudaStream_t st);
int eval_device_multi(bl_dataset* ds, const void* theta, int C, void* logp, void* grad, cudaStream_t st,
                      double* logp64, void (*fill)(const bl_dataset*, EvalParams&));
int comm_post_eval(bl_dataset* ds, EvalParams& p, cudaStream_t st);
void comm_destroy(bl_dataset* ds);

// smallest batch handed to the lane = chain kernels: their 128-thread blocks (idle warps skip the arithmetic)
// overtake the site-parallel engine at 32 chains for all three models (B200, ms per evaluation, chain /
// engine: occu 0.47 / 0.54, occu_rn 9.1 / 14.7, occu_cop 0.36 / 0.45); at 16 chains the engine wins.
constexpr int kChainKernelMinChains = 32;

static cudaError_t launch_model(const bl_dataset* ds, const EvalParams& p, dim3 grid, size_t smem, cudaStream_t st,
                                int* occ) {
  switch (ds->desc.model) {
    case BL_MODEL_OCCU: return launch_occu(p, ds->desc.dtype, grid, smem, st, occ);
    case BL_MODEL_OCCU_RN: return launch_occu_rn(p, ds->desc.dtype, grid, smem, st, occ);
    case BL_MODEL_NMIXTURE: return launch_nmixture(p, ds->desc.dtype, grid, smem, st, occ);
    case BL_MODEL_OCCU_CS: return launch_occu_cs(p, ds->desc.dtype, grid, smem, st, occ);
    default: return launch_occu_cop(p, ds->desc.dtype, grid, smem, st, occ);
  }
}

static size_t elem_size(int dtype) { return dtype == BL_F32 ? 4 : 8; }

static void fill_params(const bl_dataset* ds, EvalParams& p) {
  memset(&p, 0, sizeof(p));
  p.packed = ds->packed;
  p.L = ds->L;
  p.model = ds->desc.model;
  p.D = ds->D;
  p.DS = ds->DS;
  p.NQ = 1 + ds->D;
  p.flags = ds->desc.flags;
  p.K = ds->desc.max_abundance;
  p.cop_const = ds->cop_const;
  p.prior_beta_loc = ds->desc.prior_beta_loc;
  p.prior_beta_scale = ds->desc.prior_beta_scale;
  p.prior_alpha_loc = ds->desc.prior_alpha_loc;
  p.prior_alpha_scale = ds->desc.prior_alpha_scale;
  p.prior_beta_norm = log(p.prior_beta_scale) + 0.91893853320467274178;
  p.prior_alpha_norm = log(p.prior_alpha_scale) + 0.91893853320467274178;
  p.prior_beta_iscale = 1.0 / p.prior_beta_scale;
  p.prior_alpha_iscale = 1.0 / p.prior_alpha_scale;
  p.prior_fp_a = ds->desc.prior_fp_a;
  p.prior_fp_b = ds->desc.prior_fp_b;
  p.prior_fp_rate = ds->desc.prior_fp_rate;
  p.nch = 4;
  p.coop_reduce = 1;
  if (const char* ev = getenv("BL_COOP_REDUCE")) p.coop_reduce = atoi(ev) != 0;
  p.ring_mode = 0;
  if (const char* ev = getenv("BL_SIGNED_RING")) p.ring_mode = atoi(ev);
  p.hier_reduce = 1;
  if (const char* ev = getenv("BL_HIER_REDUCE")) p.hier_reduce = atoi(ev) != 0;
}

// geometry for C chains (cached); grows the fp64 partial workspace when needed
static int plan_for(bl_dataset* ds, int C, Plan** out) {
  auto it = ds->plans.find(C);
  if (it == ds->plans.end()) {
    Plan pl{};
    const int elem = (int)elem_size(ds->desc.dtype);
    size_t extra = 0;
    if (ds->desc.model == BL_MODEL_OCCU_RN || ds->desc.model == BL_MODEL_NMIXTURE)
      extra = occu_rn_extra_smem(ds->L, ds->desc.max_abundance, elem);  // per-thread (K+1) column
    pl.rn_global = false;
    if (extra && extra + 4096 < ds->smem_limit) {
      // the per-thread A_k column goes to shared memory when it fits beside the tile ring ...
      pl.g = plan_geometry(ds->L, elem, C, ds->D, ds->DS, ds->num_sms, 1, ds->smem_limit - extra);
      if (pl.g.smem_bytes + extra + 128 > ds->smem_limit) pl.rn_global = true;
    } else if (extra) {
      pl.rn_global = true;
    }
    if (pl.rn_global) extra = 0;  // ... else to (coalesced, L2-resident) global scratch
    if (!extra) {
      // resident blocks per SM the tile ring is sized for: small chain batches are latency-bound (one
      // tile per warp per iteration), so they trade ring depth for more resident warps
      int bps = 2;
      if (const char* ev = getenv("BL_ENGINE_BPS")) bps = atoi(ev) > 0 ? atoi(ev) : bps;
      int wc = 0;  // A/B switch: "old" = the previous rule (as many chain-groups as chains, up to 8)
      if (const char* ev = getenv("BL_ENGINE_WC"))
        wc = !strcmp(ev, "old") ? pow2_floor(C < kWarpsPerBlock ? (C < 1 ? 1 : C) : kWarpsPerBlock) : atoi(ev);
      pl.g = plan_geometry(ds->L, elem, C, ds->D, ds->DS, ds->num_sms, bps, ds->smem_limit, wc);
    }
    pl.nch = 4;  // chains a warp interleaves per pass (models that implement site_chain_n)
    if (const char* ev = getenv("BL_ENGINE_NCH")) pl.nch = atoi(ev) > 0 ? atoi(ev) : pl.nch;
    pl.rn_scratch_off = (uint32_t)((pl.g.smem_bytes + 127) & ~size_t(127));
    pl.g.smem_bytes = pl.rn_scratch_off + extra;
    int chain_min = kChainKernelMinChains;
    if (const char* ev = getenv("BL_CHAIN_MIN")) chain_min = atoi(ev) > 0 ? atoi(ev) : chain_min;
    const bool want_chain = C >= chain_min && !ds->force_engine;
    pl.chain_kernel = 0;
    if (want_chain && ds->desc.model == BL_MODEL_OCCU &&
        occu_chain_supported(ds->desc.dtype, ds->L.ks, ds->L.ko, ds->desc.flags))
      pl.chain_kernel = 1;
    if (pl.chain_kernel == 1 && ds->packed_signed) {  // K1d unless the tuning switch asks for K1c
      const char* ev = getenv("BL_OCCU_CHAIN_KERNEL");
      if (!ev || atoi(ev) != 1 || ds->strict_chain) pl.chain_kernel = 5;
    }
    if (pl.chain_kernel == 1 && ds->strict_chain) pl.chain_kernel = 0;  // K1c has no libm form
    if (want_chain && ds->desc.model == BL_MODEL_OCCU_RN &&
        occu_rn_chain_supported(ds->desc.dtype, ds->L.ks, ds->L.ko, ds->desc.flags) &&
        occu_rn_chain_smem(ds->L, 2, ds->desc.max_abundance, ds->D, occu_rn_chain_block_threads(C)) <= ds->smem_limit)
      pl.chain_kernel = 2;
    if (want_chain && ds->desc.model == BL_MODEL_OCCU_RN && ds->packed_rn2 &&
        occu_rn2_smem(ds->L, 2, ds->desc.max_abundance, 128) <= ds->smem_limit)
      pl.chain_kernel = 6;
    if (want_chain && ds->desc.model == BL_MODEL_OCCU_COP &&
        occu_cop_chain_supported(ds->desc.dtype, ds->L.ks, ds->L.ko, ds->desc.flags))
      pl.chain_kernel = 3;
    // occu_cs: the engine wins below 64 chains (B200, 500k x 10: 32 chains 0.71 vs 1.04 ms, 64: 1.41 vs 1.43)
    if (want_chain && C >= 64 && ds->desc.model == BL_MODEL_OCCU_CS &&
        occu_cs_chain_supported(ds->desc.dtype, ds->L.ks, ds->L.ko, ds->desc.flags))
      pl.chain_kernel = 4;
    // occu below the chain kernels' threshold: K1s (lane = site, register-resident sums of <= 8 chains per block)
    if (!want_chain && C < chain_min && !ds->force_engine && !extra && ds->desc.model == BL_MODEL_OCCU &&
        occu_small_supported(ds->desc.dtype, ds->L.ks, ds->L.ko, ds->desc.flags)) {
      pl.chain_kernel = 7;
      const int ncmax = occu_small_max_chains();
      pl.chain_variant = 0;
      pl.g.n_chunks = (C + ncmax - 1) / ncmax;
      pl.g.CB = (C + pl.g.n_chunks - 1) / pl.g.n_chunks;
      const size_t budget = ds->smem_limit - 24 * 1024;  // static shared memory + reserve
      pl.chain_bt = occu_small_block_threads(ds->L, pl.g.CB, budget);  // one block per SM, one TMA ring per warp
      if (pl.chain_bt > 0) {
        pl.g.WS = pl.chain_bt / kWarp; pl.g.WC = 1;
        pl.g.n_block_tiles = (ds->L.n_tiles + pl.g.WS - 1) / pl.g.WS;
        pl.g.nstage = kMaxStages;
        while (pl.g.nstage > 2 && occu_small_smem(ds->L, pl.g.nstage, pl.chain_bt) > budget) --pl.g.nstage;
        pl.g.smem_bytes = occu_small_smem(ds->L, pl.g.nstage, pl.chain_bt);
      } else {  // very wide units: the engine
        pl.chain_kernel = 0;
        pl.g = plan_geometry(ds->L, elem, C, ds->D, ds->DS, ds->num_sms, 2, ds->smem_limit, 0);
      }
    } else if (pl.chain_kernel) {  // lane = chain: one warp-tile per stage, no theta / accumulator staging
      const int bt = pl.chain_kernel == 1   ? occu_chain_block_threads(ds->L.ks, ds->L.ko, C)
                     : pl.chain_kernel == 5 ? occu_signed_block_threads(C)
                     : pl.chain_kernel == 6 ? occu_rn2_block_threads(ds->L, C, ds->desc.max_abundance, ds->smem_limit)
                     : pl.chain_kernel == 2 ? occu_rn_chain_block_threads(C)
                     : pl.chain_kernel == 3 ? occu_cop_chain_block_threads(C)
                                            : occu_cs_chain_block_threads(C);  // chains per block
      pl.chain_bt = bt;
      pl.chain_variant = occu_chain_variant();
      pl.g.n_chunks = (C + bt - 1) / bt;
      pl.g.CB = (C + pl.g.n_chunks - 1) / pl.g.n_chunks;
      pl.g.WS = 1; pl.g.WC = kWarpsPerBlock;
      pl.g.n_block_tiles = ds->L.n_tiles;
      pl.g.nstage = kMaxStages;
      if (pl.chain_kernel == 1) {
        pl.g.smem_bytes = occu_chain_smem(ds->L, pl.g.nstage, bt);
      } else if (pl.chain_kernel == 5) {
        pl.g.n_block_tiles = occu_signed_block_tiles(ds->L);  // groups of warp-tiles per ring slot
        while (pl.g.nstage > 2 && occu_signed_smem(ds->L, pl.g.nstage, bt) > ds->smem_limit * bt / 512) --pl.g.nstage;
        pl.g.smem_bytes = occu_signed_smem(ds->L, pl.g.nstage, bt);
      } else if (pl.chain_kernel == 6) {
        const size_t budget = bt == 256 ? ds->smem_limit / 2 : ds->smem_limit / 3;
        while (pl.g.nstage > 2 && occu_rn2_smem(ds->L, pl.g.nstage, ds->desc.max_abundance, bt) > budget) --pl.g.nstage;
        pl.g.smem_bytes = occu_rn2_smem(ds->L, pl.g.nstage, ds->desc.max_abundance, bt);
        pl.rn_global = false;
      } else if (pl.chain_kernel == 3) {
        pl.g.smem_bytes = occu_cop_chain_smem(ds->L, pl.g.nstage, bt);
      } else if (pl.chain_kernel == 4) {
        pl.g.smem_bytes = occu_cs_chain_smem(ds->L, pl.g.nstage, bt);
      } else {
        while (pl.g.nstage > 2 &&
               occu_rn_chain_smem(ds->L, pl.g.nstage, ds->desc.max_abundance, ds->D, bt) > ds->smem_limit)
          --pl.g.nstage;
        pl.g.smem_bytes = occu_rn_chain_smem(ds->L, pl.g.nstage, ds->desc.max_abundance, ds->D, bt);
        pl.rn_global = false;
      }
    }
    if (pl.g.smem_bytes > ds->smem_limit)
      return fail(BL_ERR_UNSUPPORTED, "shape needs %zu B of shared memory per block (> %zu)", pl.g.smem_bytes,
                  ds->smem_limit);
    EvalParams p;
    fill_params(ds, p);
    p.chain_bt = pl.chain_bt;
    p.chain_variant = pl.chain_variant;
    p.nch = pl.nch;
    p.CB = pl.g.CB;  // K1s picks its instantiation by the chains per block
    int occ = 0;
    cudaError_t e = pl.chain_kernel == 1   ? launch_occu_chain(p, dim3(1), pl.g.smem_bytes, nullptr, &occ)
                    : pl.chain_kernel == 5 ? launch_occu_signed(p, dim3(1), pl.g.smem_bytes, nullptr, &occ)
                    : pl.chain_kernel == 6 ? launch_occu_rn2(p, dim3(1), pl.g.smem_bytes, nullptr, &occ)
                    : pl.chain_kernel == 2 ? launch_occu_rn_chain(p, dim3(1), pl.g.smem_bytes, nullptr, &occ)
                    : pl.chain_kernel == 3 ? launch_occu_cop_chain(p, dim3(1), pl.g.smem_bytes, nullptr, &occ)
                    : pl.chain_kernel == 4 ? launch_occu_cs_chain(p, dim3(1), pl.g.smem_bytes, nullptr, &occ)
                    : pl.chain_kernel == 7 ? launch_occu_small(p, dim3(1), pl.g.smem_bytes, nullptr, &occ)
                                           : launch_model(ds, p, dim3(1), pl.g.smem_bytes, nullptr, &occ);
    if (e != cudaSuccess) return fail(BL_ERR_CUDA, "occupancy query: %s", cudaGetErrorString(e));
    if (occ < 1) return fail(BL_ERR_UNSUPPORTED, "kernel does not fit on an SM (smem %zu B)", pl.g.smem_bytes);
    pl.occupancy = occ;
    int64_t want = (int64_t)ds->num_sms * occ / pl.g.n_chunks;
    if (want < 1) want = 1;
    pl.g.nsplit = (int)(want < pl.g.n_block_tiles ? want : pl.g.n_block_tiles);
    if (pl.g.nsplit < 1) pl.g.nsplit = 1;
    it = ds->plans.emplace(C, pl).first;
  }
  Plan& pl = it->second;
  const size_t need = (size_t)pl.g.nsplit * C * (1 + ds->D);
  const size_t rn_need = pl.rn_global ? (size_t)(ds->desc.max_abundance + 1) * pl.g.nsplit * pl.g.n_chunks *
                                            kBlockThreads * elem_size(ds->desc.dtype)
                                      : 0;
  if (need > ds->partial_cap || (size_t)pl.g.n_chunks > ds->counters_cap || (size_t)C > ds->sums_cap ||
      rn_need > ds->rn_scratch_cap) {
    // (re)allocation synchronises; callers that must not sync pass desc.max_chains up front
    cudaDeviceSynchronize();
    if (need > ds->partial_cap) {
      cudaFree(ds->partial);
      if (cudaMalloc(&ds->partial, need * sizeof(double)) != cudaSuccess)
        return fail(BL_ERR_NOMEM, "cudaMalloc(partials, %zu B) failed", need * sizeof(double));
      ds->partial_cap = need;
    }
    if ((size_t)pl.g.n_chunks > ds->counters_cap) {
      cudaFree(ds->counters);
      size_t n = (size_t)pl.g.n_chunks + 16;  // kTicketStride tickets per chunk (final + group tickets, engine.cuh)
      if (cudaMalloc(&ds->counters, n * kTicketStride * sizeof(unsigned int)) != cudaSuccess)
        return fail(BL_ERR_NOMEM, "counters");
      cudaMemset(ds->counters, 0, n * kTicketStride * sizeof(unsigned int));
      ds->counters_cap = n;
    }
    if (rn_need > ds->rn_scratch_cap) {
      cudaFree(ds->rn_scratch);
      if (cudaMalloc(&ds->rn_scratch, rn_need) != cudaSuccess) return fail(BL_ERR_NOMEM, "occu_rn scratch (%zu B)", rn_need);
      ds->rn_scratch_cap = rn_need;
    }
    if ((size_t)C > ds->sums_cap) {
      cudaFree(ds->sums);
      if (cudaMalloc(&ds->sums, (size_t)C * (1 + ds->D) * sizeof(double)) != cudaSuccess)
        return fail(BL_ERR_NOMEM, "sums");
      ds->sums_cap = C;
    }
  }
  *out = &pl;
  return BL_OK;
}

// occu with random effects: own kernel (lane = site, elementwise gradient outputs), own workspace sizing
static int eval_device_re(bl_dataset* ds, const void* theta, int C, void* logp, void* grad, cudaStream_t st,
                          double* logp64) {
  if (ds->comm) return fail(BL_ERR_UNSUPPORTED, "random effects are not site-sharded");
  const int64_t S = ds->desc.n_sites;
  const int nqs = 1 + occu_re_n_small(ds->L, S, ds->desc.flags);
  int grid_x = (int)std::min<int64_t>((S + kBlockThreads - 1) / kBlockThreads, (int64_t)ds->num_sms * 4);
  if (grid_x < 1) grid_x = 1;
  const size_t need = (size_t)grid_x * C * nqs;
  if (need > ds->partial_cap || (size_t)C > ds->counters_cap) {
    cudaDeviceSynchronize();
    if (need > ds->partial_cap) {
      cudaFree(ds->partial);
      if (cudaMalloc(&ds->partial, need * sizeof(double)) != cudaSuccess) return fail(BL_ERR_NOMEM, "partials");
      ds->partial_cap = need;
    }
    if ((size_t)C > ds->counters_cap) {
      cudaFree(ds->counters);
      const size_t n = (size_t)C + 16;
      if (cudaMalloc(&ds->counters, n * sizeof(unsigned int)) != cudaSuccess) return fail(BL_ERR_NOMEM, "counters");
      cudaMemset(ds->counters, 0, n * sizeof(unsigned int));
      ds->counters_cap = n;
    }
  }
  EvalParams p;
  fill_params(ds, p);
  p.theta = theta; p.logp = logp; p.logp64 = logp64; p.grad = grad;
  p.partial = ds->partial; p.counters = ds->counters; p.C = C;
  const double sc_s = ds->desc.prior_fp_a > 0 ? ds->desc.prior_fp_a : 1.0;
  const double sc_o = ds->desc.prior_fp_b > 0 ? ds->desc.prior_fp_b : 1.0;
  cudaError_t e = launch_occu_re(p, ds->desc.dtype, S, grid_x, sc_s, sc_o, st);
  if (e != cudaSuccess) return fail(BL_ERR_CUDA, "eval launch: %s", cudaGetErrorString(e));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return BL_OK;
}

int eval_device(bl_dataset* ds, const void* theta, int C, void* logp, void* grad, cudaStream_t st, int allreduce,
                double* logp64) {
  if (!ds->species.empty()) return eval_device_multi(ds, theta, C, logp, grad, st, logp64, fill_params);
  if (ds->re) return eval_device_re(ds, theta, C, logp, grad, st, logp64);
  Plan* pl = nullptr;
  int rc = plan_for(ds, C, &pl);
  if (rc) return rc;
  EvalParams p;
  fill_params(ds, p);
  p.theta = theta;
  p.logp = logp;
  p.logp64 = logp64;
  p.grad = grad;
  p.partial = ds->partial;
  p.counters = ds->counters;
  p.sums = ds->sums;
  p.allreduce = (allreduce || ds->comm) ? 1 : 0;
  p.C = C;
  p.CB = pl->g.CB;
  p.WC = pl->g.WC;
  p.WS = pl->g.WS;
  p.nstage = pl->g.nstage;
  p.chain_bt = pl->chain_bt;
  p.chain_variant = pl->chain_variant;
  p.nch = pl->nch;
  p.nsplit = pl->g.nsplit;
  p.n_block_tiles = pl->g.n_block_tiles;
  p.rn_scratch_off = pl->rn_scratch_off;
  p.rn_scratch_global = pl->rn_global ? ds->rn_scratch : nullptr;
  const dim3 grid(pl->g.nsplit, pl->g.n_chunks);
  if (pl->chain_kernel == 5) p.packed = ds->packed_signed;
  if (pl->chain_kernel == 6) p.packed = ds->packed_rn2;
  cudaError_t e = pl->chain_kernel == 1   ? launch_occu_chain(p, grid, pl->g.smem_bytes, st, nullptr)
                  : pl->chain_kernel == 5 ? launch_occu_signed(p, grid, pl->g.smem_bytes, st, nullptr)
                  : pl->chain_kernel == 6 ? launch_occu_rn2(p, grid, pl->g.smem_bytes, st, nullptr)
                  : pl->chain_kernel == 2 ? launch_occu_rn_chain(p, grid, pl->g.smem_bytes, st, nullptr)
                  : pl->chain_kernel == 3 ? launch_occu_cop_chain(p, grid, pl->g.smem_bytes, st, nullptr)
                  : pl->chain_kernel == 4 ? launch_occu_cs_chain(p, grid, pl->g.smem_bytes, st, nullptr)
                  : pl->chain_kernel == 7 ? launch_occu_small(p, grid, pl->g.smem_bytes, st, nullptr)
                                          : launch_model(ds, p, grid, pl->g.smem_bytes, st, nullptr);
  if (e != cudaSuccess) return fail(BL_ERR_CUDA, "eval launch: %s", cudaGetErrorString(e));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (ds->comm) return comm_post_eval(ds, p, st);  // cross-rank sum of the raw sums, then priors
  return BL_OK;
}

}  // namespace bl

using namespace bl;

#define CU_TRY(expr)                                                                              \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess) return fail(BL_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(_e));     \
  } while (0)

extern "C" {

int bl_version(void) { return BL_ABI_VERSION; }

const char* bl_strerror(int status) {
  switch (status) {
    case BL_OK: return "ok";
    case BL_ERR_INVALID: return "invalid argument";
    case BL_ERR_UNSUPPORTED: return "option outside the accelerated path (no fallback)";
    case BL_ERR_CUDA: return "CUDA error";
    case BL_ERR_NCCL: return "NCCL error";
    case BL_ERR_BAD_DATA: return "bad data";
    case BL_ERR_NOMEM: return "out of memory";
    default: return "unknown";
  }
}

const char* bl_last_error(void) { return g_err; }

int bl_device_count(int* count) {
  if (!count) return fail(BL_ERR_INVALID, "count is NULL");
  *count = 0;
  CU_TRY(cudaGetDeviceCount(count));
  return BL_OK;
}

int64_t bl_launch_count(void) { return g_launches.load(); }

int bl_dataset_create(const bl_desc* d, const void* y, const void* X, const void* W, const void* T,
                      bl_dataset** out) {
  if (!d || !out) return fail(BL_ERR_INVALID, "desc/out is NULL");
  *out = nullptr;
  if (d->abi_version != BL_ABI_VERSION) return fail(BL_ERR_INVALID, "ABI version %d != %d", d->abi_version, BL_ABI_VERSION);
  if (d->model < 0 || d->model > BL_MODEL_OCCU_CS) return fail(BL_ERR_INVALID, "unknown model %d", d->model);
  if ((d->dtype != BL_F32 && d->dtype != BL_F64) || (d->data_dtype != BL_F32 && d->data_dtype != BL_F64))
    return fail(BL_ERR_INVALID, "dtype must be BL_F32 or BL_F64");
  if (d->n_sites < 0 || d->n_periods < 1 || d->n_replicates < 1 || d->n_site_covs < 0 || d->n_obs_covs < 0)
    return fail(BL_ERR_INVALID, "bad shape S=%lld P=%d J=%d Ks=%d Ko=%d", (long long)d->n_sites, d->n_periods,
                d->n_replicates, d->n_site_covs, d->n_obs_covs);
  if (d->n_species < 1) return fail(BL_ERR_INVALID, "n_species must be >= 1");
  if (d->n_species > 1) {
    // the species plate (occu.py:182-186): one likelihood-only child per species, extras shared (multi.cu)
    if (d->flags & (BL_FLAG_SITE_RE | BL_FLAG_OBS_RE))
      return fail(BL_ERR_UNSUPPORTED, "random effects with n_species > 1 are outside the accelerated path");
    bl_dataset* parent = new (std::nothrow) bl_dataset();
    if (!parent) return fail(BL_ERR_NOMEM, "host allocation failed");
    parent->desc = *d;
    const size_t ds_in = elem_size(d->data_dtype);
    const size_t stride = (size_t)d->n_sites * d->n_periods * d->n_replicates * ds_in;
    int rc = BL_OK;
    for (int sp = 0; sp < d->n_species && rc == BL_OK; ++sp) {
      bl_desc dc = *d;
      dc.n_species = 1;
      dc.flags &= ~BL_FLAG_PRIOR;
      bl_dataset* child = nullptr;
      rc = bl_dataset_create(&dc, (const char*)y + sp * stride, X, W, T, &child);
      if (rc == BL_OK) parent->species.push_back(child);
    }
    if (rc == BL_OK) {
      bl_dataset* c0 = parent->species[0];
      parent->L = c0->L;
      parent->n_extras = c0->n_extras;
      parent->D = d->n_species * (c0->L.ks + 1 + c0->L.ko + 1) + c0->n_extras;
      parent->DS = parent->D;
      parent->num_sms = c0->num_sms;
      parent->smem_limit = c0->smem_limit;
      for (bl_dataset* c : parent->species) { parent->n_masked += c->n_masked; parent->packed_bytes += c->packed_bytes; }
      cudaError_t e2 = cudaStreamCreateWithFlags(&parent->own_stream, cudaStreamNonBlocking);
      if (e2 == cudaSuccess) e2 = cudaEventCreate(&parent->ev0);
      if (e2 == cudaSuccess) e2 = cudaEventCreate(&parent->ev1);
      if (e2 != cudaSuccess) rc = fail(BL_ERR_CUDA, "stream/event create: %s", cudaGetErrorString(e2));
    }
    if (rc != BL_OK) { bl_dataset_destroy(parent); return rc; }
    *out = parent;
    return BL_OK;
  }
  if (d->n_site_covs > kMaxCov || d->n_obs_covs > kMaxCov)
    return fail(BL_ERR_UNSUPPORTED, "more than %d covariates per predictor", kMaxCov);
  const bool fpc = d->flags & BL_FLAG_FP_CONSTANT, fpu = d->flags & BL_FLAG_FP_UNOCCUPIED;
  if (d->model == BL_MODEL_OCCU && fpc && fpu)
    return fail(BL_ERR_INVALID, "false_positives_constant and false_positives_unoccupied cannot both be True");  // occu.py:112-114
  if (d->model == BL_MODEL_OCCU_RN && fpu) return fail(BL_ERR_INVALID, "occu_rn has no false_positives_unoccupied");
  if (d->model == BL_MODEL_NMIXTURE && (fpc || fpu)) return fail(BL_ERR_INVALID, "nmixture has no false-positive options");
  if (d->model == BL_MODEL_OCCU_CS && (fpc || fpu)) return fail(BL_ERR_INVALID, "occu_cs has no false-positive options");
  if (d->model == BL_MODEL_OCCU_CS && (d->flags & BL_FLAG_PRIOR) &&
      (d->prior_fp_a <= 0 || d->prior_fp_b <= 0 || d->prior_fp_rate <= 0))
    return fail(BL_ERR_INVALID, "occu_cs priors: Gamma(a, b) of sigma and the Normal scale of mu must be positive");
  if ((d->model == BL_MODEL_OCCU_RN || d->model == BL_MODEL_NMIXTURE) && (d->max_abundance < 1 || d->max_abundance > 1023))
    return fail(BL_ERR_INVALID, "max_abundance must be in [1, 1023]");
  const bool re = (d->flags & (BL_FLAG_SITE_RE | BL_FLAG_OBS_RE)) != 0;
  if (re && (d->model != BL_MODEL_OCCU || fpc || fpu))
    return fail(BL_ERR_UNSUPPORTED, "random effects are accelerated for occu without false-positive extras only");
  if (d->n_sites > 0 && (!y || !X || !W)) return fail(BL_ERR_INVALID, "y/X/W is NULL");
  if ((d->flags & BL_FLAG_PRIOR) && (d->prior_beta_scale <= 0 || d->prior_alpha_scale <= 0))
    return fail(BL_ERR_INVALID, "prior scales must be positive");

  CU_TRY(cudaSetDevice(d->device));
  bl_dataset* ds = new (std::nothrow) bl_dataset();
  if (!ds) return fail(BL_ERR_NOMEM, "host allocation failed");
  ds->desc = *d;
  ds->L = make_layout(d->model, d->n_sites, d->n_periods, d->n_replicates, d->n_site_covs, d->n_obs_covs);
  // strict math: the libm-accurate engine, except for occu shapes K1d covers (its STRICT instantiations)
  ds->strict_chain = (d->flags & BL_FLAG_STRICT_MATH) && d->model == BL_MODEL_OCCU && !re && d->n_species <= 1 &&
                     occu_signed_supported(d->dtype, d->n_site_covs, d->n_obs_covs, d->flags);
  ds->force_engine = (d->flags & BL_FLAG_STRICT_MATH) != 0 && !ds->strict_chain;
  ds->n_extras = d->model == BL_MODEL_OCCU_CS ? 4 : (fpc ? 1 : 0) + (fpu ? 1 : 0);
  ds->D = d->n_site_covs + 1 + d->n_obs_covs + 1 + ds->n_extras;
  int derived = d->model == BL_MODEL_OCCU ? occu_derived_slots(d->flags)
                : d->model == BL_MODEL_OCCU_RN ? occu_rn_derived_slots(d->flags)
                : d->model == BL_MODEL_OCCU_COP ? occu_cop_derived_slots(d->flags)
                : d->model == BL_MODEL_OCCU_CS ? 4 : 0;
  ds->DS = ds->D + derived;
  if (re) {
    const int64_t Dre = occu_re_theta_dim(ds->L, d->n_sites, d->flags);
    if (Dre > (int64_t)1 << 30) { delete ds; return fail(BL_ERR_UNSUPPORTED, "random-effect vector too long"); }
    ds->re = true;
    ds->force_engine = true;
    ds->D = (int)Dre;
    ds->DS = ds->D;
    ds->n_extras = ((d->flags & BL_FLAG_SITE_RE) ? 1 : 0) + ((d->flags & BL_FLAG_OBS_RE) ? 1 : 0);
  }
  cudaDeviceProp prop;
  cudaError_t e = cudaGetDeviceProperties(&prop, d->device);
  if (e != cudaSuccess) { delete ds; return fail(BL_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e)); }
  ds->num_sms = prop.multiProcessorCount;
  // per-block budget for DYNAMIC shared memory: minus the kernels' static part (s_is_last, the 512-byte staging of the
  // final reduction) and the 1 KB the driver reserves per block, so that limit / n blocks really are resident
  ds->smem_limit = prop.sharedMemPerBlockOptin - 4096;

  const Layout& L = ds->L;
  const size_t es = elem_size(d->dtype), ds_in = elem_size(d->data_dtype);
  ds->packed_bytes = (size_t)L.n_tiles_padded * L.F * kWarp * es;
  if (ds->packed_bytes == 0) ds->packed_bytes = 16;
  void *dy = nullptr, *dX = nullptr, *dW = nullptr, *dT = nullptr;
  int* d_err = nullptr;
  unsigned long long* d_nm = nullptr;
  const size_t nobs = (size_t)L.n_units * L.J;
  int rc = BL_OK;
  do {
#define CU_BRK(expr)                                                                   \
    if ((e = (expr)) != cudaSuccess) { rc = fail(BL_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e)); break; }
    CU_BRK(cudaMalloc(&ds->packed, ds->packed_bytes));
    CU_BRK(cudaMemset(ds->packed, 0, ds->packed_bytes));
    CU_BRK(cudaMalloc(&d_err, sizeof(int)));
    CU_BRK(cudaMemset(d_err, 0, sizeof(int)));
    CU_BRK(cudaMalloc(&d_nm, sizeof(unsigned long long)));
    CU_BRK(cudaMemset(d_nm, 0, sizeof(unsigned long long)));
    if (L.n_units > 0) {
      CU_BRK(cudaMalloc(&dy, nobs * ds_in));
      CU_BRK(cudaMalloc(&dX, (size_t)d->n_sites * (L.ks > 0 ? L.ks : 1) * ds_in));
      CU_BRK(cudaMalloc(&dW, nobs * (L.ko > 0 ? L.ko : 1) * ds_in));
      CU_BRK(cudaMemcpy(dy, y, nobs * ds_in, cudaMemcpyDefault));
      if (L.ks > 0) CU_BRK(cudaMemcpy(dX, X, (size_t)d->n_sites * L.ks * ds_in, cudaMemcpyDefault));
      if (L.ko > 0) CU_BRK(cudaMemcpy(dW, W, nobs * L.ko * ds_in, cudaMemcpyDefault));
      if (T && d->model == BL_MODEL_OCCU_COP) {
        CU_BRK(cudaMalloc(&dT, nobs * ds_in));
        CU_BRK(cudaMemcpy(dT, T, nobs * ds_in, cudaMemcpyDefault));
      }
      CU_BRK(launch_pack(d->data_dtype, d->dtype, dy, dX, dW, dT, ds->packed, L, d->model, d->max_abundance, d_err,
                         d_nm, nullptr));
      g_launches.fetch_add(1);
      if (d->model == BL_MODEL_OCCU && !ds->force_engine &&
          occu_signed_supported(d->dtype, L.ks, L.ko, d->flags)) {
        // second packing for the lane = chain kernel K1d: AoS signed site records (occu_signed.cu)
        ds->packed_signed_bytes = occu_signed_bytes(L);
        CU_BRK(cudaMalloc(&ds->packed_signed, ds->packed_signed_bytes));
        CU_BRK(launch_repack_signed(ds->packed, ds->packed_signed, L, nullptr));
        g_launches.fetch_add(1);
      }
      if (d->model == BL_MODEL_OCCU_RN && !ds->force_engine &&
          occu_rn2_supported(d->dtype, L.ks, L.ko, L.J, d->max_abundance, d->flags)) {
        // second packing for the lane = chain kernel K2d: visits sorted detections-first (occu_rn2.cu)
        ds->packed_rn2_bytes = occu_rn2_bytes(L);
        CU_BRK(cudaMalloc(&ds->packed_rn2, ds->packed_rn2_bytes));
        CU_BRK(launch_repack_rn2(ds->packed, ds->packed_rn2, L, nullptr));
        g_launches.fetch_add(1);
      }
      CU_BRK(cudaDeviceSynchronize());
    }
    int h_err = 0;
    unsigned long long h_nm = 0;
    CU_BRK(cudaMemcpy(&h_err, d_err, sizeof(int), cudaMemcpyDeviceToHost));
    CU_BRK(cudaMemcpy(&h_nm, d_nm, sizeof(h_nm), cudaMemcpyDeviceToHost));
    ds->n_masked = (int64_t)h_nm;
    if (h_err & 1) { rc = fail(BL_ERR_BAD_DATA, "detections must be binary (0/1) or non-finite (missing)"); break; }
    if (h_err & 2) { rc = fail(BL_ERR_BAD_DATA, "counts must be >= 0 and session_duration finite"); break; }
    if (h_err & 4) { rc = fail(BL_ERR_BAD_DATA, "nmixture counts must be integers in [0, max_abundance]"); break; }
#undef CU_BRK
  } while (0);
  cudaFree(dy); cudaFree(dX); cudaFree(dW); cudaFree(dT); cudaFree(d_err); cudaFree(d_nm);
  if (rc == BL_OK && d->model == BL_MODEL_OCCU_COP) {
    // data-only constant  sum m (y log T - lgamma(y+1)), in double, fixed order (host arrays only)
    double acc = 0.0;
    auto rd = [&](const void* a, size_t i) -> double {
      return d->data_dtype == BL_F32 ? (double)((const float*)a)[i] : ((const double*)a)[i];
    };
    auto rd_c = [&](const void* a, size_t i) -> double {  // value as the compute dtype sees it
      double v = rd(a, i);
      return d->dtype == BL_F32 ? (double)(float)v : v;
    };
    for (int64_t u = 0; u < L.n_units; ++u) {
      const int64_t s = u / L.P;
      bool site_nan = false;
      for (int k = 0; k < L.ks; ++k) site_nan |= std::isnan(rd(X, (size_t)s * L.ks + k));
      for (int j = 0; j < L.J; ++j) {
        const size_t o = (size_t)u * L.J + j;
        bool cov_nan = site_nan;
        for (int k = 0; k < L.ko; ++k) cov_nan |= std::isnan(rd(W, o * L.ko + k));
        const double yv = rd_c(y, o);
        if (!std::isfinite(yv) || cov_nan) continue;
        const double tv = T ? rd_c(T, o) : 1.0;
        acc += (yv > 0 ? yv * std::log(tv) : 0.0) - std::lgamma(yv + 1.0);
      }
    }
    ds->cop_const = acc;
  }
  if (rc == BL_OK) {
    cudaError_t e2 = cudaStreamCreateWithFlags(&ds->own_stream, cudaStreamNonBlocking);
    if (e2 == cudaSuccess) e2 = cudaEventCreate(&ds->ev0);
    if (e2 == cudaSuccess) e2 = cudaEventCreate(&ds->ev1);
    if (e2 != cudaSuccess) rc = fail(BL_ERR_CUDA, "stream/event create: %s", cudaGetErrorString(e2));
  }
  if (rc == BL_OK && d->max_chains > 0 && !ds->re) {
    Plan* pl = nullptr;
    rc = plan_for(ds, d->max_chains, &pl);
  }
  if (rc != BL_OK) { bl_dataset_destroy(ds); return rc; }
  *out = ds;
  return BL_OK;
}

int bl_dataset_destroy(bl_dataset* ds) {
  if (!ds) return BL_OK;
  cudaSetDevice(ds->desc.device);
  for (bl_dataset* c : ds->species) bl_dataset_destroy(c);
  ds->species.clear();
  cudaFree(ds->ms_theta); cudaFree(ds->ms_lp); cudaFree(ds->ms_lp64); cudaFree(ds->ms_grad);
  comm_destroy(ds);
  cudaFree(ds->packed); cudaFree(ds->packed_signed); cudaFree(ds->packed_rn2); cudaFree(ds->partial); cudaFree(ds->counters); cudaFree(ds->sums);
  cudaFree(ds->rn_scratch);
  cudaFree(ds->d_theta); cudaFree(ds->d_out);
  if (ds->h_theta) cudaFreeHost(ds->h_theta);
  if (ds->h_out) cudaFreeHost(ds->h_out);
  if (ds->ev0) cudaEventDestroy(ds->ev0);
  if (ds->ev1) cudaEventDestroy(ds->ev1);
  if (ds->own_stream) cudaStreamDestroy(ds->own_stream);
  delete ds;
  return BL_OK;
}

int bl_dataset_info(const bl_dataset* ds, bl_info* info) {
  if (!ds || !info) return fail(BL_ERR_INVALID, "NULL argument");
  const Layout& L = ds->L;
  const int64_t es = (int64_t)elem_size(ds->desc.dtype);
  info->theta_dim = ds->D;
  info->n_extras = ds->n_extras;
  info->n_units = L.n_units;
  info->packed_bytes = (int64_t)ds->packed_bytes;
  // bytes(y) + bytes(X) + bytes(W) (+ bytes(T)) + theta + (logp, grad), in the compute dtype
  int64_t per_unit = L.J + (int64_t)L.J * L.ko + (ds->desc.model == BL_MODEL_OCCU_COP ? L.J : 0);
  info->algorithmic_bytes = es * (L.n_units * per_unit + ds->desc.n_sites * L.ks) + es * (2 * ds->D + 1);
  info->n_masked = ds->n_masked;
  info->fields_per_unit = L.F;
  const bool fp = ds->n_extras > 0;
  info->kernel_variant = ds->desc.model == BL_MODEL_OCCU ? occu_has_specialisation(L.ks, L.ko, fp)
                         : ds->desc.model == BL_MODEL_OCCU_CS ? occu_cs_has_specialisation(L.ks, L.ko) : 0;
  return BL_OK;
}

int bl_plan_kernel(bl_dataset* ds, int32_t n_chains, int32_t* kernel_id, int32_t* grid_x, int32_t* grid_y,
                   int32_t* block_threads) {
  if (!ds || !kernel_id || n_chains < 1) return fail(BL_ERR_INVALID, "bad argument");
  if (!ds->species.empty() || ds->re) return fail(BL_ERR_UNSUPPORTED, "composite / random-effects handles plan per child");
  CU_TRY(cudaSetDevice(ds->desc.device));
  Plan* pl = nullptr;
  const int rc = plan_for(ds, n_chains, &pl);
  if (rc) return rc;
  *kernel_id = pl->chain_kernel;
  if (grid_x) *grid_x = pl->g.nsplit;
  if (grid_y) *grid_y = pl->g.n_chunks;
  if (block_threads) *block_threads = pl->chain_kernel ? pl->chain_bt : kBlockThreads;
  return BL_OK;
}

int bl_dataset_export_mask(const bl_dataset* ds, uint8_t* mask_out) {
  if (!ds || !mask_out) return fail(BL_ERR_INVALID, "NULL argument");
  const size_t n = (size_t)ds->L.n_units * ds->L.J;
  if (n == 0) return BL_OK;
  if (!ds->species.empty()) {  // n_species > 1: (Sp, S, P, J), species-major like obs
    for (size_t sp = 0; sp < ds->species.size(); ++sp) {
      const int rc = bl_dataset_export_mask(ds->species[sp], mask_out + sp * n);
      if (rc) return rc;
    }
    return BL_OK;
  }
  CU_TRY(cudaSetDevice(ds->desc.device));
  uint8_t* d_mask = nullptr;
  CU_TRY(cudaMalloc(&d_mask, n));
  cudaError_t e = launch_export_mask(ds->desc.dtype, ds->packed, d_mask, ds->L, nullptr);
  g_launches.fetch_add(1);
  if (e == cudaSuccess) e = cudaMemcpy(mask_out, d_mask, n, cudaMemcpyDeviceToHost);
  cudaFree(d_mask);
  if (e != cudaSuccess) return fail(BL_ERR_CUDA, "export mask: %s", cudaGetErrorString(e));
  return BL_OK;
}

int bl_eval(bl_dataset* ds, const void* theta, int32_t n_chains, void* logp, void* grad, bl_stream stream) {
  if (!ds || !theta || !logp || !grad) return fail(BL_ERR_INVALID, "NULL argument");
  if (n_chains < 1) return fail(BL_ERR_INVALID, "n_chains must be >= 1");
  CU_TRY(cudaSetDevice(ds->desc.device));
  return eval_device(ds, theta, n_chains, logp, grad, (cudaStream_t)stream, 0, nullptr);
}

static int ensure_host_staging(bl_dataset* ds, int C) {
  if (C <= ds->host_cap) return BL_OK;
  const size_t es = elem_size(ds->desc.dtype);
  cudaDeviceSynchronize();
  cudaFree(ds->d_theta); cudaFree(ds->d_out);
  if (ds->h_theta) cudaFreeHost(ds->h_theta);
  if (ds->h_out) cudaFreeHost(ds->h_out);
  ds->d_theta = ds->d_out = ds->h_theta = ds->h_out = nullptr;
  ds->host_cap = 0;
  const size_t nt = (size_t)C * ds->D * es, no = (size_t)C * (1 + ds->D) * es;
  CU_TRY(cudaMalloc(&ds->d_theta, nt));
  CU_TRY(cudaMalloc(&ds->d_out, no));
  CU_TRY(cudaMallocHost(&ds->h_theta, nt));
  CU_TRY(cudaMallocHost(&ds->h_out, no));
  ds->host_cap = C;
  return BL_OK;
}

int bl_eval_host(bl_dataset* ds, const void* theta, int32_t n_chains, void* logp, void* grad) {
  if (!ds || !theta || !logp || !grad) return fail(BL_ERR_INVALID, "NULL argument");
  if (n_chains < 1) return fail(BL_ERR_INVALID, "n_chains must be >= 1");
  CU_TRY(cudaSetDevice(ds->desc.device));
  int rc = ensure_host_staging(ds, n_chains);
  if (rc) return rc;
  const size_t es = elem_size(ds->desc.dtype);
  const size_t nt = (size_t)n_chains * ds->D * es, nl = (size_t)n_chains * es;
  cudaStream_t st = ds->own_stream;
  memcpy(ds->h_theta, theta, nt);  // pageable caller buffer -> pinned staging
  CU_TRY(cudaMemcpyAsync(ds->d_theta, ds->h_theta, nt, cudaMemcpyHostToDevice, st));
  char* d_logp = (char*)ds->d_out;
  char* d_grad = d_logp + nl;
  rc = eval_device(ds, ds->d_theta, n_chains, d_logp, d_grad, st, 0, nullptr);
  if (rc) return rc;
  CU_TRY(cudaMemcpyAsync(ds->h_out, ds->d_out, nl + nt, cudaMemcpyDeviceToHost, st));
  CU_TRY(cudaStreamSynchronize(st));
  memcpy(logp, ds->h_out, nl);
  memcpy(grad, (char*)ds->h_out + nl, nt);
  return BL_OK;
}

int bl_obs_loglik(bl_dataset* ds, const void* theta, int32_t n_draws, float* lppd_out, float* var_out) {
  if (!ds || !theta || !lppd_out) return fail(BL_ERR_INVALID, "NULL argument");
  if (n_draws < 1) return fail(BL_ERR_INVALID, "n_draws must be >= 1");
  if (ds->desc.model != BL_MODEL_OCCU || ds->re || !ds->species.empty() ||
      (ds->desc.flags & (BL_FLAG_FP_CONSTANT | BL_FLAG_FP_UNOCCUPIED)))
    return fail(BL_ERR_UNSUPPORTED, "per-observation log-likelihood: occu without false-positive / random-effect extras");
  CU_TRY(cudaSetDevice(ds->desc.device));
  const size_t es = elem_size(ds->desc.dtype);
  const size_t nobs = (size_t)ds->L.n_units * ds->L.J;
  if (nobs == 0) return BL_OK;
  void* d_theta = nullptr;
  float *d_l = nullptr, *d_v = nullptr;
  int rc = BL_OK;
  cudaError_t e = cudaSuccess;
  do {
#define CU_BRK(expr) if ((e = (expr)) != cudaSuccess) { rc = fail(BL_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e)); break; }
    CU_BRK(cudaMalloc(&d_theta, (size_t)n_draws * ds->D * es));
    CU_BRK(cudaMalloc(&d_l, nobs * sizeof(float)));
    if (var_out) CU_BRK(cudaMalloc(&d_v, nobs * sizeof(float)));
    CU_BRK(cudaMemcpy(d_theta, theta, (size_t)n_draws * ds->D * es, cudaMemcpyHostToDevice));
    EvalParams p;
    fill_params(ds, p);
    p.theta = d_theta;
    CU_BRK(launch_obs_loglik(p, ds->desc.dtype, n_draws, d_l, d_v, ds->own_stream));
    g_launches.fetch_add(1);
    CU_BRK(cudaStreamSynchronize(ds->own_stream));
    CU_BRK(cudaMemcpy(lppd_out, d_l, nobs * sizeof(float), cudaMemcpyDeviceToHost));
    if (var_out) CU_BRK(cudaMemcpy(var_out, d_v, nobs * sizeof(float), cudaMemcpyDeviceToHost));
#undef CU_BRK
  } while (0);
  cudaFree(d_theta); cudaFree(d_l); cudaFree(d_v);
  return rc;
}

int bl_site_summary(bl_dataset* ds, const void* theta, int32_t n_draws, float* out) {
  if (!ds || !theta || !out) return fail(BL_ERR_INVALID, "NULL argument");
  if (n_draws < 1) return fail(BL_ERR_INVALID, "n_draws must be >= 1");
  if (ds->re) return fail(BL_ERR_UNSUPPORTED, "per-unit summaries are not built for the random-effects likelihood");
  if (!ds->species.empty()) return fail(BL_ERR_UNSUPPORTED, "per-unit summaries: use one handle per species");
  CU_TRY(cudaSetDevice(ds->desc.device));
  const size_t es = elem_size(ds->desc.dtype);
  const size_t U = (size_t)ds->L.n_units;
  if (U == 0) return BL_OK;
  void* d_theta = nullptr;
  float* d_out = nullptr;
  void* d_scratch = nullptr;
  int rc = BL_OK;
  cudaError_t e = cudaSuccess;
  do {
#define CU_BRK(expr) if ((e = (expr)) != cudaSuccess) { rc = fail(BL_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e)); break; }
    CU_BRK(cudaMalloc(&d_theta, (size_t)n_draws * ds->D * es));
    CU_BRK(cudaMalloc(&d_out, 4 * U * sizeof(float)));
    CU_BRK(cudaMemcpy(d_theta, theta, (size_t)n_draws * ds->D * es, cudaMemcpyHostToDevice));
    EvalParams p;
    fill_params(ds, p);
    p.theta = d_theta;
    p.C = n_draws;
    if (ds->desc.model == BL_MODEL_OCCU_RN || ds->desc.model == BL_MODEL_NMIXTURE) {
      const size_t blocks = (U + kBlockThreads - 1) / kBlockThreads;
      CU_BRK(cudaMalloc(&d_scratch, (size_t)(ds->desc.max_abundance + 1) * blocks * kBlockThreads * es));
      p.rn_scratch_global = d_scratch;
    }
    switch (ds->desc.model) {
      case BL_MODEL_OCCU: e = launch_occu_summary(p, ds->desc.dtype, d_out, nullptr); break;
      case BL_MODEL_OCCU_RN: e = launch_occu_rn_summary(p, ds->desc.dtype, d_out, nullptr); break;
      case BL_MODEL_NMIXTURE: e = launch_nmixture_summary(p, ds->desc.dtype, d_out, nullptr); break;
      case BL_MODEL_OCCU_CS: e = launch_occu_cs_summary(p, ds->desc.dtype, d_out, nullptr); break;
      default: e = launch_occu_cop_summary(p, ds->desc.dtype, d_out, nullptr); break;
    }
    if (e != cudaSuccess) { rc = fail(BL_ERR_CUDA, "summary launch: %s", cudaGetErrorString(e)); break; }
    g_launches.fetch_add(1);
    CU_BRK(cudaMemcpy(out, d_out, 4 * U * sizeof(float), cudaMemcpyDeviceToHost));
#undef CU_BRK
  } while (0);
  cudaFree(d_theta); cudaFree(d_out); cudaFree(d_scratch);
  return rc;
}

int bl_eval_timed(bl_dataset* ds, const void* theta, int32_t n_chains, void* logp, void* grad, bl_stream stream,
                  int32_t iters, float* ms_per_eval) {
  if (!ds || !ms_per_eval || iters < 1) return fail(BL_ERR_INVALID, "bad argument");
  CU_TRY(cudaSetDevice(ds->desc.device));
  cudaStream_t st = (cudaStream_t)stream;
  CU_TRY(cudaEventRecord(ds->ev0, st));
  for (int i = 0; i < iters; ++i) {
    int rc = bl_eval(ds, theta, n_chains, logp, grad, stream);
    if (rc) return rc;
  }
  CU_TRY(cudaEventRecord(ds->ev1, st));
  CU_TRY(cudaEventSynchronize(ds->ev1));
  float ms = 0.f;
  CU_TRY(cudaEventElapsedTime(&ms, ds->ev0, ds->ev1));
  *ms_per_eval = ms / iters;
  return BL_OK;
}

int bl_device_malloc(int32_t device, size_t bytes, void** ptr) {
  if (!ptr) return fail(BL_ERR_INVALID, "ptr is NULL");
  CU_TRY(cudaSetDevice(device));
  CU_TRY(cudaMalloc(ptr, bytes ? bytes : 16));
  return BL_OK;
}
int bl_device_free(void* ptr) { CU_TRY(cudaFree(ptr)); return BL_OK; }
int bl_memcpy_h2d(void* dst, const void* src, size_t bytes, bl_stream stream) {
  CU_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
  return BL_OK;
}
int bl_memcpy_d2h(void* dst, const void* src, size_t bytes, bl_stream stream) {
  CU_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  return BL_OK;
}
int bl_host_malloc_pinned(size_t bytes, void** ptr) {
  if (!ptr) return fail(BL_ERR_INVALID, "ptr is NULL");
  CU_TRY(cudaMallocHost(ptr, bytes ? bytes : 16));
  return BL_OK;
}
int bl_host_free_pinned(void* ptr) { CU_TRY(cudaFreeHost(ptr)); return BL_OK; }
int bl_stream_create(int32_t device, bl_stream* stream) {
  if (!stream) return fail(BL_ERR_INVALID, "stream is NULL");
  CU_TRY(cudaSetDevice(device));
  cudaStream_t s;
  CU_TRY(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  *stream = (bl_stream)s;
  return BL_OK;
}
int bl_stream_destroy(bl_stream stream) { CU_TRY(cudaStreamDestroy((cudaStream_t)stream)); return BL_OK; }
int bl_stream_sync(bl_stream stream) { CU_TRY(cudaStreamSynchronize((cudaStream_t)stream)); return BL_OK; }

int bl_flush_l2(int32_t device, bl_stream stream) {
  static std::mutex mu;
  static std::map<int, void*> bufs;
  const size_t bytes = 256u << 20;  // > 126 MB L2
  CU_TRY(cudaSetDevice(device));
  void* buf = nullptr;
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = bufs.find(device);
    if (it == bufs.end()) {
      CU_TRY(cudaMalloc(&buf, bytes));
      bufs[device] = buf;
    } else {
      buf = it->second;
    }
  }
  CU_TRY(cudaMemsetAsync(buf, 0, bytes, (cudaStream_t)stream));
  return BL_OK;
}

}  // extern "C"
