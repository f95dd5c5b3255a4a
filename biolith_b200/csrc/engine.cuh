// The site-tile engine shared by the occu / occu_rn / occu_cop likelihood kernels.
//
// Mapping (see DESIGN.md "kernel design"):
//   lane  = unit (site x period)          -> every field read is one coalesced 128 B line
//   warp  = (site-group ws, chain-group wc); a block of 8 warps is WS x WC
//   block = (site split bx, chain chunk by); persistent over a contiguous range of block-tiles
//   tile  = WS warp-tiles, staged HBM -> SMEM by ONE cp.async.bulk (TMA) per block-tile through an
//           NSTAGE mbarrier ring, then re-used by every chain of the chunk (CB <= 256 chains).
// Per (unit, chain) the Model functor returns NQ = 1 + D numbers (log-marginal, d/dtheta); they are
// summed over the 32 lanes by xor-shuffles and accumulated in fp64 shared memory by a fixed owner
// (warp (ws,wc) owns chains wc, wc+WC, ... of site-group ws) -> no atomics, deterministic order.
// The last block of each chain chunk (ticket counter) reduces the per-split partials in split
// order and writes logp / grad (+ priors): one launch per evaluation.
#pragma once

#include "common.cuh"

namespace bl {

// lgamma for small positive arguments used by the prior normaliser (host+device, double)
__device__ __forceinline__ double log_sigmoid_d(double x) { return fmin(x, 0.0) - log1p(exp(-fabs(x))); }

// occu_cs extras [mu0, x1 = log(mu1 - mu0), log sigma0, log sigma1] (occu_cs.py:146-154): Normal(0, s) on
// mu0, Normal(0, s) left-truncated at mu0 on mu1 (+ log|J| = x1), Gamma(a, b) on each sigma (+ log|J| = x);
// s = prior_fp_rate, (a, b) = (prior_fp_a, prior_fp_b) for this model.  which = -1: log-density, 0..3: d/dx_i.
__device__ inline double cs_prior(const EvalParams& p, const double x[4], int which) {
  const double s = p.prior_fp_rate, a = p.prior_fp_a, b = p.prior_fp_b;
  const double h2pi = 0.91893853320467274178, rs2 = 0.70710678118654752440;
  const double mu0 = x[0], e1 = exp(x[1]), mu1 = mu0 + e1, t = mu0 / s;
  // log(1 - Phi(t)) and the hazard phi(t) / (1 - Phi(t)), stable for t > 0 through erfcx
  double log_sf, hazard;
  if (t > 0.0) {
    const double ex = erfcx(t * rs2);
    log_sf = log(0.5 * ex) - 0.5 * t * t;
    hazard = 0.79788456080286535588 / ex;  // sqrt(2/pi) / erfcx
  } else {
    const double sf = 0.5 * erfc(t * rs2);
    log_sf = log(sf);
    hazard = exp(-0.5 * t * t - h2pi) / sf;
  }
  switch (which) {
    case -1: {
      double lp = -0.5 * t * t - log(s) - h2pi;
      lp += -0.5 * (mu1 / s) * (mu1 / s) - log(s) - h2pi - log_sf + x[1];
      for (int i = 2; i < 4; ++i) lp += a * log(b) + (a - 1.0) * x[i] - b * exp(x[i]) - lgamma(a) + x[i];
      return lp;
    }
    case 0: return -mu0 / (s * s) - mu1 / (s * s) + hazard / s;
    case 1: return -mu1 / (s * s) * e1 + 1.0;
    case 2: return a - b * exp(x[2]);
    default: return a - b * exp(x[3]);
  }
}

// Adds priors (+ Jacobians) to the raw sums of one chain and writes the outputs.
template <typename T>
__device__ void finalize_chain(const EvalParams& p, int c, int q, double total, bool add_const) {
  const T* theta = reinterpret_cast<const T*>(p.theta) + (size_t)c * p.D;
  T* logp = reinterpret_cast<T*>(p.logp);
  T* grad = reinterpret_cast<T*>(p.grad);
  const int KB = p.L.ks + 1, KA = p.L.ko + 1;
  const bool prior = (p.flags & BL_FLAG_PRIOR) != 0;
  if (q == 0) {
    double lp = total + (add_const ? p.cop_const : 0.0);
    if (prior) {
      const double h2pi = 0.91893853320467274178;  // 0.5*log(2*pi)
      // log(scale) + 0.5 log(2 pi) and 1 / scale come from the host (fill_params): a serial fp64 log / division
      // chain in the one thread every evaluation waits for
      const double nb = p.prior_beta_norm, na = p.prior_alpha_norm;
      const double ib = p.prior_beta_iscale, ia = p.prior_alpha_iscale;
      (void)h2pi;
      for (int i = 0; i < KB; ++i) {
        const double z = ((double)theta[i] - p.prior_beta_loc) * ib;
        lp += -0.5 * z * z - nb;
      }
      for (int i = 0; i < KA; ++i) {
        const double z = ((double)theta[KB + i] - p.prior_alpha_loc) * ia;
        lp += -0.5 * z * z - na;
      }
      if (p.model == BL_MODEL_OCCU_CS) {
        const double x[4] = {(double)theta[KB + KA], (double)theta[KB + KA + 1], (double)theta[KB + KA + 2],
                             (double)theta[KB + KA + 3]};
        lp += cs_prior(p, x, -1);
      }
      for (int i = KB + KA; i < p.D && p.model != BL_MODEL_OCCU_CS; ++i) {
        double x = (double)theta[i];
        if (p.model == BL_MODEL_OCCU_COP) {  // occu_cop: Exponential(rate) on exp(x), + log|J| = x
          lp += log(p.prior_fp_rate) - p.prior_fp_rate * exp(x) + x;
        } else {               // Beta(a,b) on sigmoid(x), + log|J| = log c + log(1-c)
          double a = p.prior_fp_a, b = p.prior_fp_b;
          lp += a * log_sigmoid_d(x) + b * log_sigmoid_d(-x) + lgamma(a + b) - lgamma(a) - lgamma(b);
        }
      }
    }
    logp[c] = (T)lp;
    if (p.logp64) p.logp64[c] = lp;
  } else {
    const int i = q - 1;
    double g = total;
    if (prior) {
      double x = (double)theta[i];
      if (i < KB) g -= (x - p.prior_beta_loc) * (p.prior_beta_iscale * p.prior_beta_iscale);
      else if (i < KB + KA) g -= (x - p.prior_alpha_loc) * (p.prior_alpha_iscale * p.prior_alpha_iscale);
      else if (p.model == BL_MODEL_OCCU_CS) {
        const double xe[4] = {(double)theta[KB + KA], (double)theta[KB + KA + 1], (double)theta[KB + KA + 2],
                              (double)theta[KB + KA + 3]};
        g += cs_prior(p, xe, i - KB - KA);
      } else if (p.model == BL_MODEL_OCCU_COP) g += 1.0 - p.prior_fp_rate * exp(x);
      else {
        double c1 = 1.0 / (1.0 + exp(-x));
        g += p.prior_fp_a * (1.0 - c1) - p.prior_fp_b * c1;
      }
    }
    grad[(size_t)c * p.D + i] = (T)g;
  }
}

// Stand-alone finalize (site-sharded mode: runs after the collective on the raw sums).
template <typename T>
__global__ void finalize_kernel(EvalParams p) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= p.C * p.NQ) return;
  finalize_chain<T>(p, idx / p.NQ, idx % p.NQ, p.sums[idx], false);
}

// Called by every thread of a block after it has published its [bx][c][q] partials: the last block
// of the chain chunk (ticket counter) sums the site splits in split order and writes the outputs.
//
// Tickets of chunk y live at counters[y * kTicketStride + ...]: [0] the chunk's final ticket, [1 + g] group g's.
// With >= kHierMinSplits site splits the sum is taken in TWO levels (measured on B200 at config 2, K1s: one block
// walking 444..592 splits x (chains x quantities) through L2 on its own was 15 of 37 us at C = 1 and 29 of 78 us at
// C = 5 -- every other SM idle): splits are cut into groups of ~sqrt(nsplit) consecutive blocks; the last block of
// a group to arrive adds the group's rows (<= gs independent L2 loads per item, in flight together) IN PLACE into
// the group's first row, then takes the chunk's final ticket; the last group adds the <= 63 group rows and
// finalises.  Group membership and both summation orders are fixed by (grid, block index): deterministic.
constexpr int kTicketStride = 64;
constexpr unsigned kHierMinSplits = 64;

// src[0] + src[stride] + ... (n terms, in index order) with the L2 loads issued 16 at a time: written out because
// `total += __ldcg(..)` in an unrolled loop is compiled into load -> add -> load (measured with %globaltimer stamps,
// K1s at C = 8: 74 dependent loads = 21 us of a 107 us evaluation).
static __device__ __noinline__ double sum_rows(const double* __restrict__ src, size_t stride, unsigned n) {
  double total = 0.0;
  for (unsigned b = 0; b < n; b += 16) {
    double v[16];
#pragma unroll
    for (unsigned u = 0; u < 16; ++u) v[u] = (b + u < n) ? __ldcg(src + (size_t)(b + u) * stride) : 0.0;
#pragma unroll
    for (unsigned u = 0; u < 16; ++u) total += v[u];
  }
  return total;
}

// __noinline__: one copy per translation unit instead of one per kernel instantiation (it runs once per block, and
// inlining its three reduction paths into ~150 kernels was a third of the library's compile time).
template <typename T>
__device__ __noinline__ void finish_block(const EvalParams& p, int c0, int ncb, int* s_is_last,
                                          double* coop_scratch = nullptr) {
  const int tid = threadIdx.x, NQ = p.NQ;
  unsigned int* tickets = p.counters + (size_t)blockIdx.y * kTicketStride;
  const int items = ncb * NQ;
  const unsigned nsplit = gridDim.x;
  if (coop_scratch && items <= (int)blockDim.x && nsplit <= 256) {
    // Few items, one block per SM (K1s; `coop_scratch` = blockDim.x doubles of shared memory): the last block sums
    // with ALL its threads -- thread (g, j) adds rows g, g + G, ... of item j (G = blockDim.x / items; the loads of
    // a thread are independent and in flight together), then thread j adds the G sums in order: one ticket, one or
    // two L2 round trips, a fixed order.
    __threadfence();
    __syncthreads();
    if (tid == 0) *s_is_last = (atomicAdd(&tickets[0], 1u) == nsplit - 1);
    __syncthreads();
    if (!*s_is_last) return;
    __threadfence();
    const int G = (int)blockDim.x / items;
    const size_t rstride = (size_t)p.C * NQ;
    if (tid < G * items) {
      const int g = tid / items, j = tid % items;
      const double* src = p.partial + (size_t)c0 * NQ + j + (size_t)g * rstride;  // rows g, g + G, ...
      const unsigned nrows = (unsigned)g < nsplit ? (nsplit - g + G - 1) / G : 0u;
      coop_scratch[tid] = sum_rows(src, (size_t)G * rstride, nrows);
    }
    __syncthreads();
    if (tid < items) {
      double total = 0.0;
      for (int g0 = 0; g0 < G; g0 += 16) {  // shared-memory reads in flight together, adds in index order
        double v[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) v[u] = (g0 + u < G) ? coop_scratch[(g0 + u) * items + tid] : 0.0;
#pragma unroll
        for (int u = 0; u < 16; ++u) total += v[u];
      }
      const int c = c0 + tid / NQ, q = tid % NQ;
      if (p.allreduce) p.sums[(size_t)c * NQ + q] = total + (q == 0 ? p.cop_const : 0.0);
      else finalize_chain<T>(p, c, q, total, true);
    }
    if (tid == 0) tickets[0] = 0;
    return;
  }
  if (p.hier_reduce && nsplit >= kHierMinSplits) {
    unsigned gs = 16;
    while (gs * gs < nsplit) ++gs;  // <= 63 groups for any grid up to 3969 splits
    const unsigned ngroups = (nsplit + gs - 1) / gs;
    const unsigned g = blockIdx.x / gs, gfirst = g * gs;
    const unsigned gcount = min(gs, nsplit - gfirst);
    const size_t rstride = (size_t)p.C * NQ;
    __threadfence();
    __syncthreads();
    if (tid == 0) *s_is_last = (atomicAdd(&tickets[1 + g], 1u) == gcount - 1);
    __syncthreads();
    if (!*s_is_last) return;
    __threadfence();
    double* row0 = p.partial + (size_t)gfirst * rstride + (size_t)c0 * NQ;
    for (int i = tid; i < items; i += blockDim.x) {
      row0[i] = sum_rows(row0 + i, rstride, gcount);
    }
    if (tid == 0) tickets[1 + g] = 0;  // self-reset for the next launch
    __threadfence();
    __syncthreads();
    if (tid == 0) *s_is_last = (atomicAdd(&tickets[0], 1u) == ngroups - 1);
    __syncthreads();
    if (!*s_is_last) return;
    __threadfence();
    const double* col0 = p.partial + (size_t)c0 * NQ;
    for (int i = tid; i < items; i += blockDim.x) {
      const double total = sum_rows(col0 + i, (size_t)gs * rstride, ngroups);
      const int c = c0 + i / NQ, q = i % NQ;
      if (p.allreduce) p.sums[(size_t)c * NQ + q] = total + (q == 0 ? p.cop_const : 0.0);
      else finalize_chain<T>(p, c, q, total, true);
    }
    if (tid == 0) tickets[0] = 0;
    return;
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const unsigned int ticket = atomicAdd(&tickets[0], 1u);
    *s_is_last = (ticket == gridDim.x - 1);
  }
  __syncthreads();
  if (!*s_is_last) return;
  __threadfence();
  // measured (config 2, us per evaluation, thread-per-item / warp-per-item): C=1 (11 items) 45.8 / 41.6, C=2 63.4 / 63.1,
  // C=4 (44 items) 90.1 / 99.5, C=5 110.6 / 126.5 -> cooperative only up to 32 items
  constexpr int kCoopItems = 32;
  __shared__ double s_tot[kCoopItems];
  if (items <= kCoopItems && p.coop_reduce) {
    // Few items (small chain batches): one WARP per (chain, quantity) -- the lanes sum the site splits b = lane,
    // lane + 32, ... and a fixed xor-butterfly adds the 32 partial sums (a fixed order: deterministic), so the
    // splits' L2 round trips overlap instead of queueing behind one thread; then one THREAD per item adds the
    // priors (fp64 log / lgamma) and writes the outputs.
    const int lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    for (int j = warp; j < items; j += nwarps) {
      double total = 0.0;
      const double* src = p.partial + (size_t)c0 * NQ + j;
      for (unsigned int b = lane; b < gridDim.x; b += 32) total += __ldcg(src + (size_t)b * p.C * NQ);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
      if (lane == 0) s_tot[j] = total;
    }
    __syncthreads();
    if (tid < items) {
      const int c = c0 + tid / NQ, q = tid % NQ;
      if (p.allreduce) p.sums[(size_t)c * NQ + q] = s_tot[tid] + (q == 0 ? p.cop_const : 0.0);
      else finalize_chain<T>(p, c, q, s_tot[tid], true);
    }
  } else {
    for (int i = tid; i < items; i += blockDim.x) {
      const double total = sum_rows(p.partial + (size_t)c0 * NQ + i, (size_t)p.C * NQ, gridDim.x);
      const int c = c0 + i / NQ, q = i % NQ;
      if (p.allreduce) p.sums[(size_t)c * NQ + q] = total + (q == 0 ? p.cop_const : 0.0);
      else finalize_chain<T>(p, c, q, total, true);
    }
  }
  if (tid == 0) tickets[0] = 0;  // self-reset for the next launch
}

// Sums q[0..NQ) of one (warp-tile, chain) over the 32 lanes into the warp's fp64 accumulator row, in
// chunks of <= 32 quantities through the transposed butterfly (fixed order -> deterministic).
template <typename T, int OFF, int NQM>
__device__ __forceinline__ void reduce_into(const T* __restrict__ q, bool valid, int lane, int NQ,
                                            double* __restrict__ acc) {
  if constexpr (OFF < NQM) {
    constexpr int REM = NQM - OFF;
    constexpr int P = REM >= 32 ? 32 : (REM > 16 ? 32 : REM > 8 ? 16 : REM > 4 ? 8 : REM > 2 ? 4 : REM > 1 ? 2 : 1);
    constexpr int SH = 5 - Log2<P>::v;
    T v[P];
#pragma unroll
    for (int i = 0; i < P; ++i) v[i] = (valid && OFF + i < NQ) ? q[OFF + i < NQM ? OFF + i : 0] : T(0);
    const T r = transpose_reduce<T, P>(v, lane);
    const int idx = OFF + (lane >> SH);
    if ((lane & ((1 << SH) - 1)) == 0 && idx < NQ) acc[idx] += (double)r;
    reduce_into<T, OFF + 32, NQM>(q, valid, lane, NQ, acc);
  }
}

// Tried and rejected for small chain batches (measured, config 2): per-lane fp32 running sums [chain][q][thread] in
// shared memory with one butterfly per 32 tiles instead of one per (warp-tile, chain): 11 shared-memory
// read-modify-writes per (tile, chain) cost more than the 16-shuffle transposed butterfly they replace and shorten
// the TMA ring (C = 5: 130 us against 112 us, C = 8: 182 against 154, C = 1 unchanged).
template <typename T, class Model, int MINB>
__global__ void __launch_bounds__(kBlockThreads, MINB) eval_kernel(const __grid_constant__ EvalParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
  const int F = p.L.F;
  const int WS = p.WS, WC = p.WC, NQ = p.NQ, D = p.D, DS = p.DS;
  const uint32_t tile_elems = (uint32_t)WS * F * kWarp;        // one block-tile
  const uint32_t tile_bytes = tile_elems * sizeof(T);
  T* stage0 = reinterpret_cast<T*>(smem_raw + 128);
  T* s_theta = stage0 + (size_t)p.nstage * tile_elems;
  size_t theta_bytes = ((size_t)p.CB * DS * sizeof(T) + 15) & ~size_t(15);
  double* s_acc = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(s_theta) + theta_bytes);
  __shared__ int s_is_last;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ws = warp / WC, wc = warp % WC;
  const int c0 = blockIdx.y * p.CB;
  const int ncb = min(p.CB, p.C - c0);

  // contiguous range of block-tiles for this site split
  const int64_t nbt = p.n_block_tiles;
  const int64_t bt_begin = nbt * blockIdx.x / gridDim.x;
  const int64_t bt_end = nbt * (blockIdx.x + 1) / gridDim.x;
  const int n_it = (int)(bt_end - bt_begin);
  const T* packed = reinterpret_cast<const T*>(p.packed);

  if (tid == 0) {
    for (int s = 0; s < p.nstage; ++s) mbar_init(&bars[s], 1);
    fence_mbar_init();
  }
  for (int i = tid; i < ncb * D; i += kBlockThreads)
    s_theta[(i / D) * DS + (i % D)] = reinterpret_cast<const T*>(p.theta)[(size_t)c0 * D + i];
  for (int i = tid; i < WS * p.CB * NQ; i += kBlockThreads) s_acc[i] = 0.0;
  __syncthreads();
  if constexpr (Model::kDerived > 0) {  // per-chain constants derived from theta, once per block
    for (int ci = tid; ci < ncb; ci += kBlockThreads) Model::derive(p, s_theta + (size_t)ci * DS);
    __syncthreads();
  }

  if (tid == 0) {
    const int pre = min(p.nstage, n_it);
    for (int s = 0; s < pre; ++s) {
      mbar_expect_tx(&bars[s], tile_bytes);
      tma_load_bulk(stage0 + (size_t)s * tile_elems, packed + (size_t)(bt_begin + s) * tile_elems, tile_bytes,
                    &bars[s]);
    }
  }

  for (int it = 0; it < n_it; ++it) {
    const int s = it % p.nstage;
    mbar_wait(&bars[s], (uint32_t)((it / p.nstage) & 1));
    const T* tile = stage0 + (size_t)s * tile_elems + (size_t)ws * F * kWarp;
    const int64_t unit = ((bt_begin + it) * WS + ws) * kWarp + lane;
    const bool valid = unit < p.L.n_units;

    typename Model::Site site;
    Model::load_site(p, tile, lane, site);

    int ci = wc;
    if constexpr (Model::kMultiChain >= 4) {
      if (p.nch >= 4) {
        for (; ci + 3 * WC < ncb; ci += 4 * WC) {
          T q4[4][Model::kNQMax];
          Model::template site_chain_n<4>(p, tile, lane, site, s_theta + (size_t)ci * DS, WC * DS, q4);
#pragma unroll
          for (int c = 0; c < 4; ++c)
            reduce_into<T, 0, Model::kNQMax>(q4[c], valid, lane, NQ, s_acc + ((size_t)ws * p.CB + ci + c * WC) * NQ);
        }
      }
    }
    if constexpr (Model::kMultiChain >= 2) {
      if (p.nch >= 2) {
        for (; ci + WC < ncb; ci += 2 * WC) {
          T q2[2][Model::kNQMax];
          Model::template site_chain_n<2>(p, tile, lane, site, s_theta + (size_t)ci * DS, WC * DS, q2);
#pragma unroll
          for (int c = 0; c < 2; ++c)
            reduce_into<T, 0, Model::kNQMax>(q2[c], valid, lane, NQ, s_acc + ((size_t)ws * p.CB + ci + c * WC) * NQ);
        }
      }
    }
    for (; ci < ncb; ci += WC) {
      T q[Model::kNQMax];
      Model::site_chain(p, tile, lane, site, s_theta + (size_t)ci * DS, q);
      reduce_into<T, 0, Model::kNQMax>(q, valid, lane, NQ, s_acc + ((size_t)ws * p.CB + ci) * NQ);
    }
    __syncthreads();  // every warp is done reading stage s
    if (tid == 0 && it + p.nstage < n_it) {
      mbar_expect_tx(&bars[s], tile_bytes);
      tma_load_bulk(stage0 + (size_t)s * tile_elems, packed + (size_t)(bt_begin + it + p.nstage) * tile_elems,
                    tile_bytes, &bars[s]);
    }
  }
  __syncthreads();

  // block partial: sum the site-groups in fixed order, publish [bx][c][q]
  double* my_partial = p.partial + ((size_t)blockIdx.x * p.C + c0) * NQ;
  for (int i = tid; i < ncb * NQ; i += kBlockThreads) {
    double v = 0.0;
    for (int g = 0; g < WS; ++g) v += s_acc[(size_t)g * p.CB * NQ + i];
    my_partial[i] = v;
  }
  finish_block<T>(p, c0, ncb, &s_is_last);
}

// ------------------------------------------------------------------------------------------
// Per-unit posterior summaries over a batch of draws (the deterministic
// `psi` / `abundance` sites and the pointwise log-likelihood that the reference materialises per draw,
// biolith/models/occu.py:207, utils/predict.py:67-72, evaluation/lppd.py, waic.py -- here streamed:
// nothing of size draws x sites ever exists).  lane = unit, draws staged through shared memory.
//   out[0][u] = mean_n psi_u(theta_n)            (occu_rn: mean lambda_u)
//   out[1][u] = mean_n P(z_u = 1 | y, theta_n)   (occu_rn: mean E[N_u | y, theta_n])
//   out[2][u] = log mean_n exp(l_u(theta_n))     (pointwise marginal lppd)
//   out[3][u] = var_n l_u(theta_n)               (pointwise p_waic)
// ------------------------------------------------------------------------------------------
constexpr int kSummaryDraws = 32;  // draws staged per chunk

template <typename T, class Model>
__global__ void __launch_bounds__(kBlockThreads) summary_kernel(const EvalParams p, float* __restrict__ out) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  T* s_theta = reinterpret_cast<T*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31;
  const int D = p.D, DS = p.DS, N = p.C;
  const int64_t unit = (int64_t)blockIdx.x * kBlockThreads + tid;
  const bool valid = unit < p.L.n_units;
  const int64_t tile_idx = valid ? unit / kWarp : 0;
  const T* tile = reinterpret_cast<const T*>(p.packed) + tile_idx * (int64_t)p.L.F * kWarp;
  typename Model::Site site;
  Model::load_site(p, tile, lane, site);
  double s1 = 0.0, s2 = 0.0, mean = 0.0, m2 = 0.0, mx = -1e300, se = 0.0;
  for (int c0 = 0; c0 < N; c0 += kSummaryDraws) {
    const int nc = min(kSummaryDraws, N - c0);
    __syncthreads();
    for (int i = tid; i < nc * D; i += kBlockThreads)
      s_theta[(i / D) * DS + (i % D)] = reinterpret_cast<const T*>(p.theta)[(size_t)c0 * D + i];
    __syncthreads();
    if constexpr (Model::kDerived > 0) {
      for (int ci = tid; ci < nc; ci += kBlockThreads) Model::derive(p, s_theta + (size_t)ci * DS);
      __syncthreads();
    }
    for (int ci = 0; ci < nc; ++ci) {
      T q[Model::kNQMax];
      T ex[2] = {T(0), T(0)};
      Model::site_chain(p, tile, lane, site, s_theta + (size_t)ci * DS, q, ex);
      const double ell = (double)q[0] + (double)Model::unit_const(p, tile, lane);
      s1 += (double)ex[0];
      s2 += (double)ex[1];
      const double n = (double)(c0 + ci + 1);
      const double d = ell - mean;
      mean += d / n;
      m2 += d * (ell - mean);
      if (ell > mx) { se = se * exp(mx - ell) + 1.0; mx = ell; }
      else se += exp(ell - mx);
    }
  }
  if (valid) {
    const int64_t U = p.L.n_units;
    out[unit] = (float)(s1 / N);
    out[U + unit] = (float)(s2 / N);
    out[2 * U + unit] = (float)(mx + log(se / N));
    out[3 * U + unit] = (float)(N > 1 ? m2 / (N - 1) : 0.0);
  }
}

template <typename T, class Model>
cudaError_t launch_summary(const EvalParams& p, float* out, cudaStream_t st) {
  auto kern = summary_kernel<T, Model>;
  const size_t smem = (size_t)kSummaryDraws * p.DS * sizeof(T) + 128;
  const unsigned blocks = (unsigned)((p.L.n_units + kBlockThreads - 1) / kBlockThreads);
  if (blocks == 0) return cudaSuccess;
  kern<<<blocks, kBlockThreads, smem, st>>>(p, out);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// host-side launch geometry
// ------------------------------------------------------------------------------------------
struct Geometry {
  int CB, n_chunks, WC, WS, nstage, nsplit;
  int64_t n_block_tiles;
  size_t smem_bytes;
};

inline int pow2_floor(int x) { int p = 1; while (p * 2 <= x) p *= 2; return p; }

inline size_t eval_smem_bytes(const Layout& L, int elem, int WS, int nstage, int CB, int D, int DS) {
  size_t tile = (size_t)WS * L.F * kWarp * elem;
  size_t theta = ((size_t)CB * DS * elem + 15) & ~size_t(15);
  size_t acc = (size_t)WS * CB * (1 + D) * sizeof(double);
  return 128 + nstage * tile + theta + acc;
}

inline Geometry plan_geometry(const Layout& L, int elem, int C, int D, int DS, int num_sms, int blocks_per_sm,
                              size_t smem_limit, int wc_override = 0) {
  Geometry g{};
  g.n_chunks = (C + kMaxChainsPerBlock - 1) / kMaxChainsPerBlock;
  g.CB = (C + g.n_chunks - 1) / g.n_chunks;
  // fewest chain-groups whose fp64 accumulators (WS x CB x NQ doubles) still fit beside the tile ring:
  // WC = 1 keeps every warp busy for any CB (a warp walks all chains of the chunk on its own warp-tile)
  // and stages 8 warp-tiles per TMA; larger chunks trade site-groups for accumulator space
  const size_t acc_budget = smem_limit / blocks_per_sm / 3;
  g.WC = 1;
  while (g.WC < kWarpsPerBlock && (size_t)(kWarpsPerBlock / g.WC) * g.CB * (1 + D) * sizeof(double) > acc_budget)
    g.WC *= 2;
  // ... and wide units (many visits / fp64) trade them for a ring of at least two stages
  while (g.WC < kWarpsPerBlock &&
         eval_smem_bytes(L, elem, kWarpsPerBlock / g.WC, 2, g.CB, D, DS) > smem_limit / blocks_per_sm)
    g.WC *= 2;
  if (wc_override > 0) g.WC = wc_override;
  g.WS = kWarpsPerBlock / g.WC;
  g.n_block_tiles = (L.n_tiles + g.WS - 1) / g.WS;
  g.nstage = kMaxStages;
  while (g.nstage > 1 && eval_smem_bytes(L, elem, g.WS, g.nstage, g.CB, D, DS) > smem_limit / blocks_per_sm) --g.nstage;
  g.smem_bytes = eval_smem_bytes(L, elem, g.WS, g.nstage, g.CB, D, DS);
  int64_t want = (int64_t)num_sms * blocks_per_sm / g.n_chunks;
  if (want < 1) want = 1;
  g.nsplit = (int)(want < g.n_block_tiles ? want : g.n_block_tiles);
  if (g.nsplit < 1) g.nsplit = 1;
  return g;
}

}  // namespace bl
