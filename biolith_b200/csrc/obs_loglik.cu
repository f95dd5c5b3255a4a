// Per-OBSERVATION pointwise log-likelihood over a batch of posterior draws, streamed.
//
// Reference: biolith/evaluation/lppd.py:51-61 and waic.py:60-82 reduce a (draws, n_species, n_sites, n_periods,
// n_replicates) array of per-observation log-likelihoods -- numpyro's log_likelihood of site "y" given a z drawn
// per posterior draw (utils/predict.py:67-72) -- to  lppd = sum_obs log mean_draws exp(ll)  and
// p_waic = sum_obs var_draws(ll).  The reference also ships the closed form of the same quantity with z integrated
// out per draw, `log_likelihood_manual` (evaluation/log_likelihood.py:55-98):
//     ll_nij = y_ij log clip(psi_ni p_nij, e, 1-e) + (1 - y_ij) log clip(1 - psi_ni p_nij, e, 1-e),   e = 1e-10,
// (mean_z exp(ll | z) = psi p for y = 1 and 1 - psi p for y = 0, so the two agree in expectation; the reference tests
// them against each other at rtol 1e-1, lppd.py:109-123).  This kernel evaluates the closed form: it is deterministic
// and its inputs are exactly the deterministic sites psi / prob_detection the reference registers (occu.py:207,221).
//
// lane = unit; visits outer, draws inner (theta staged through shared memory): per observation an online
// log-sum-exp and a Welford variance -- nothing of size draws x observations is ever formed.  occu without
// false-positive extras; fp32 / fp64 arithmetic as the dataset.
#include "engine.cuh"
#include "handle.h"

namespace bl {

constexpr int kOllDraws = 64;  // draws staged per chunk

template <typename T>
__global__ void __launch_bounds__(kBlockThreads) obs_loglik_kernel(const EvalParams p, int n_draws,
                                                                   float* __restrict__ lppd,
                                                                   float* __restrict__ var) {
  using N = Num<T>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* s_theta = reinterpret_cast<T*>(smem_raw);
  const int tid = threadIdx.x;
  const int ks = p.L.ks, ko = p.L.ko, J = p.L.J, F = p.L.F, D = p.D;
  const int64_t u = (int64_t)blockIdx.x * kBlockThreads + tid;
  const bool valid = u < p.L.n_units;
  const int64_t uu = valid ? u : 0;
  const T* base = reinterpret_cast<const T*>(p.packed) + (uu / kWarp) * (int64_t)F * kWarp + (uu % kWarp);
  const T eps = T(1e-10);
  T x[kMaxCov];
  for (int k = 0; k < ks; ++k) x[k] = base[k * kWarp];
  for (int j = 0; j < J; ++j) {
    const uint32_t mw = N::as_bits(base[(p.L.off_m + (j >> 5)) * kWarp]);
    const uint32_t yw = N::as_bits(base[(p.L.off_y + (j >> 5)) * kWarp]);
    const bool m = (mw >> (j & 31)) & 1u, y = (yw >> (j & 31)) & 1u;
    T w[kMaxCov];
    for (int k = 0; k < ko; ++k) w[k] = base[(p.L.off_w + j * ko + k) * kWarp];
    double mx = -1e300, se = 0.0, mean = 0.0, m2 = 0.0;
    for (int c0 = 0; c0 < n_draws; c0 += kOllDraws) {
      const int nc = min(kOllDraws, n_draws - c0);
      __syncthreads();
      for (int i = tid; i < nc * D; i += kBlockThreads)
        s_theta[i] = reinterpret_cast<const T*>(p.theta)[(size_t)c0 * D + i];
      __syncthreads();
      for (int ci = 0; ci < nc; ++ci) {
        const T* th = s_theta + (size_t)ci * D;
        T eta = th[0];
        for (int k = 0; k < ks; ++k) eta = N::fma_(x[k], th[1 + k], eta);
        T nu = th[ks + 1];
        for (int k = 0; k < ko; ++k) nu = N::fma_(w[k], th[ks + 2 + k], nu);
        const T psi = T(1) / (T(1) + N::exp_(-eta)), pd = T(1) / (T(1) + N::exp_(-nu));
        const T q = pd * psi;
        const T ll = y ? N::log_(N::min_(N::max_(q, eps), T(1) - eps))
                       : N::log_(N::min_(N::max_(T(1) - q, eps), T(1) - eps));
        const double l = (double)ll, n = (double)(c0 + ci + 1);
        const double d = l - mean;
        mean += d / n;
        m2 += d * (l - mean);
        if (l > mx) { se = se * exp(mx - l) + 1.0; mx = l; }
        else se += exp(l - mx);
      }
    }
    if (valid) {
      const float nanv = __int_as_float(0x7fc00000);
      lppd[u * J + j] = m ? (float)(mx + log(se / n_draws)) : nanv;
      if (var) var[u * J + j] = m ? (float)(n_draws > 1 ? m2 / (n_draws - 1) : 0.0) : nanv;
    }
  }
}

cudaError_t launch_obs_loglik(const EvalParams& p, int dtype, int n_draws, float* lppd, float* var, cudaStream_t st) {
  const unsigned blocks = (unsigned)((p.L.n_units + kBlockThreads - 1) / kBlockThreads);
  if (blocks == 0) return cudaSuccess;
  const size_t es = dtype == BL_F32 ? 4 : 8;
  const size_t smem = (size_t)kOllDraws * p.D * es + 16;
  if (dtype == BL_F32) obs_loglik_kernel<float><<<blocks, kBlockThreads, smem, st>>>(p, n_draws, lppd, var);
  else obs_loglik_kernel<double><<<blocks, kBlockThreads, smem, st>>>(p, n_draws, lppd, var);
  return cudaGetLastError();
}

}  // namespace bl
