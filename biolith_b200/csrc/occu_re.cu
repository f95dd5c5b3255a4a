// K9: occu with site / observation random effects (reference: biolith/models/occu.py:170-173
// site_re_sd / obs_re_sd ~ HalfNormal, :191-196 site_re_occ, site_re_det ~ Normal(0, site_re_sd) per site,
// :215-218 obs_re ~ Normal(0, obs_re_sd) per observation, added to the two linear predictors :199-202, :224-228).
//
// The parameter vector of one chain grows to
//     theta = [ beta (Ks+1) | alpha (Ko+1) | log sd_site | log sd_obs | a (S) | d (S) | o (S*P*J) ]
// (oracle/occupancy.py:_split_re; the sd's are sampled in log space like numpyro's biject_to(positive)), so the
// gradient has two kinds of entries: the usual REDUCTIONS over sites (beta, alpha, the log sd's) and ELEMENTWISE
// outputs, one per random effect:  dl/da_s = sum_p (r - psi) - a_s / sd^2,  dl/dd_s = sum_pj dl/dnu - d_s / sd^2,
// dl/do_spj = dl/dnu_spj - o_spj / sd_o^2.  Mapping: lane = SITE (a thread walks the periods of its site, so the
// per-site sums need no atomics and every elementwise gradient is written once, coalesced); grid = (site blocks,
// chains); the reductions use the engine's fp64 block partials + ticketed last-block sum (deterministic).
// fp32 (libm-accurate forms) and fp64; any Ks, Ko <= 16; no false-positive extras.
#include "engine.cuh"
#include "handle.h"

namespace bl {

struct ReLayout {
  int site_re, obs_re;
  int n_small;      // Ks + Ko + 2 + number of sd parameters
  int off_sd_site, off_sd_obs;
  int64_t off_a, off_d, off_o, D;
  int64_t S;
};

inline ReLayout make_re_layout(const Layout& L, int64_t S, uint32_t flags) {
  ReLayout r{};
  r.site_re = (flags & BL_FLAG_SITE_RE) != 0;
  r.obs_re = (flags & BL_FLAG_OBS_RE) != 0;
  r.S = S;
  int i = L.ks + L.ko + 2;
  r.off_sd_site = r.site_re ? i++ : -1;
  r.off_sd_obs = r.obs_re ? i++ : -1;
  r.n_small = i;
  int64_t o = i;
  r.off_a = o; if (r.site_re) o += S;
  r.off_d = o; if (r.site_re) o += S;
  r.off_o = o; if (r.obs_re) o += S * L.P * L.J;
  r.D = o;
  return r;
}

int64_t occu_re_theta_dim(const Layout& L, int64_t S, uint32_t flags) { return make_re_layout(L, S, flags).D; }
int occu_re_n_small(const Layout& L, int64_t S, uint32_t flags) { return make_re_layout(L, S, flags).n_small; }

template <typename T>
__global__ void __launch_bounds__(kBlockThreads) occu_re_kernel(const EvalParams p, const ReLayout R,
                                                                 double sd_scale_site, double sd_scale_obs) {
  using N = Num<T>;
  const int c = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ks = p.L.ks, ko = p.L.ko, J = p.L.J, P = p.L.P, F = p.L.F;
  const int NQs = 1 + R.n_small;
  const T* theta = reinterpret_cast<const T*>(p.theta) + (size_t)c * R.D;
  T* grad = reinterpret_cast<T*>(p.grad) + (size_t)c * R.D;
  const T* packed = reinterpret_cast<const T*>(p.packed);
  __shared__ double s_part[kWarpsPerBlock][2 * kMaxCov + 8];
  __shared__ int s_is_last;

  const T sd_s = R.site_re ? N::exp_(theta[R.off_sd_site]) : T(1);
  const T sd_o = R.obs_re ? N::exp_(theta[R.off_sd_obs]) : T(1);
  const T iv_s = T(1) / (sd_s * sd_s), iv_o = T(1) / (sd_o * sd_o);
  const T h2pi = T(0.91893853320467274178);
  const T log_tiny = N::log_tiny();

  // per-thread fp64 partial sums of the reductions: [0] log-density, [1..] beta, alpha, log sd's
  double acc[2 * kMaxCov + 8];
  for (int i = 0; i < NQs; ++i) acc[i] = 0.0;

  for (int64_t s = (int64_t)blockIdx.x * kBlockThreads + tid; s < R.S; s += (int64_t)gridDim.x * kBlockThreads) {
    const T a_s = R.site_re ? theta[R.off_a + s] : T(0);
    const T d_s = R.site_re ? theta[R.off_d + s] : T(0);
    T ga_site = T(0), gd_site = T(0);
    for (int pp = 0; pp < P; ++pp) {
      const int64_t u = s * P + pp;
      const T* base = packed + (u / kWarp) * (int64_t)F * kWarp + (u % kWarp);
      T eta = theta[0];
      for (int k = 0; k < ks; ++k) eta = N::fma_(base[k * kWarp], theta[1 + k], eta);
      eta += a_s;
      const LogSig<T> ps = log_sigmoid_pair<T>(eta);
      T L1 = T(0);
      int n1 = 0, n0 = 0;
      // pass 1: the z = 1 branch
      for (int j = 0; j < J; ++j) {
        const uint32_t mw = N::as_bits(base[(p.L.off_m + (j >> 5)) * kWarp]);
        if (!((mw >> (j & 31)) & 1u)) continue;
        const uint32_t yw = N::as_bits(base[(p.L.off_y + (j >> 5)) * kWarp]);
        const bool y = (yw >> (j & 31)) & 1u;
        T nu = theta[ks + 1];
        for (int k = 0; k < ko; ++k) nu = N::fma_(base[(p.L.off_w + j * ko + k) * kWarp], theta[ks + 2 + k], nu);
        nu += d_s;
        if (R.obs_re) nu += theta[R.off_o + u * J + j];
        const LogSig<T> pj = log_sigmoid_pair<T>(nu);
        L1 += y ? pj.lp : pj.l1mp;
        n1 += y; n0 += !y;
      }
      const T L0 = (T)n1 * log_tiny + (T)n0 * N::neg_tiny();
      const T av = ps.lp + L1, bv = ps.l1mp + L0;
      const T mx = N::max_(av, bv);
      const T ell = mx + N::log1p_(N::exp_(-N::abs_(av - bv)));
      const T dd = av - bv;
      const T td = N::exp_(-N::abs_(dd));
      const T r = dd >= T(0) ? T(1) / (T(1) + td) : td / (T(1) + td);  // P(z = 1 | y)
      const T d_eta = ps.inr ? r - ps.p : T(0);
      acc[0] += (double)ell;
      acc[1] += (double)d_eta;
      for (int k = 0; k < ks; ++k) acc[2 + k] += (double)(d_eta * base[k * kWarp]);
      ga_site += d_eta;
      // pass 2: dl/dnu_j = r m (y - p) [in range]
      for (int j = 0; j < J; ++j) {
        const uint32_t mw = N::as_bits(base[(p.L.off_m + (j >> 5)) * kWarp]);
        const bool m = (mw >> (j & 31)) & 1u;
        T d_nu = T(0);
        if (m) {
          const uint32_t yw = N::as_bits(base[(p.L.off_y + (j >> 5)) * kWarp]);
          const bool y = (yw >> (j & 31)) & 1u;
          T nu = theta[ks + 1];
          for (int k = 0; k < ko; ++k) nu = N::fma_(base[(p.L.off_w + j * ko + k) * kWarp], theta[ks + 2 + k], nu);
          nu += d_s;
          if (R.obs_re) nu += theta[R.off_o + u * J + j];
          const LogSig<T> pj = log_sigmoid_pair<T>(nu);
          d_nu = pj.inr ? r * ((y ? T(1) : T(0)) - pj.p) : T(0);
          acc[2 + ks] += (double)d_nu;
          for (int k = 0; k < ko; ++k) acc[3 + ks + k] += (double)(d_nu * base[(p.L.off_w + j * ko + k) * kWarp]);
          gd_site += d_nu;
        }
        if (R.obs_re) {
          const T o = theta[R.off_o + u * J + j];  // every observation slot carries a random effect, masked or not
          grad[R.off_o + u * J + j] = d_nu - o * iv_o;
          acc[0] += (double)(-T(0.5) * o * o * iv_o - N::log_(sd_o) - h2pi);
          acc[1 + R.off_sd_obs] += (double)(o * o * iv_o - T(1));
        }
      }
    }
    if (R.site_re) {
      grad[R.off_a + s] = ga_site - a_s * iv_s;
      grad[R.off_d + s] = gd_site - d_s * iv_s;
      acc[0] += (double)(-T(0.5) * (a_s * a_s + d_s * d_s) * iv_s - T(2) * N::log_(sd_s) - T(2) * h2pi);
      acc[1 + R.off_sd_site] += (double)((a_s * a_s + d_s * d_s) * iv_s - T(2));
    }
  }
  // block reduction (fixed order): lanes by xor-butterfly, warps through shared memory
  for (int i = 0; i < NQs; ++i) {
    double v = acc[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) s_part[warp][i] = v;
  }
  __syncthreads();
  if (tid < NQs) {
    double v = 0.0;
    for (int w = 0; w < kWarpsPerBlock; ++w) v += s_part[w][tid];
    p.partial[((size_t)blockIdx.x * p.C + c) * NQs + tid] = v;
  }
  // ticket: the last block of this chain sums the site blocks in order and adds the priors
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const unsigned int ticket = atomicAdd(&p.counters[c], 1u);
    s_is_last = (ticket == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_is_last) return;
  __threadfence();
  if (tid < NQs) {
    double total = 0.0;
    for (unsigned int b = 0; b < gridDim.x; ++b) total += __ldcg(p.partial + ((size_t)b * p.C + c) * NQs + tid);
    const bool prior = (p.flags & BL_FLAG_PRIOR) != 0;
    const int KB = ks + 1, KA = ko + 1;
    if (tid == 0) {
      double lp = total;
      if (prior) {
        const double hp = 0.91893853320467274178;
        for (int i = 0; i < KB; ++i) {
          const double z = ((double)theta[i] - p.prior_beta_loc) / p.prior_beta_scale;
          lp += -0.5 * z * z - log(p.prior_beta_scale) - hp;
        }
        for (int i = 0; i < KA; ++i) {
          const double z = ((double)theta[KB + i] - p.prior_alpha_loc) / p.prior_alpha_scale;
          lp += -0.5 * z * z - log(p.prior_alpha_scale) - hp;
        }
        // HalfNormal(scale) on sd = exp(x), + log|d sd / dx| = x
        if (R.site_re) {
          const double x = (double)theta[R.off_sd_site], sd = exp(x) / sd_scale_site;
          lp += log(2.0) - 0.5 * sd * sd - log(sd_scale_site) - hp + x;
        }
        if (R.obs_re) {
          const double x = (double)theta[R.off_sd_obs], sd = exp(x) / sd_scale_obs;
          lp += log(2.0) - 0.5 * sd * sd - log(sd_scale_obs) - hp + x;
        }
      }
      reinterpret_cast<T*>(p.logp)[c] = (T)lp;
      if (p.logp64) p.logp64[c] = lp;
    } else {
      const int i = tid - 1;
      double g = total;
      if (prior) {
        const double x = (double)theta[i];
        if (i < KB) g -= (x - p.prior_beta_loc) / (p.prior_beta_scale * p.prior_beta_scale);
        else if (i < KB + KA) g -= (x - p.prior_alpha_loc) / (p.prior_alpha_scale * p.prior_alpha_scale);
        else {
          const double sc = (i == R.off_sd_site) ? sd_scale_site : sd_scale_obs;
          const double sd = exp(x) / sc;
          g += 1.0 - sd * sd;
        }
      }
      grad[i] = (T)g;
    }
  }
  if (tid == 0) p.counters[c] = 0;
}

cudaError_t launch_occu_re(const EvalParams& p, int dtype, int64_t S, int grid_x, double sd_scale_site,
                           double sd_scale_obs, cudaStream_t st) {
  const ReLayout R = make_re_layout(p.L, S, p.flags);
  const dim3 grid(grid_x, p.C);
  if (dtype == BL_F32) occu_re_kernel<float><<<grid, kBlockThreads, 0, st>>>(p, R, sd_scale_site, sd_scale_obs);
  else occu_re_kernel<double><<<grid, kBlockThreads, 0, st>>>(p, R, sd_scale_site, sd_scale_obs);
  return cudaGetLastError();
}

}  // namespace bl
