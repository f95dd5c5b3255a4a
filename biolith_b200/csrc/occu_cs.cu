// K8: fused log-marginal + gradient of the continuous-score occupancy model (Rhinehart et al. 2022) --
// two enumerated latents (z per unit, f per visit) and a two-component
// Normal mixture on the classifier score.
//
// Replaces value_and_grad(potential_fn) of biolith/models/occu_cs.py:146-223 (reference): per unit
//   psi = sigmoid(beta0 + X.beta_1:)                         (occu_cs.py:179-188)
//   z ~ Bernoulli(psi~), enumerated                          (occu_cs.py:189-191)
//   f_j ~ Bernoulli(clamp(z sigmoid(nu_j))), enumerated      (occu_cs.py:202-213)
//   s_j ~ Normal((1-f) mu0 + f mu1, (1-f) sigma0 + f sigma1), NaN-masked   (occu_cs.py:215-223)
// With n0_j / n1_j the two Normal log-densities of the score and q the clamped probability of f = 1
// in a z-branch (z = 1: p~_j, z = 0: tiny):
//   L_j(z) = logaddexp(log(1-q) + n0_j, log q + n1_j),   w_j(z) = P(f_j = 1 | s_j, z)
//   l = logaddexp(log psi~ + sum_j m_j L_j(1), log(1-psi~) + sum_j m_j L_j(0)),  r = P(z = 1 | s)
//   dl/deta = r - psi;  dl/dnu_j = r m_j (w_j(1) - p_j);
//   dl/dmu_f = sum_j m_j wbar_fj e_fj / sigma_f,  dl/dlog sigma_f = sum_j m_j wbar_fj (e_fj^2 - 1),
//   e_fj = (s_j - mu_f)/sigma_f,  wbar = r w(1) + (1-r) w(0)  (P(f = 1 | s), mixed over z).
// theta extras (unconstrained, numpyro's biject_to): mu0, x1 = log(mu1 - mu0) (mu1 is left-truncated
// at mu0, occu_cs.py:149), log sigma0, log sigma1; the chain rule mu1 = mu0 + exp(x1) is applied here,
// the priors in engine.cuh:finalize_chain.  Closed form: oracle/occupancy.py:occu_cs_logp_grad.
#include <type_traits>

#include "engine.cuh"

namespace bl {

template <typename T, int KS, int KO, bool STRICT>
struct OccuCsModel {
  using N = Num<T>;
  static constexpr bool kSfu = std::is_same<T, float>::value && !STRICT;
  using M = Mth<T, kSfu>;
  static constexpr bool kGeneric = (KS < 0);
  static constexpr int KSM = kGeneric ? kMaxCov : KS;
  static constexpr int KOM = kGeneric ? kMaxCov : KO;
  static constexpr int kNQMax = kGeneric ? 40 : (1 + KS + 1 + KO + 1 + 4);
  // derived per-chain slots after the D raw parameters: mu1, 1/sigma0, 1/sigma1, exp(x1)
  static constexpr int kDerived = 4;
  static constexpr int kMultiChain = 1;  // chains per pass over a warp-tile (engine.cuh)

  struct Site {
    T x[KSM];
    T nm;  // unmasked visits (data only)
  };
  static __device__ __forceinline__ T unit_const(const EvalParams&, const T*, int) { return T(0); }

  static __device__ __forceinline__ void derive(const EvalParams& p, T* th) {
    const int i0 = p.L.ks + p.L.ko + 2;
    T* d = th + p.D;
    const T e1 = N::exp_(th[i0 + 1]);
    d[0] = th[i0] + e1;
    d[1] = N::exp_(-th[i0 + 2]);
    d[2] = N::exp_(-th[i0 + 3]);
    d[3] = e1;
  }

  static __device__ __forceinline__ void load_site(const EvalParams& p, const T* __restrict__ tile, int lane,
                                                   Site& s) {
    const int ks = kGeneric ? p.L.ks : KS;
#pragma unroll
    for (int k = 0; k < KSM; ++k) s.x[k] = (k < ks) ? tile[k * kWarp + lane] : T(0);
    int nm = 0;
    for (int w = 0; w < p.L.nw; ++w) nm += __popc(N::as_bits(tile[(p.L.off_m + w) * kWarp + lane]));
    s.nm = (T)nm;
  }

  // clamped log q~, log(1-q~), q = sigmoid(nu), and the in-range flag (numpyro clamp_probs)
  static __device__ __forceinline__ void clamped_pair(T nu, T& lq, T& l1q, T& pj, bool& inr) {
    if constexpr (kSfu) {
      const sfu::SoftSig ss = sfu::softsig<true>(nu);
      lq = ss.xc - ss.s;
      l1q = -ss.s;
      pj = ss.p;
      inr = ss.inr;
    } else {
      const LogSig<T> ls = log_sigmoid_pair<T>(nu);
      lq = ls.lp;
      l1q = ls.l1mp;
      pj = ls.p;
      inr = ls.inr;
    }
  }

  static __device__ __forceinline__ void site_chain(const EvalParams& p, const T* __restrict__ tile, int lane,
                                                    const Site& s, const T* __restrict__ th, T* __restrict__ q,
                                                    T* __restrict__ extra = nullptr) {
    const int ks = kGeneric ? p.L.ks : KS;
    const int ko = kGeneric ? p.L.ko : KO;
    const int J = p.L.J;
    const int i0 = ks + ko + 2;
    T eta = th[0];
#pragma unroll
    for (int k = 0; k < KSM; ++k)
      if (k < ks) eta = N::fma_(s.x[k], th[1 + k], eta);
    const T* al = th + ks + 1;
    const T a0 = al[0];
    T a[KOM], ga[KOM];
#pragma unroll
    for (int k = 0; k < KOM; ++k) {
      a[k] = (k < ko) ? al[1 + k] : T(0);
      ga[k] = T(0);
    }
    const T mu0 = th[i0], xs0 = th[i0 + 2], xs1 = th[i0 + 3];
    const T* d = th + p.D;
    const T mu1 = d[0], is0 = d[1], is1 = d[2], e1x = d[3];
    const T cdiff = xs0 - xs1;                               // log sigma0 - log sigma1
    const T dz0 = N::log_tiny() - N::neg_tiny();             // log q~ - log(1-q~) in the z = 0 branch

    T L1 = T(0), L0 = T(0), ga0 = T(0);
    T A_e0 = T(0), A_q0 = T(0), A_e1 = T(0), A_q1 = T(0);    // z = 1 branch sums
    T B_e0 = T(0), B_q0 = T(0), B_e1 = T(0), B_q1 = T(0);    // z = 0 branch sums
    uint32_t mw = 0;
    const T* wrow = tile + p.L.off_w * kWarp + lane;
#pragma unroll 2
    for (int j = 0; j < J; ++j) {
      if ((j & 31) == 0) mw = N::as_bits(tile[(p.L.off_m + (j >> 5)) * kWarp + lane]);
      const T mf = ((mw >> (j & 31)) & 1u) ? T(1) : T(0);
      T w[KOM];
      T nu = a0;
#pragma unroll
      for (int k = 0; k < KOM; ++k) {
        w[k] = (k < ko) ? wrow[(j * ko + k) * kWarp] : T(0);
        nu = N::fma_(w[k], a[k], nu);
      }
      const T sc = tile[(p.L.off_y + j) * kWarp + lane];  // score, 0 where masked
      const T e0 = (sc - mu0) * is0, e1 = (sc - mu1) * is1;
      const T h0 = T(-0.5) * e0 * e0, h1 = T(-0.5) * e1 * e1;
      const T delta = (h1 - h0) + cdiff;                  // n1_j - n0_j
      T lq, l1q, pj;
      bool inr;
      clamped_pair(nu, lq, l1q, pj, inr);
      T sp1, w1, sp0, v1;
      M::softsig((lq - l1q) + delta, sp1, w1);            // z = 1: L = log(1-q~) + n0 + softplus(.)
      M::softsig(dz0 + delta, sp0, v1);                   // z = 0
      L1 = N::fma_(mf, l1q + sp1 + h0, L1);
      L0 = N::fma_(mf, N::neg_tiny() + sp0 + h0, L0);
      const T g = inr ? mf * (w1 - pj) : T(0);
      ga0 += g;
#pragma unroll
      for (int k = 0; k < KOM; ++k)
        if (k < ko) ga[k] = N::fma_(g, w[k], ga[k]);
      const T q0 = N::fma_(e0, e0, T(-1)), q1 = N::fma_(e1, e1, T(-1));
      const T w1m = mf * w1, w0m = mf - w1m, v1m = mf * v1, v0m = mf - v1m;
      A_e0 = N::fma_(w0m, e0, A_e0); A_q0 = N::fma_(w0m, q0, A_q0);
      A_e1 = N::fma_(w1m, e1, A_e1); A_q1 = N::fma_(w1m, q1, A_q1);
      B_e0 = N::fma_(v0m, e0, B_e0); B_q0 = N::fma_(v0m, q0, B_q0);
      B_e1 = N::fma_(v1m, e1, B_e1); B_q1 = N::fma_(v1m, q1, B_q1);
    }

    T lpsi, l1psi, psi;
    bool in_psi;
    clamped_pair(eta, lpsi, l1psi, psi, in_psi);
    const T av = lpsi + L1, bv = l1psi + L0;
    T spd, r;
    M::softsig(av - bv, spd, r);
    // the Normal constants shared by both f states of every unmasked visit: -(log sigma0 + log sqrt(2 pi))
    const T ell = bv + spd - s.nm * (xs0 + T(0.91893853320467274178));
    const T geta = in_psi ? (r - psi) : T(0);
    const T r0 = T(1) - r;
    if (extra) { extra[0] = psi; extra[1] = r; }
    q[0] = ell;
    q[1] = geta;
#pragma unroll
    for (int k = 0; k < KSM; ++k)
      if (k < ks) q[2 + k] = geta * s.x[k];
    q[2 + ks] = r * ga0;
#pragma unroll
    for (int k = 0; k < KOM; ++k)
      if (k < ko) q[3 + ks + k] = r * ga[k];
    const T g_mu0 = (r * A_e0 + r0 * B_e0) * is0;
    const T g_mu1 = (r * A_e1 + r0 * B_e1) * is1;
    q[1 + i0] = g_mu0 + g_mu1;          // mu1 = mu0 + exp(x1)
    q[2 + i0] = g_mu1 * e1x;
    q[3 + i0] = r * A_q0 + r0 * B_q0;
    q[4 + i0] = r * A_q1 + r0 * B_q1;
  }
};

template <typename T, int KS, int KO, bool STRICT>
static cudaError_t launch_cs_one(const EvalParams& p, dim3 grid, size_t smem, cudaStream_t stream, int* occ) {
  auto kern = eval_kernel<T, OccuCsModel<T, KS, KO, STRICT>, 2>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  if (occ) return cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, kern, kBlockThreads, smem);
  kern<<<grid, kBlockThreads, smem, stream>>>(p);
  return cudaGetLastError();
}

template <typename T, bool STRICT>
static cudaError_t dispatch_cs(const EvalParams& p, dim3 grid, size_t smem, cudaStream_t st, int* occ) {
  const int ks = p.L.ks, ko = p.L.ko;
  if (ks == 1 && ko == 1) return launch_cs_one<T, 1, 1, STRICT>(p, grid, smem, st, occ);
  if (ks == 5 && ko == 3) return launch_cs_one<T, 5, 3, STRICT>(p, grid, smem, st, occ);
  return launch_cs_one<T, -1, -1, STRICT>(p, grid, smem, st, occ);
}

cudaError_t launch_occu_cs(const EvalParams& p, int dtype, dim3 grid, size_t smem, cudaStream_t stream, int* occ) {
  if (dtype == BL_F64) return dispatch_cs<double, true>(p, grid, smem, stream, occ);
  return (p.flags & BL_FLAG_STRICT_MATH) ? dispatch_cs<float, true>(p, grid, smem, stream, occ)
                                         : dispatch_cs<float, false>(p, grid, smem, stream, occ);
}

cudaError_t launch_occu_cs_summary(const EvalParams& p, int dtype, float* out, cudaStream_t st) {
  if (dtype == BL_F64) return launch_summary<double, OccuCsModel<double, -1, -1, true>>(p, out, st);
  return launch_summary<float, OccuCsModel<float, -1, -1, true>>(p, out, st);
}

int occu_cs_has_specialisation(int ks, int ko) { return (ks == 1 && ko == 1) || (ks == 5 && ko == 3); }

}  // namespace bl
