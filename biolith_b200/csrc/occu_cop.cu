// K3: fused log-marginal + gradient of the count-detection occupancy model (Pautrel et al. 2024).
//
// Replaces value_and_grad(potential_fn) of biolith/models/occu_cop.py:197-255 (reference): per unit
//   psi = sigmoid(beta0 + X.beta_1:)                                           (occu_cop.py:215-225)
//   mu_j = exp(alpha0 + W_j.alpha_1:),  rate_zj = T_j (z mu_j + (1-z) u + c)    (occu_cop.py:236-255)
//   L_z = sum_j m_j Poisson(rate_zj).log_prob(y_j);  l = logaddexp(log psi~ + L_1, log1p(-psi~) + L_0)
// The data-only part  sum_j m_j (y_j log T_j - lgamma(y_j + 1))  is common to both branches and is
// added once per evaluation as EvalParams::cop_const (computed in fp64 at pack time).  Masked visits
// are packed as (y, T) = (0, 0) and contribute exactly zero.  c = rate_fp_constant, u =
// rate_fp_unoccupied enter through their logs (unconstrained ExpTransform space).
// Closed form: oracle/occupancy.py:occu_cop_logp_grad.
#include <type_traits>

#include "engine.cuh"

namespace bl {

template <typename T, int KS, int KO, bool STRICT>
struct OccuCopModel {
  using N = Num<T>;
  static constexpr bool kSfu = std::is_same<T, float>::value && !STRICT;
  using M = Mth<T, kSfu>;
  static constexpr bool kGeneric = (KS < 0);
  static constexpr int KSM = kGeneric ? kMaxCov : KS;
  static constexpr int KOM = kGeneric ? kMaxCov : KO;
  static constexpr int kNQMax = 40;   // runtime NQ loop in the engine
  static constexpr int kDerived = 6;  // c, u, rho0, log rho0, 1/rho0, -

  struct Site {
    T x[KSM];
    T sy, st;
  };

  static __device__ __forceinline__ void derive(const EvalParams& p, T* th) {
    T* d = th + p.D;
    int i = p.L.ks + p.L.ko + 2;
    const T c = (p.flags & BL_FLAG_FP_CONSTANT) ? N::exp_(th[i++]) : T(0);
    const T u = (p.flags & BL_FLAG_FP_UNOCCUPIED) ? N::exp_(th[i++]) : T(0);
    const T rho0 = u + c;
    d[0] = c;
    d[1] = u;
    d[2] = rho0;
    d[3] = rho0 > T(0) ? N::log_(rho0) : -N::inf();
    d[4] = rho0 > T(0) ? T(1) / rho0 : T(0);
    d[5] = T(0);
  }

  static __device__ __forceinline__ void load_site(const EvalParams& p, const T* __restrict__ tile, int lane,
                                                   Site& s) {
    const int ks = kGeneric ? p.L.ks : KS;
#pragma unroll
    for (int k = 0; k < KSM; ++k) s.x[k] = (k < ks) ? tile[k * kWarp + lane] : T(0);
    s.sy = tile[p.L.off_sy * kWarp + lane];
    s.st = tile[(p.L.off_sy + 1) * kWarp + lane];
  }

  static __device__ __forceinline__ void site_chain(const EvalParams& p, const T* __restrict__ tile, int lane,
                                                    const Site& s, const T* __restrict__ th, T* __restrict__ q) {
    const int ks = kGeneric ? p.L.ks : KS;
    const int ko = kGeneric ? p.L.ko : KO;
    const int J = p.L.J;
    const T* d = th + p.D;
    const T c = d[0], u = d[1], rho0 = d[2];
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
    T L1 = T(0), ga0 = T(0), s1 = T(0);
    const T* wrow = tile + p.L.off_w * kWarp + lane;
    const T* yrow = tile + p.L.off_y * kWarp + lane;
    const T* trow = tile + p.L.off_t * kWarp + lane;
#pragma unroll 2
    for (int j = 0; j < J; ++j) {
      T w[KOM];
      T nu = a0;
#pragma unroll
      for (int k = 0; k < KOM; ++k) {
        w[k] = (k < ko) ? wrow[(j * ko + k) * kWarp] : T(0);
        nu = N::fma_(w[k], a[k], nu);
      }
      const T y = yrow[j * kWarp], Tj = trow[j * kWarp];
      const T mu = M::exp_(nu);
      const T rho1 = mu + c;
      const bool ypos = y > T(0);
      const T t1 = (ypos ? y * M::log_(rho1) : T(0)) - Tj * rho1;
      const T d1 = (ypos ? y * M::rcp_(rho1) : T(0)) - Tj;  // dt1/drho1
      L1 += t1;
      s1 += d1;
      const T g = d1 * mu;  // drho1/dnu = mu
      ga0 += g;
#pragma unroll
      for (int k = 0; k < KOM; ++k)
        if (k < ko) ga[k] = N::fma_(g, w[k], ga[k]);
    }
    // z = 0 branch from the per-unit data sums
    const T L0 = (s.sy > T(0) ? s.sy * d[3] : T(0)) - s.st * rho0;
    const T d0 = (s.sy > T(0) ? s.sy * d[4] : T(0)) - s.st;  // dL0/drho0
    T psi, lpsi, l1psi;
    bool in_psi;
    if constexpr (kSfu) {
      const sfu::SoftSig se = sfu::softsig<true>(eta);
      psi = se.p; lpsi = se.xc - se.s; l1psi = -se.s; in_psi = se.inr;
    } else {
      const LogSig<T> se = log_sigmoid_pair<T>(eta);
      psi = se.p; lpsi = se.lp; l1psi = se.l1mp; in_psi = se.inr;
    }
    const T av = lpsi + L1;
    const T bv = l1psi + L0;
    T r, ell;
    if (bv == -N::inf()) {  // Poisson(0) saw a count: the unit is occupied with certainty
      r = T(1);
      ell = av;
    } else {
      const T dd = av - bv;
      const T td = M::exp_(-N::abs_(dd));
      const T inv = M::rcp_(T(1) + td);
      r = (dd >= T(0)) ? inv : td * inv;
      ell = N::max_(av, bv) + (kSfu ? M::log_(T(1) + td) : N::log1p_(td));
    }
    const T geta = in_psi ? (r - psi) : T(0);
    q[0] = ell;
    q[1] = geta;
#pragma unroll
    for (int k = 0; k < KSM; ++k)
      if (k < ks) q[2 + k] = geta * s.x[k];
    q[2 + ks] = r * ga0;
#pragma unroll
    for (int k = 0; k < KOM; ++k)
      if (k < ko) q[3 + ks + k] = r * ga[k];
    int i = 3 + ks + ko;
    const T w0 = (r < T(1)) ? (T(1) - r) * d0 : T(0);
    if (p.flags & BL_FLAG_FP_CONSTANT) q[i++] = (r * s1 + w0) * c;  // x = log c: dc/dx = c
    if (p.flags & BL_FLAG_FP_UNOCCUPIED) q[i++] = w0 * u;
  }
};

template <typename T, int KS, int KO, bool STRICT>
static cudaError_t launch_cop_one(const EvalParams& p, dim3 grid, size_t smem, cudaStream_t stream, int* occ) {
  auto kern = eval_kernel<T, OccuCopModel<T, KS, KO, STRICT>, 2>;
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

cudaError_t launch_occu_cop(const EvalParams& p, int dtype, dim3 grid, size_t smem, cudaStream_t stream, int* occ) {
  const bool strict = (p.flags & BL_FLAG_STRICT_MATH) != 0;
  const bool s53 = p.L.ks == 5 && p.L.ko == 3;
  if (dtype == BL_F64)
    return s53 ? launch_cop_one<double, 5, 3, true>(p, grid, smem, stream, occ)
               : launch_cop_one<double, -1, -1, true>(p, grid, smem, stream, occ);
  if (strict)
    return s53 ? launch_cop_one<float, 5, 3, true>(p, grid, smem, stream, occ)
               : launch_cop_one<float, -1, -1, true>(p, grid, smem, stream, occ);
  return s53 ? launch_cop_one<float, 5, 3, false>(p, grid, smem, stream, occ)
             : launch_cop_one<float, -1, -1, false>(p, grid, smem, stream, occ);
}

int occu_cop_derived_slots(uint32_t) { return 6; }

}  // namespace bl
