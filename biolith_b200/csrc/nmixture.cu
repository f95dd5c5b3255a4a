// K7: fused log-marginal + gradient of the N-mixture model (Royle 2004) -- SURVEY.md section 8 row f4,
// the closest sibling of occu_rn.
//
// Replaces value_and_grad(potential_fn) of biolith/models/nmixture.py:173-220 (reference): per unit
//   lambda = exp(beta0 + X.beta_1:)                                             (nmixture.py:173-180)
//   N in {max_j y_j .. K} with the truncated, NOT renormalised Poisson(lambda) weights: the reference
//     masks logits below the largest count, adds numpyro.factor(logsumexp(logits)) and samples
//     Categorical(logits) -> the normaliser cancels                             (nmixture.py:181-193)
//   y_j ~ Binomial(N, p_j), p_j = sigmoid(alpha0 + W_j.alpha_1:), NaN-masked     (nmixture.py:203-220)
// With u_j = log(1-p_j) the enumerated sum collapses to
//   A_k = k (eta + sum_j m_j u_j) - lambda - lgamma(k+1) + c_k,   c_k = sum_j m_j log C(k, y_j)
//   l   = sum_j m_j y_j nu_j + logsumexp_k A_k
//   dl/deta = E_w[k] - lambda,   dl/dnu_j = m_j (y_j - p_j E_w[k])
// c_k is data-only: each thread (lane = unit) builds its (K+1)-column once per staged tile in shared
// memory and reuses it for every chain of the block; the data-only parts of the alpha gradient
// (sum m y, sum m y W) are packed per unit.  Closed form: oracle/occupancy.py:nmixture_logp_grad.
#include <cmath>
#include <type_traits>

#include "engine.cuh"

namespace bl {

constexpr int kNmixMaxAbundance = 1023;
__device__ double d_lgf64[2 * kNmixMaxAbundance + 2];  // lgamma(k + 1), global (divergent lookups)
__device__ float d_lgf32[2 * kNmixMaxAbundance + 2];

template <typename T> struct LgfTable;
template <> struct LgfTable<float> { static __device__ __forceinline__ float at(int k) { return d_lgf32[k]; } };
template <> struct LgfTable<double> { static __device__ __forceinline__ double at(int k) { return d_lgf64[k]; } };

template <typename T, int KS, int KO, bool STRICT>
struct NmixModel {
  using N = Num<T>;
  static constexpr bool kSfu = std::is_same<T, float>::value && !STRICT;
  using M = Mth<T, kSfu>;
  static constexpr bool kGeneric = (KS < 0);
  static constexpr int KSM = kGeneric ? kMaxCov : KS;
  static constexpr int KOM = kGeneric ? kMaxCov : KO;
  static constexpr int kNQMax = 40;  // runtime NQ loop in the engine
  static constexpr int kDerived = 0;
  static constexpr int kMultiChain = 1;  // chains per pass over a warp-tile (engine.cuh)

  struct Site {
    T x[KSM];
    T y0;
    int kmin;
  };
  static __device__ __forceinline__ T unit_const(const EvalParams&, const T*, int) { return T(0); }
  static __device__ __forceinline__ void derive(const EvalParams&, T*) {}

  static __device__ __forceinline__ T* column(const EvalParams& p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    if (p.rn_scratch_global)
      return reinterpret_cast<T*>(p.rn_scratch_global) +
             ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * kBlockThreads + threadIdx.x;
    return reinterpret_cast<T*>(smem_raw + p.rn_scratch_off) + threadIdx.x;
  }
  static __device__ __forceinline__ size_t stride(const EvalParams& p) {
    return p.rn_scratch_global ? (size_t)gridDim.x * gridDim.y * kBlockThreads : (size_t)kBlockThreads;
  }

  // once per (tile, thread): covariates + the data-only column c_k = sum_j m_j log C(k, y_j) - lgamma(k+1)
  static __device__ __forceinline__ void load_site(const EvalParams& p, const T* __restrict__ tile, int lane,
                                                   Site& s) {
    const int ks = kGeneric ? p.L.ks : KS;
#pragma unroll
    for (int k = 0; k < KSM; ++k) s.x[k] = (k < ks) ? tile[k * kWarp + lane] : T(0);
    s.kmin = (int)tile[p.L.off_sy * kWarp + lane];
    s.y0 = tile[(p.L.off_sy + 1) * kWarp + lane];
    T* Cc = column(p);
    const size_t ST = stride(p);
    const int K = p.K, J = p.L.J;
    int nm = 0;
    T sum_lgy = T(0);
    uint32_t mw = 0;
    for (int j = 0; j < J; ++j) {
      if ((j & 31) == 0) mw = N::as_bits(tile[(p.L.off_m + (j >> 5)) * kWarp + lane]);
      if ((mw >> (j & 31)) & 1u) {
        ++nm;
        sum_lgy += LgfTable<T>::at((int)tile[(p.L.off_y + j) * kWarp + lane]);
      }
    }
    for (int k = 0; k <= K; ++k) {
      T c = -N::inf();
      if (k >= s.kmin) {
        c = (T)(nm - 1) * LgfTable<T>::at(k) - sum_lgy;
        for (int j = 0; j < J; ++j) {
          if ((j & 31) == 0) mw = N::as_bits(tile[(p.L.off_m + (j >> 5)) * kWarp + lane]);
          if ((mw >> (j & 31)) & 1u) c -= LgfTable<T>::at(k - (int)tile[(p.L.off_y + j) * kWarp + lane]);
        }
      }
      Cc[(size_t)k * ST] = c;
    }
  }

  static __device__ __forceinline__ void site_chain(const EvalParams& p, const T* __restrict__ tile, int lane,
                                                    const Site& s, const T* __restrict__ th, T* __restrict__ q,
                                                    T* __restrict__ extra = nullptr) {
    const int ks = kGeneric ? p.L.ks : KS;
    const int ko = kGeneric ? p.L.ko : KO;
    const int J = p.L.J, K = p.K;
    const T* Cc = column(p);
    const size_t ST = stride(p);
    T eta = th[0];
#pragma unroll
    for (int k = 0; k < KSM; ++k)
      if (k < ks) eta = N::fma_(s.x[k], th[1 + k], eta);
    const T lam = M::exp_(eta);
    const T* al = th + ks + 1;
    const T a0 = al[0];
    T a[KOM], pw[KOM];
#pragma unroll
    for (int k = 0; k < KOM; ++k) {
      a[k] = (k < ko) ? al[1 + k] : T(0);
      pw[k] = T(0);
    }
    // visits: U = sum m log(1-p), V = sum m y nu, P0 = sum m p, Pw_k = sum m p W_k
    T U = T(0), V = T(0), P0 = T(0);
    uint32_t mw = 0;
    const T* wrow = tile + p.L.off_w * kWarp + lane;
#pragma unroll 2
    for (int j = 0; j < J; ++j) {
      if ((j & 31) == 0) mw = N::as_bits(tile[(p.L.off_m + (j >> 5)) * kWarp + lane]);
      const bool m = (mw >> (j & 31)) & 1u;
      T w[KOM];
      T nu = a0;
#pragma unroll
      for (int k = 0; k < KOM; ++k) {
        w[k] = (k < ko) ? wrow[(j * ko + k) * kWarp] : T(0);
        nu = N::fma_(w[k], a[k], nu);
      }
      T sp, pj;
      M::softsig(nu, sp, pj);
      const T y = tile[(p.L.off_y + j) * kWarp + lane];  // 0 where masked
      pj = m ? pj : T(0);
      U -= m ? sp : T(0);
      V = N::fma_(y, nu, V);
      P0 += pj;
#pragma unroll
      for (int k = 0; k < KOM; ++k)
        if (k < ko) pw[k] = N::fma_(pj, w[k], pw[k]);
    }
    // online log-sum-exp over the abundance states k >= kmin
    const T slope = eta + U;
    T Mx = -N::inf(), Z = T(0), E = T(0);
    T kf = (T)s.kmin;
    for (int k = s.kmin; k <= K; ++k) {
      const T ak = N::fma_(kf, slope, Cc[(size_t)k * ST]);
      const T Mn = N::max_(Mx, ak);
      const T sc = M::exp_(Mx - Mn);  // exp(-inf) = 0 on the first state
      const T e = M::exp_(ak - Mn);
      Z = N::fma_(Z, sc, e);
      E = N::fma_(E, sc, kf * e);
      Mx = Mn;
      kf += T(1);
    }
    const T Ek = E * M::rcp_(Z);
    const T ell = V + (Mx + M::log_(Z)) - lam;
    const T geta = Ek - lam;
    if (extra) { extra[0] = lam; extra[1] = Ek; }
    q[0] = ell;
    q[1] = geta;
#pragma unroll
    for (int k = 0; k < KSM; ++k)
      if (k < ks) q[2 + k] = geta * s.x[k];
    q[2 + ks] = s.y0 - Ek * P0;
#pragma unroll
    for (int k = 0; k < KOM; ++k)
      if (k < ko) q[3 + ks + k] = tile[(p.L.off_sy + 2 + k) * kWarp + lane] - Ek * pw[k];
  }
};

template <typename T, int KS, int KO, bool STRICT>
static cudaError_t launch_nmix_one(const EvalParams& p, dim3 grid, size_t smem, cudaStream_t stream, int* occ) {
  auto kern = eval_kernel<T, NmixModel<T, KS, KO, STRICT>, 2>;
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

static cudaError_t ensure_nmix_tables() {
  static bool done[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 64 && done[dev]) return cudaSuccess;
  constexpr int n = 2 * kNmixMaxAbundance + 2;
  static double h[n];
  static float hf[n];
  for (int k = 0; k < n; ++k) {
    h[k] = std::lgamma((double)k + 1.0);
    hf[k] = (float)h[k];
  }
  cudaError_t e = cudaMemcpyToSymbol(d_lgf64, h, sizeof(h));
  if (e == cudaSuccess) e = cudaMemcpyToSymbol(d_lgf32, hf, sizeof(hf));
  if (e == cudaSuccess && dev < 64) done[dev] = true;
  return e;
}

cudaError_t launch_nmixture(const EvalParams& p, int dtype, dim3 grid, size_t smem, cudaStream_t stream, int* occ) {
  if (!occ) {
    cudaError_t e = ensure_nmix_tables();
    if (e != cudaSuccess) return e;
  }
  const bool strict = (p.flags & BL_FLAG_STRICT_MATH) != 0;
  const bool s53 = p.L.ks == 5 && p.L.ko == 3;
  if (dtype == BL_F64)
    return s53 ? launch_nmix_one<double, 5, 3, true>(p, grid, smem, stream, occ)
               : launch_nmix_one<double, -1, -1, true>(p, grid, smem, stream, occ);
  if (strict)
    return s53 ? launch_nmix_one<float, 5, 3, true>(p, grid, smem, stream, occ)
               : launch_nmix_one<float, -1, -1, true>(p, grid, smem, stream, occ);
  return s53 ? launch_nmix_one<float, 5, 3, false>(p, grid, smem, stream, occ)
             : launch_nmix_one<float, -1, -1, false>(p, grid, smem, stream, occ);
}

cudaError_t launch_nmixture_summary(const EvalParams& p, int dtype, float* out, cudaStream_t st) {
  cudaError_t e = ensure_nmix_tables();
  if (e != cudaSuccess) return e;
  if (dtype == BL_F64) return launch_summary<double, NmixModel<double, -1, -1, true>>(p, out, st);
  return launch_summary<float, NmixModel<float, -1, -1, true>>(p, out, st);
}

}  // namespace bl
