// K2: fused log-marginal + gradient of the Royle-Nichols abundance-induced heterogeneity model.
//
// Replaces value_and_grad(potential_fn) of biolith/models/occu_rn.py:175-222 (reference): per unit
//   eta = beta0 + X.beta_1:, lambda = exp(eta)                              (occu_rn.py:181-188)
//   N ~ Categorical(logits_k = k eta - lambda - lgamma(k+1)), k = 0..K      (utils/distributions.py:31-40,
//                                                                            normalised by CategoricalLogits)
//   r_j = sigmoid(alpha0 + W_j.alpha_1:),  P_kj = 1 - (1-c)(1-r_j)^k         (occu_rn.py:205-222)
//   A_k = log pi_k + sum_j m_j Bernoulli(P~_kj).log_prob(y_j);  l = logsumexp_k A_k
//   dl/deta = E_post[k] - E_prior[k];  dl/dnu_j = sum_k w_k dt_kj/dnu
// log(1 - P_kj) = k log(1-r_j) + log(1-c) is carried in log space and 1 - q^k by the all-positive
// recurrence P_k = P_{k-1} + q^{k-1} r, so neither the clamp decisions nor small P suffer the
// 1-(1-r)**N cancellation of the reference formulation (oracle/occupancy.py:occu_rn_logp_grad).
// Per-thread state A_k lives in shared memory ([K+1][256], conflict-free), visits outer / k inner.
#include <cmath>
#include <cstdlib>
#include <mutex>
#include <type_traits>

#include "engine.cuh"

namespace bl {

constexpr int kMaxAbundance = 1023;
__constant__ double c_lgamma[kMaxAbundance + 1];  // lgamma(k + 1)
__constant__ float c_lgamma_f[kMaxAbundance + 1];

template <typename T, int KS, int KO, bool STRICT>
struct OccuRnModel {
  using N = Num<T>;
  static constexpr bool kSfu = std::is_same<T, float>::value && !STRICT;
  using M = Mth<T, kSfu>;
  static constexpr bool kGeneric = (KS < 0);
  static constexpr int KSM = kGeneric ? kMaxCov : KS;
  static constexpr int KOM = kGeneric ? kMaxCov : KO;
  static constexpr int kNQMax = kGeneric ? (1 + 2 * (kMaxCov + 1) + 1) : 33;  // runtime NQ loop either way
  static constexpr int kDerived = 4;  // l1mc = log(1-c), c, 1-c, dc/dx = c(1-c)
  static constexpr int kMultiChain = 1;  // chains per pass over a warp-tile (engine.cuh)

  struct Site {
    T x[KSM];
  };
  static __device__ __forceinline__ T unit_const(const EvalParams&, const T*, int) { return T(0); }

  static __device__ __forceinline__ void derive(const EvalParams& p, T* th) {
    T* d = th + p.D;
    if (p.flags & BL_FLAG_FP_CONSTANT) {
      const T x = th[p.D - 1];
      const T cv = T(1) / (T(1) + N::exp_(-x));
      d[0] = -(N::max_(x, T(0)) + N::log1p_(N::exp_(-N::abs_(x))));
      d[1] = cv;
      d[2] = T(1) - cv;
      d[3] = cv * (T(1) - cv);
    } else {
      d[0] = T(0); d[1] = T(0); d[2] = T(1); d[3] = T(0);
    }
  }

  static __device__ __forceinline__ void load_site(const EvalParams& p, const T* __restrict__ tile, int lane,
                                                   Site& s) {
    const int ks = kGeneric ? p.L.ks : KS;
#pragma unroll
    for (int k = 0; k < KSM; ++k) s.x[k] = (k < ks) ? tile[k * kWarp + lane] : T(0);
  }

  static __device__ __forceinline__ void site_chain(const EvalParams& p, const T* __restrict__ tile, int lane,
                                                    const Site& s, const T* __restrict__ th, T* __restrict__ q,
                                                    T* __restrict__ extra = nullptr) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // per-thread column A[k] at scratch[k * 256 + tid] (placed after the engine's regions)
    // (or, when (K+1) x 256 elements do not fit, at global scratch[k * n_threads + gtid], coalesced)
    T* A;
    size_t ST;
    if (p.rn_scratch_global) {
      ST = (size_t)gridDim.x * gridDim.y * kBlockThreads;
      A = reinterpret_cast<T*>(p.rn_scratch_global) +
          ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * kBlockThreads + threadIdx.x;
    } else {
      ST = kBlockThreads;
      A = reinterpret_cast<T*>(smem_raw + p.rn_scratch_off) + threadIdx.x;
    }
    const int ks = kGeneric ? p.L.ks : KS;
    const int ko = kGeneric ? p.L.ko : KO;
    const int J = p.L.J, K = p.K;
    const bool fpc = (p.flags & BL_FLAG_FP_CONSTANT) != 0;
    const T l1mc = th[p.D + 0], cval = th[p.D + 1], omc = th[p.D + 2];
    T eta = th[0];
#pragma unroll
    for (int k = 0; k < KSM; ++k)
      if (k < ks) eta = N::fma_(s.x[k], th[1 + k], eta);
    const T* al = th + ks + 1;
    const T a0 = al[0];
    T a[KOM];
#pragma unroll
    for (int k = 0; k < KOM; ++k) a[k] = (k < ko) ? al[1 + k] : T(0);

    // ---- prior logits (the -lambda term cancels in the normalisation) and their normaliser
    T Mp = -N::inf();
    for (int k = 0; k <= K; ++k) {
      const T lk = N::fma_((T)k, eta, -(T)c_lgamma[k]);
      A[k * ST] = lk;
      Mp = N::max_(Mp, lk);
    }
    T Zp = T(0), Ep = T(0);
    for (int k = 0; k <= K; ++k) {
      const T e = M::exp_(A[k * ST] - Mp);
      Zp += e;
      Ep = N::fma_((T)k, e, Ep);
    }
    const T logZp = Mp + M::log_(Zp);
    Ep = Ep * M::rcp_(Zp);

    // ---- pass 1: A_k += sum_j m_j log Bernoulli(y_j | P~_kj)
    uint32_t yw = 0, mw = 0;
    const T* wrow = tile + p.L.off_w * kWarp + lane;
    for (int j = 0; j < J; ++j) {
      if ((j & 31) == 0) {
        yw = N::as_bits(tile[(p.L.off_y + (j >> 5)) * kWarp + lane]);
        mw = N::as_bits(tile[(p.L.off_m + (j >> 5)) * kWarp + lane]);
      }
      if (!((mw >> (j & 31)) & 1u)) continue;
      const bool y = (yw >> (j & 31)) & 1u;
      T nu = a0;
#pragma unroll
      for (int k = 0; k < KOM; ++k)
        if (k < ko) nu = N::fma_(wrow[(j * ko + k) * kWarp], a[k], nu);
      T sp, r;
      M::softsig(nu, sp, r);
      const T u = -sp;  // log(1 - r)
      if (!y) {
        for (int k = 0; k <= K; ++k) {
          const T lq = N::fma_((T)k, u, l1mc);
          A[k * ST] += N::max_(lq, N::log_eps());  // P <= tiny gives -tiny == lq in this precision
        }
      } else {
        const T qv = T(1) - r;
        T qk = T(1), P0 = T(0);  // q^k and 1 - q^k (all-positive recurrence)
        for (int k = 0; k <= K; ++k) {
          const T lq = N::fma_((T)k, u, l1mc);
          const T P = N::fma_(omc, P0, cval);  // c + (1-c)(1-q^k)
          const bool lo = P <= -N::neg_tiny();
          const bool hi = lq <= N::log_eps();
          const T t = lo ? N::log_tiny() : (hi ? N::log1m_eps() : M::log_(P));
          A[k * ST] += t;
          P0 = N::fma_(qk, r, P0);
          qk *= qv;
        }
      }
    }
    // ---- posterior over N
    T Mx = -N::inf();
    for (int k = 0; k <= K; ++k) Mx = N::max_(Mx, A[k * ST]);
    T Z = T(0);
    for (int k = 0; k <= K; ++k) {
      const T e = M::exp_(A[k * ST] - Mx);
      A[k * ST] = e;
      Z += e;
    }
    const T iZ = M::rcp_(Z);
    T Eq = T(0);
    for (int k = 0; k <= K; ++k) {
      const T w = A[k * ST] * iZ;
      A[k * ST] = w;
      Eq = N::fma_((T)k, w, Eq);
    }
    const T ell = (Mx + M::log_(Z)) - logZp;
    const T geta = Eq - Ep;
    if (extra) { extra[0] = N::exp_(eta); extra[1] = Eq; }  // lambda, E[N | y]

    // ---- pass 2: dl/dnu_j = sum_k w_k dt_kj/dnu  (and dl/dc)
    T ga0 = T(0), gc = T(0);
    T ga[KOM];
#pragma unroll
    for (int k = 0; k < KOM; ++k) ga[k] = T(0);
    for (int j = 0; j < J; ++j) {
      if ((j & 31) == 0) {
        yw = N::as_bits(tile[(p.L.off_y + (j >> 5)) * kWarp + lane]);
        mw = N::as_bits(tile[(p.L.off_m + (j >> 5)) * kWarp + lane]);
      }
      if (!((mw >> (j & 31)) & 1u)) continue;
      const bool y = (yw >> (j & 31)) & 1u;
      T w[KOM];
      T nu = a0;
#pragma unroll
      for (int k = 0; k < KOM; ++k) {
        w[k] = (k < ko) ? wrow[(j * ko + k) * kWarp] : T(0);
        nu = N::fma_(w[k], a[k], nu);
      }
      T sp, r;
      M::softsig(nu, sp, r);
      const T u = -sp;
      T g = T(0);  // sum_k w_k dt/dlq * k   (then times dlq/dnu = -r per unit k)
      T gcj = T(0);
      if (!y) {
        for (int k = 0; k <= K; ++k) {
          const T lq = N::fma_((T)k, u, l1mc);
          const T wk = (lq > N::log_eps()) ? A[k * ST] : T(0);  // dt/dlq = 1 inside the clamp
          g = N::fma_((T)k, wk, g);
          gcj += wk;
        }
      } else {
        const T qv = T(1) - r;
        T qk = T(1), P0 = T(0);
        for (int k = 0; k <= K; ++k) {
          const T lq = N::fma_((T)k, u, l1mc);
          const T P = N::fma_(omc, P0, cval);
          const bool inr = (P > -N::neg_tiny()) && (lq > N::log_eps());
          // dt/dlq = -(1-P)/P with 1-P = (1-c) q^k
          const T dt = inr ? -(omc * qk) * M::rcp_(P) : T(0);
          const T wd = A[k * ST] * dt;
          g = N::fma_((T)k, wd, g);
          gcj += wd;
          P0 = N::fma_(qk, r, P0);
          qk *= qv;
        }
      }
      const T gnu = -r * g;  // dlq/dnu = -k r
      ga0 += gnu;
#pragma unroll
      for (int k = 0; k < KOM; ++k)
        if (k < ko) ga[k] = N::fma_(gnu, w[k], ga[k]);
      gc += gcj;
    }
    q[0] = ell;
    q[1] = geta;
#pragma unroll
    for (int k = 0; k < KSM; ++k)
      if (k < ks) q[2 + k] = geta * s.x[k];
    q[2 + ks] = ga0;
#pragma unroll
    for (int k = 0; k < KOM; ++k)
      if (k < ko) q[3 + ks + k] = ga[k];
    if (fpc) q[3 + ks + ko] = -gc * th[p.D + 3] / omc;  // dlq/dc = -1/(1-c); times dc/dx
  }
};

// ------------------------------------------------------------------------------------------------
// K2c: chain-parallel Royle-Nichols kernel (fp32, SFU math, C >= 64): lane = chain, sites and their
// detection / mask bits are warp-uniform.  Two consequences make it ~4x cheaper than the site-parallel
// form above: (1) the y branch is uniform, so the expensive k-loop with log(P_k) runs only for visits
// with a detection; (2) non-detections are linear in k, log(1-P_k) = k u_j + log(1-c), except where
// the clamp binds (k u_j + log(1-c) <= log eps), so they are folded into one slope U = sum u_j plus a
// short descending correction loop over the clamped tail states only.
// ------------------------------------------------------------------------------------------------
// RT = true: KS / KO are capacities and the actual covariate counts come from the layout (the k-loops
// dominate, so the predicated site-level loops cost nothing measurable)
template <int KS, int KO, int BT, bool RT>
__global__ void __launch_bounds__(BT, BT == 128 ? 4 : 2) occu_rn_chain_kernel(const __grid_constant__ EvalParams p) {
  using N = Num<float>;
  using M = Mth<float, true>;
  constexpr int KB = KS + 1, KA = KO + 1;
  const int ks = RT ? p.L.ks : KS, ko = RT ? p.L.ko : KO;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
  float* stage0 = reinterpret_cast<float*>(smem_raw + 128);
  __shared__ int s_is_last;
  const int F = p.L.F, J = p.L.J, K = p.K, NQ = p.NQ, D = p.D;
  const bool fpc = (p.flags & BL_FLAG_FP_CONSTANT) != 0;
  const uint32_t tile_elems = (uint32_t)F * kWarp;
  const uint32_t tile_bytes = tile_elems * sizeof(float);
  float* yfx = stage0 + (size_t)p.nstage * tile_elems;
  float* mfx = yfx + (size_t)J * kWarp;
  const int tid = threadIdx.x;
  double* g64 = reinterpret_cast<double*>(mfx + (size_t)J * kWarp) + tid;      // [NQ][BT]
  float* A = reinterpret_cast<float*>(g64 - tid + (size_t)NQ * BT) + tid;      // [K+1][BT]
  float* U0 = A - tid + (size_t)(K + 1) * BT + tid;                            // [J][BT] log(1-r_j) of non-detections
  const int c0 = blockIdx.y * p.CB;
  const int ncb = min(p.CB, p.C - c0);
  const bool chain_ok = tid < ncb;
  const bool warp_on = (tid & ~31) < ncb;  // warps past the end of the batch only help stage and expand
  const int64_t nbt = p.n_block_tiles;
  const int64_t bt_begin = nbt * blockIdx.x / gridDim.x;
  const int64_t bt_end = nbt * (blockIdx.x + 1) / gridDim.x;
  const int n_it = (int)(bt_end - bt_begin);
  const float* packed = reinterpret_cast<const float*>(p.packed);

  if (tid == 0) {
    for (int s = 0; s < p.nstage; ++s) mbar_init(&bars[s], 1);
    fence_mbar_init();
  }
  float b[KB], a[KA];
  float l1mc = 0.f, cval = 0.f, omc = 1.f, dcdx = 0.f;
  {
    const float* th = reinterpret_cast<const float*>(p.theta) + (size_t)(c0 + (chain_ok ? tid : 0)) * D;
#pragma unroll
    for (int k = 0; k < KB; ++k) b[k] = (k <= ks) ? th[k] : 0.f;
#pragma unroll
    for (int k = 0; k < KA; ++k) a[k] = (k <= ko) ? th[ks + 1 + k] : 0.f;
    if (fpc) {
      const float x = th[D - 1];
      cval = 1.f / (1.f + expf(-x));
      l1mc = -(fmaxf(x, 0.f) + log1pf(expf(-fabsf(x))));
      omc = 1.f - cval;
      dcdx = cval * omc;
    }
  }
  for (int i = 0; i < NQ; ++i) g64[(size_t)i * BT] = 0.0;
  __syncthreads();
  if (tid == 0) {
    const int pre = min(p.nstage, n_it);
    for (int s = 0; s < pre; ++s) {
      mbar_expect_tx(&bars[s], tile_bytes);
      tma_load_bulk(stage0 + (size_t)s * tile_elems, packed + (size_t)(bt_begin + s) * tile_elems, tile_bytes,
                    &bars[s]);
    }
  }
  const float log_eps = N::log_eps();

  for (int it = 0; it < n_it; ++it) {
    const int s = it % p.nstage;
    mbar_wait(&bars[s], (uint32_t)((it / p.nstage) & 1));
    const float* tile = stage0 + (size_t)s * tile_elems;
    const int64_t unit0 = (bt_begin + it) * kWarp;
    const int n_valid = (int)max((int64_t)0, min((int64_t)kWarp, p.L.n_units - unit0));
    for (int e = tid; e < J * kWarp; e += BT) {
      const int site = e & 31, j = e >> 5;
      const uint32_t yw = __float_as_uint(tile[(p.L.off_y + (j >> 5)) * kWarp + site]);
      const uint32_t mw = __float_as_uint(tile[(p.L.off_m + (j >> 5)) * kWarp + site]);
      yfx[e] = ((yw >> (j & 31)) & 1u) ? 1.f : 0.f;
      mfx[e] = ((mw >> (j & 31)) & 1u) ? 1.f : 0.f;
    }
    __syncthreads();
    float acc_gb[KB], acc_ga[KA], acc_gc = 0.f;
    double logp_tile = 0.0;
#pragma unroll
    for (int k = 0; k < KB; ++k) acc_gb[k] = 0.f;
#pragma unroll
    for (int k = 0; k < KA; ++k) acc_ga[k] = 0.f;

    const int n_mine = warp_on ? n_valid : 0;
    for (int si = 0; si < n_mine; ++si) {
      float x[KS > 0 ? KS : 1];
      float eta = b[0];
#pragma unroll
      for (int k = 0; k < KS; ++k) {
        x[k] = (k < ks) ? tile[k * kWarp + si] : 0.f;
        eta = fmaf(x[k], b[1 + k], eta);
      }
      // ---- prior logits k eta - lgamma(k+1) (the -lambda term cancels) and their normaliser.  The
      // sequence is concave in k with its maximum at floor(lambda) or the next integer, so the shift of
      // the log-sum-exp is known up front and one pass suffices.
      float Mp;
      {
        const float lam = M::exp_(fminf(eta, 80.f));
        const int k0 = (int)fminf(lam, (float)K);
        const int k1 = min(k0 + 1, K);
        Mp = fmaxf(fmaf((float)k0, eta, -c_lgamma_f[k0]), fmaf((float)k1, eta, -c_lgamma_f[k1]));
      }
      float Zp = 0.f, Ep = 0.f, kf = 0.f;
#pragma unroll 4
      for (int k = 0; k <= K; ++k) {
        const float lk = fmaf(kf, eta, -c_lgamma_f[k]);
        A[(size_t)k * BT] = lk;
        const float e = M::exp_(lk - Mp);
        Zp += e;
        Ep = fmaf(kf, e, Ep);
        kf += 1.f;
      }
      const float logZp = Mp + M::log_(Zp);
      Ep *= M::rcp_(Zp);
      // ---- pass 1: detections run the k-loop; non-detections only record u_j = log(1-r_j)
      float Utot = 0.f, umin = 0.f;
      int n0i = 0;
      for (int j = 0; j < J; ++j) {
        if (mfx[j * kWarp + si] == 0.f) continue;  // warp-uniform
        float nu = a[0];
#pragma unroll
        for (int k = 0; k < KO; ++k)
          if (k < ko) nu = fmaf(tile[(p.L.off_w + j * ko + k) * kWarp + si], a[1 + k], nu);
        float sp, r;
        M::softsig(nu, sp, r);
        const float u = -sp;
        if (yfx[j * kWarp + si] == 0.f) {  // warp-uniform
          Utot += u;
          umin = fminf(umin, u);
          U0[(size_t)n0i * BT] = u;
          ++n0i;
        } else {
          const float qv = 1.f - r;
          float qk = 1.f, P0 = 0.f, kk = 0.f;
#pragma unroll 4
          for (int k = 0; k <= K; ++k) {
            const float lq = fmaf(kk, u, l1mc);
            const float P = fmaf(omc, P0, cval);
            float t = M::log_(P);
            t = (lq <= log_eps) ? N::log1m_eps() : t;
            t = (P <= -N::neg_tiny()) ? N::log_tiny() : t;
            A[(size_t)k * BT] += t;
            P0 = fmaf(qk, r, P0);
            qk *= qv;
            kk += 1.f;
          }
        }
      }
      // non-detections are linear in k except on the clamped tail (k u_j + log(1-c) <= log eps): one
      // descending k-loop adds max(log eps - lq_kj, 0) over all of them and stops at the first state
      // where no visit is clamped (lq grows as k falls)
      if (fmaf((float)K, umin, l1mc) <= log_eps) {
        float kd = (float)K;
        for (int k = K; k >= 0; --k) {
          float corr = 0.f;
          for (int i = 0; i < n0i; ++i) corr += fmaxf(log_eps - fmaf(kd, U0[(size_t)i * BT], l1mc), 0.f);
          if (corr == 0.f) break;
          A[(size_t)k * BT] += corr;
          kd -= 1.f;
        }
      }
      // ---- posterior over N (adds the linear non-detection part k U + n0 log(1-c)); the weights stay
      // unnormalised (e_k = exp(A_k - max)) and the sums are scaled by 1/Z afterwards
      const float off0 = (float)n0i * l1mc;
      float Mx = -N::inf();
      kf = 0.f;
#pragma unroll 4
      for (int k = 0; k <= K; ++k) {
        const float v = A[(size_t)k * BT] + fmaf(kf, Utot, off0);
        A[(size_t)k * BT] = v;
        Mx = fmaxf(Mx, v);
        kf += 1.f;
      }
      float Z = 0.f, Eqn = 0.f;
      kf = 0.f;
#pragma unroll 4
      for (int k = 0; k <= K; ++k) {
        const float e = M::exp_(A[(size_t)k * BT] - Mx);
        A[(size_t)k * BT] = e;
        Z += e;
        Eqn = fmaf(kf, e, Eqn);
        kf += 1.f;
      }
      const float iZ = M::rcp_(Z);
      const float ell = (Mx + M::log_(Z)) - logZp;
      const float geta = Eqn * iZ - Ep;
      // ---- pass 2 (sums over the unnormalised weights, scaled by 1/Z per visit).  Sweep 1: detections.
      float ga0 = 0.f, gc = 0.f, ga[KO > 0 ? KO : 1];
#pragma unroll
      for (int k = 0; k < KO; ++k) ga[k] = 0.f;
      for (int j = 0; j < J; ++j) {
        if (mfx[j * kWarp + si] == 0.f || yfx[j * kWarp + si] == 0.f) continue;
        float w[KO > 0 ? KO : 1];
        float nu = a[0];
#pragma unroll
        for (int k = 0; k < KO; ++k) {
          w[k] = (k < ko) ? tile[(p.L.off_w + j * ko + k) * kWarp + si] : 0.f;
          nu = fmaf(w[k], a[1 + k], nu);
        }
        float sp, r;
        M::softsig(nu, sp, r);
        const float u = -sp;
        const float qv = 1.f - r;
        float qk = 1.f, P0 = 0.f, kk = 0.f, g = 0.f, gcj = 0.f;
#pragma unroll 4
        for (int k = 0; k <= K; ++k) {
          const float lq = fmaf(kk, u, l1mc);
          const float P = fmaf(omc, P0, cval);
          const bool inr = (P > -N::neg_tiny()) && (lq > log_eps);
          float dt = -(omc * qk) * M::rcp_(P);  // dt/dlq = -(1-P)/P, 1-P = (1-c) q^k
          dt = inr ? dt : 0.f;
          const float wd = A[(size_t)k * BT] * dt;
          g = fmaf(kk, wd, g);
          gcj += wd;
          P0 = fmaf(qk, r, P0);
          qk *= qv;
          kk += 1.f;
        }
        const float gnu = -r * g * iZ;  // dlq/dnu = -k r
        ga0 += gnu;
#pragma unroll
        for (int k = 0; k < KO; ++k) ga[k] = fmaf(gnu, w[k], ga[k]);
        gc = fmaf(gcj, iZ, gc);
      }
      // Sweep 2: non-detections need sum_{k <= kc_j} k e_k (kc_j = last unclamped state): turn the
      // column into that prefix sum in place and look it up per visit.  With a false-positive constant
      // the k-unweighted prefix is needed too and the (rare) tail loop is kept instead.
      if (!fpc) {
        float run = 0.f;
        kf = 0.f;
#pragma unroll 4
        for (int k = 0; k <= K; ++k) {
          run = fmaf(kf, A[(size_t)k * BT], run);
          A[(size_t)k * BT] = run;
          kf += 1.f;
        }
      }
      for (int j = 0; j < J; ++j) {
        if (mfx[j * kWarp + si] == 0.f || yfx[j * kWarp + si] != 0.f) continue;
        float w[KO > 0 ? KO : 1];
        float nu = a[0];
#pragma unroll
        for (int k = 0; k < KO; ++k) {
          w[k] = (k < ko) ? tile[(p.L.off_w + j * ko + k) * kWarp + si] : 0.f;
          nu = fmaf(w[k], a[1 + k], nu);
        }
        float sp, r;
        M::softsig(nu, sp, r);
        const float u = -sp;
        float g, gcj = 0.f;
        if (!fpc) {
          // kc = largest k with k u + log(1-c) > log eps, decided by the same fmaf as pass 1
          int kc = K;
          if (fmaf((float)K, u, l1mc) <= log_eps) {
            kc = (int)fminf((log_eps - l1mc) / u, (float)K);  // u < 0 here
            while (kc < K && fmaf((float)(kc + 1), u, l1mc) > log_eps) ++kc;
            while (kc >= 0 && !(fmaf((float)kc, u, l1mc) > log_eps)) --kc;
          }
          g = kc >= 0 ? A[(size_t)kc * BT] : 0.f;
        } else {
          float tk = 0.f, t0 = 0.f, kd = (float)K;  // (k-weighted) weight of the clamped tail states
          for (int k = K; k >= 0; --k) {
            const float lq = fmaf(kd, u, l1mc);
            if (lq > log_eps) break;
            const float wk = A[(size_t)k * BT];
            tk = fmaf(kd, wk, tk);
            t0 += wk;
            kd -= 1.f;
          }
          g = Eqn - tk;
          gcj = Z - t0;
        }
        const float gnu = -r * g * iZ;
        ga0 += gnu;
#pragma unroll
        for (int k = 0; k < KO; ++k) ga[k] = fmaf(gnu, w[k], ga[k]);
        gc = fmaf(gcj, iZ, gc);
      }
      logp_tile += (double)ell;
      acc_gb[0] += geta;
#pragma unroll
      for (int k = 0; k < KS; ++k) acc_gb[1 + k] = fmaf(geta, x[k], acc_gb[1 + k]);
      acc_ga[0] += ga0;
#pragma unroll
      for (int k = 0; k < KO; ++k) acc_ga[1 + k] += ga[k];
      acc_gc += gc;
    }
    g64[0] += logp_tile;
#pragma unroll
    for (int k = 0; k < KB; ++k)
      if (k <= ks) g64[(size_t)(1 + k) * BT] += (double)acc_gb[k];
#pragma unroll
    for (int k = 0; k < KA; ++k)
      if (k <= ko) g64[(size_t)(2 + ks + k) * BT] += (double)acc_ga[k];
    if (fpc) g64[(size_t)(3 + ks + ko) * BT] += (double)(-acc_gc * dcdx / omc);
    __syncthreads();
    if (tid == 0 && it + p.nstage < n_it) {
      mbar_expect_tx(&bars[s], tile_bytes);
      tma_load_bulk(stage0 + (size_t)s * tile_elems, packed + (size_t)(bt_begin + it + p.nstage) * tile_elems,
                    tile_bytes, &bars[s]);
    }
  }
  if (chain_ok) {
    double* my = p.partial + ((size_t)blockIdx.x * p.C + c0 + tid) * NQ;
    for (int i = 0; i < NQ; ++i) my[i] = g64[(size_t)i * BT];
  }
  finish_block<float>(p, c0, ncb, &s_is_last);
}

static int rn_chain_variant() {
  const char* e = getenv("BL_CHAIN_VARIANT");  // tuning switch: 2 = 128-thread blocks, 3 = 256-thread blocks
  return e ? atoi(e) : 0;
}

bool occu_rn_chain_supported(int dtype, int ks, int ko, uint32_t flags) {
  if (dtype != BL_F32 || (flags & BL_FLAG_STRICT_MATH)) return false;
  return ks >= 0 && ks <= 8 && ko >= 0 && ko <= 4;  // (5,3) specialised, the rest through the capacity variant
}

// threads (= chains) per block for a batch of C chains; see occu_chain.cu:occu_chain_block_threads.  The
// per-thread A_k / U0 columns make 128-thread blocks 10 % slower on full batches, so they are used only
// where they pad the batch less.  Measured on B200 (config 3, K = 50, ms per evaluation, 128 / 256 threads):
// C=32 9.1 / 13.3, 64 10.8 / 15.9, 128 13.1 / 15.9, 192 25.2 / 23.3, 256 25.2 / 23.4, 384 37.9 / 45.0,
// 512 49.7 / 45.1, 1024 99.5 / 89.1 (site-parallel engine: 32 14.7, 64 29.5, 128 53.8).
int occu_rn_chain_block_threads(int C) {
  if (rn_chain_variant() == 2) return 128;
  if (rn_chain_variant() == 3) return 256;
  return (C > 128 && (C + 127) / 128 % 2 == 0) ? 256 : 128;
}

size_t occu_rn_chain_smem(const Layout& L, int nstage, int K, int D, int bt) {
  size_t bts = 128 + (size_t)nstage * L.F * kWarp * sizeof(float) + 2 * (size_t)L.J * kWarp * sizeof(float);
  bts = (bts + 15) & ~size_t(15);
  bts += (size_t)(1 + D) * bt * sizeof(double);
  return bts + (size_t)(K + 1 + L.J) * bt * sizeof(float);
}

template <int KS, int KO, int BT, bool RT>
static cudaError_t launch_rn_chain_one(const EvalParams& p, dim3 grid, size_t smem, cudaStream_t st, int* occ) {
  auto kern = occu_rn_chain_kernel<KS, KO, BT, RT>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  if (occ) return cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, kern, BT, smem);
  kern<<<grid, BT, smem, st>>>(p);
  return cudaGetLastError();
}

template <int KS, int KO, bool RT>
static cudaError_t launch_rn_chain_bt(const EvalParams& p, dim3 grid, size_t smem, cudaStream_t st, int* occ) {
  if (p.chain_bt == 128) return launch_rn_chain_one<KS, KO, 128, RT>(p, grid, smem, st, occ);
  return launch_rn_chain_one<KS, KO, 256, RT>(p, grid, smem, st, occ);
}

static cudaError_t ensure_lgamma_table();

cudaError_t launch_occu_rn_chain(const EvalParams& p, dim3 grid, size_t smem, cudaStream_t st, int* occ) {
  if (!occ) {
    cudaError_t e = ensure_lgamma_table();
    if (e != cudaSuccess) return e;
  }
  if (p.L.ks == 5 && p.L.ko == 3) return launch_rn_chain_bt<5, 3, false>(p, grid, smem, st, occ);
  if (p.L.ks <= 8 && p.L.ko <= 4) return launch_rn_chain_bt<8, 4, true>(p, grid, smem, st, occ);
  return cudaErrorNotSupported;
}

template <typename T, int KS, int KO, bool STRICT>
static cudaError_t launch_rn_one(const EvalParams& p, dim3 grid, size_t smem, cudaStream_t stream, int* occ) {
  auto kern = eval_kernel<T, OccuRnModel<T, KS, KO, STRICT>, 1>;
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

static cudaError_t ensure_lgamma_table() {
  static std::once_flag once;
  static cudaError_t status = cudaSuccess;
  // per device: constant memory is per context; guard by device id
  static bool done[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 64 && done[dev]) return cudaSuccess;
  double h[kMaxAbundance + 1];
  for (int k = 0; k <= kMaxAbundance; ++k) h[k] = std::lgamma((double)k + 1.0);
  status = cudaMemcpyToSymbol(c_lgamma, h, sizeof(h));
  if (status == cudaSuccess) {
    static float hf[kMaxAbundance + 1];
    for (int k = 0; k <= kMaxAbundance; ++k) hf[k] = (float)h[k];
    status = cudaMemcpyToSymbol(c_lgamma_f, hf, sizeof(hf));
  }
  if (status == cudaSuccess && dev < 64) done[dev] = true;
  (void)once;
  return status;
}

cudaError_t launch_occu_rn(const EvalParams& p, int dtype, dim3 grid, size_t smem, cudaStream_t stream, int* occ) {
  if (!occ) {
    cudaError_t e = ensure_lgamma_table();
    if (e != cudaSuccess) return e;
  }
  const bool strict = (p.flags & BL_FLAG_STRICT_MATH) != 0;
  const bool s53 = p.L.ks == 5 && p.L.ko == 3;
  if (dtype == BL_F64)
    return s53 ? launch_rn_one<double, 5, 3, true>(p, grid, smem, stream, occ)
               : launch_rn_one<double, -1, -1, true>(p, grid, smem, stream, occ);
  if (strict)
    return s53 ? launch_rn_one<float, 5, 3, true>(p, grid, smem, stream, occ)
               : launch_rn_one<float, -1, -1, true>(p, grid, smem, stream, occ);
  return s53 ? launch_rn_one<float, 5, 3, false>(p, grid, smem, stream, occ)
             : launch_rn_one<float, -1, -1, false>(p, grid, smem, stream, occ);
}

cudaError_t launch_occu_rn_summary(const EvalParams& p, int dtype, float* out, cudaStream_t st) {
  cudaError_t e = ensure_lgamma_table();
  if (e != cudaSuccess) return e;
  if (dtype == BL_F64) return launch_summary<double, OccuRnModel<double, -1, -1, true>>(p, out, st);
  return launch_summary<float, OccuRnModel<float, -1, -1, true>>(p, out, st);
}

int occu_rn_derived_slots(uint32_t) { return 4; }

size_t occu_rn_extra_smem(const Layout&, int K, int elem) { return (size_t)(K + 1) * kBlockThreads * elem + 128; }

}  // namespace bl
