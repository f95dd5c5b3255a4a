// K1: fused log-marginal + gradient of the Bernoulli occupancy model (MacKenzie 2002).
//
// Replaces value_and_grad(potential_fn) of biolith/models/occu.py:182-242 (reference): per unit
//   eta = beta0 + X.beta_1:            (regression/linear.py:59-66)      psi = sigmoid(eta)   (occu.py:207)
//   nu_j = alpha0 + W_j.alpha_1:       (occu.py:221-228)                 p_j = sigmoid(nu_j)
//   L1 = sum_j m_j [y_j log p~_j + (1-y_j) log1p(-p~_j)]                (occu.py:229-242, z = 1 branch)
//   L0 = n1 log(P0~) + n0 log1p(-P0~),  P0 = 1-(1-c)(1-u)  (tiny-clamped when no false positives)
//   l  = logaddexp(log psi~ + L1, log1p(-psi~) + L0)                    (funsor sum-product over z)
//   r  = P(z=1 | y) = sigmoid(a - b);  dl/deta = r - psi;  dl/dnu_j = r m_j (y_j - p_j)
// with numpyro's clamp_probs semantics (zero derivative outside [tiny, 1-eps]) decided in log space.
// The closed form is oracle/occupancy.py:occu_logp_grad; parity tests compare against it.
#include <type_traits>

#include "engine.cuh"

namespace bl {

// FP = a false-positive flag is set (exactly one extra parameter: logit c or logit u)
template <typename T, int KS, int KO, bool FP, bool STRICT>
struct OccuModel {
  using N = Num<T>;
  // bounded-error SFU math for fp32 unless BL_FLAG_STRICT_MATH (fp64 is always libm)
  static constexpr bool kSfu = std::is_same<T, float>::value && !STRICT && !FP;
  static constexpr bool kGeneric = (KS < 0);
  static constexpr int KSM = kGeneric ? kMaxCov : KS;
  static constexpr int KOM = kGeneric ? kMaxCov : KO;
  static constexpr int kNQMax = kGeneric ? (1 + 2 * (kMaxCov + 1) + 1) : (1 + KS + 1 + KO + 1 + (FP ? 1 : 0));
  // derived per-chain slots (after the D raw parameters): l1mc, lP0, l1mP0, iP0, i1mP0, sP0, sx, -
  static constexpr int kDerived = FP ? 8 : 0;

  struct Site {
    T x[KSM];
    T n1, n0;  // masked detections / non-detections (data only)
  };
  static __device__ __forceinline__ T unit_const(const EvalParams&, const T*, int) { return T(0); }

  static __device__ __forceinline__ void derive(const EvalParams& p, T* th) {
    if constexpr (FP) {
      const int D = p.D;
      const T x = th[D - 1];
      const bool is_c = (p.flags & BL_FLAG_FP_CONSTANT) != 0;
      // c = sigmoid(x): log(1-c) = -softplus(x)
      const T l1m = -(N::max_(x, T(0)) + N::log1p_(N::exp_(-N::abs_(x))));
      const T cv = T(1) / (T(1) + N::exp_(-x));
      T* d = th + D;
      d[0] = is_c ? l1m : T(0);                 // log(1-c) entering the z=1 branch
      const T l1mP0 = l1m;                      // log(1-P0) = log(1-c) + log(1-u), one of them is 0
      const T P0 = -N::expm1_(l1mP0);
      const bool in0 = (l1mP0 > N::log_eps()) && (P0 > -N::neg_tiny());
      d[1] = in0 ? N::log_(P0) : (P0 <= -N::neg_tiny() ? N::log_tiny() : N::log1m_eps());
      d[2] = in0 ? l1mP0 : (P0 <= -N::neg_tiny() ? N::neg_tiny() : N::log_eps());
      d[3] = in0 ? T(1) / P0 : T(0);
      d[4] = in0 ? T(1) / (T(1) - P0) : T(0);
      d[5] = T(1);                              // dP0/dc = (1-u), dP0/du = (1-c): the other one is 0
      d[6] = cv * (T(1) - cv);                  // d(c or u)/dx
      d[7] = T(1) / (T(1) - cv);                // 1/(1-c)
    }
  }

  static __device__ __forceinline__ void load_site(const EvalParams& p, const T* __restrict__ tile, int lane,
                                                   Site& s) {
    const int ks = kGeneric ? p.L.ks : KS;
#pragma unroll
    for (int k = 0; k < KSM; ++k) s.x[k] = (k < ks) ? tile[k * kWarp + lane] : T(0);
    int n1 = 0, n0 = 0;
    for (int w = 0; w < p.L.nw; ++w) {
      const uint32_t yw = N::as_bits(tile[(p.L.off_y + w) * kWarp + lane]);
      const uint32_t mw = N::as_bits(tile[(p.L.off_m + w) * kWarp + lane]);
      n1 += __popc(yw & mw);
      n0 += __popc(~yw & mw);
    }
    s.n1 = (T)n1;
    s.n0 = (T)n0;
  }

  static __device__ __forceinline__ void site_chain(const EvalParams& p, const T* __restrict__ tile, int lane,
                                                    const Site& s, const T* __restrict__ th, T* __restrict__ q,
                                                    T* __restrict__ extra = nullptr) {
    const int ks = kGeneric ? p.L.ks : KS;
    const int ko = kGeneric ? p.L.ko : KO;
    const int J = p.L.J;
    T eta = th[0];
#pragma unroll
    for (int k = 0; k < KSM; ++k)
      if (k < ks) eta = N::fma_(s.x[k], th[1 + k], eta);
    const T* al = th + ks + 1;
    const T a0 = al[0];
    T a[KOM];
#pragma unroll
    for (int k = 0; k < KOM; ++k) a[k] = (k < ko) ? al[1 + k] : T(0);
    T l1mc = T(0);
    if constexpr (FP) l1mc = th[p.D + 0];

    T L1 = T(0), ga0 = T(0), gc = T(0);
    T ga[KOM];
#pragma unroll
    for (int k = 0; k < KOM; ++k) ga[k] = T(0);

    uint32_t yw = 0, mw = 0;
    const T* wrow = tile + p.L.off_w * kWarp + lane;
#pragma unroll 4
    for (int j = 0; j < J; ++j) {
      if ((j & 31) == 0) {
        yw = N::as_bits(tile[(p.L.off_y + (j >> 5)) * kWarp + lane]);
        mw = N::as_bits(tile[(p.L.off_m + (j >> 5)) * kWarp + lane]);
      }
      const bool m = (mw >> (j & 31)) & 1u;
      const bool y = (yw >> (j & 31)) & 1u;
      T w[KOM];
      T nu = a0;
#pragma unroll
      for (int k = 0; k < KOM; ++k) {
        if (k < ko) {
          w[k] = wrow[(j * ko + k) * kWarp];
          nu = N::fma_(w[k], a[k], nu);
        } else {
          w[k] = T(0);
        }
      }
      T term, g;
      if constexpr (kSfu) {
        const sfu::SoftSig ss = sfu::softsig<true>(nu);
        const float yf = y ? 1.f : 0.f;
        term = fmaf(yf, ss.xc, -ss.s);
        g = ss.inr ? (yf - ss.p) : 0.f;
      } else if constexpr (!FP) {
        const LogSig<T> ls = log_sigmoid_pair<T>(nu);
        term = y ? ls.lp : ls.l1mp;
        g = ls.inr ? (y ? ls.q : -ls.p) : T(0);
      } else {
        // log(1-P1) = log(1-p) + log(1-c) carried in log space; clamps decided on exact quantities
        const T t = N::exp_(-N::abs_(nu));
        const T l = N::log1p_(t);
        const T inv = N::rcp_(T(1) + t);
        const T pj = (nu >= T(0)) ? inv : t * inv;
        const T lq = -N::max_(nu, T(0)) - l + l1mc;
        const T P1 = -N::expm1_(lq);
        const bool lo = P1 <= -N::neg_tiny();
        const bool inr = (lq > N::log_eps()) && !lo;
        T dt;  // dt/dlq
        if (y) {
          term = inr ? N::log_(P1) : (lo ? N::log_tiny() : N::log1m_eps());
          dt = inr ? -(T(1) - P1) / P1 : T(0);
        } else {
          term = inr ? lq : (lo ? N::neg_tiny() : N::log_eps());
          dt = inr ? T(1) : T(0);
        }
        g = -dt * pj;                 // dlq/dnu = -p
        gc += m ? dt : T(0);          // dlq/dc = -1/(1-c), applied once per unit below
      }
      term = m ? term : T(0);
      g = m ? g : T(0);
      L1 += term;
      ga0 += g;
#pragma unroll
      for (int k = 0; k < KOM; ++k)
        if (k < ko) ga[k] = N::fma_(g, w[k], ga[k]);
    }

    T L0, dL0 = T(0);
    if constexpr (!FP) {
      L0 = N::fma_(s.n1, N::log_tiny(), s.n0 * N::neg_tiny());
    } else {
      const T* d = th + p.D;
      L0 = N::fma_(s.n1, d[1], s.n0 * d[2]);
      dL0 = s.n1 * d[3] - s.n0 * d[4];  // dL0/dP0 (zero when P0 is clipped)
    }
    T ell, r, geta, psi_out;
    if constexpr (kSfu) {
      const sfu::SoftSig se = sfu::softsig<true>(eta);
      psi_out = se.p;
      const float av = (se.xc - se.s) + L1;
      const float bv = L0 - se.s;
      const float dd = av - bv;
      const float td = sfu::ex2(-fabsf(dd) * sfu::kLog2e);
      const float ud = 1.0f + td;
      const float inv = sfu::rcp(ud);
      r = (dd >= 0.f) ? inv : td * inv;
      ell = fmaf(sfu::lg2(ud), sfu::kLn2, fmaxf(av, bv));
      geta = se.inr ? (r - se.p) : 0.f;
    } else {
      const LogSig<T> se = log_sigmoid_pair<T>(eta);
      psi_out = se.p;
      const T av = se.lp + L1;
      const T bv = se.l1mp + L0;
      const T dd = av - bv;
      const T td = N::exp_(-N::abs_(dd));
      const T inv = N::rcp_(T(1) + td);
      r = (dd >= T(0)) ? inv : td * inv;
      ell = N::max_(av, bv) + N::log1p_(td);
      geta = se.inr ? (r - se.p) : T(0);
    }
    if (extra) { extra[0] = psi_out; extra[1] = r; }
    q[0] = ell;
    q[1] = geta;
#pragma unroll
    for (int k = 0; k < KSM; ++k)
      if (k < ks) q[2 + k] = geta * s.x[k];
    q[2 + ks] = r * ga0;
#pragma unroll
    for (int k = 0; k < KOM; ++k)
      if (k < ko) q[3 + ks + k] = r * ga[k];
    if constexpr (FP) {
      const T* d = th + p.D;
      const bool is_c = (p.flags & BL_FLAG_FP_CONSTANT) != 0;
      // constant fp: c enters both branches; unoccupied fp: only the z = 0 branch
      T gx = (T(1) - r) * dL0 * d[5];
      if (is_c) gx = N::fma_(r, -gc * d[7], gx);
      q[3 + ks + ko] = gx * d[6];
    }
  }

  // NCH chains of the same warp-tile in one pass over the visits (no false-positive extras): the chains'
  // dependent ex2 -> rcp -> lg2 sequences interleave (the engine is latency-bound at one chain per warp:
  // 16 resident warps per SM), and every covariate is read from shared memory once for all of them.
  // th0 + c * th_stride is the staged theta row of chain c; q[c] receives its NQ numbers.
  static constexpr int kMultiChain = (FP || kGeneric) ? 1 : (sizeof(T) == 8 ? 2 : 4);  // register budget
  template <int NCH>
  static __device__ __forceinline__ void site_chain_n(const EvalParams& p, const T* __restrict__ tile, int lane,
                                                      const Site& s, const T* __restrict__ th0, int th_stride,
                                                      T (*__restrict__ q)[kNQMax]) {
    static_assert(!FP, "multi-chain pass is for the plain model");
    const int ks = kGeneric ? p.L.ks : KS;
    const int ko = kGeneric ? p.L.ko : KO;
    const int J = p.L.J;
    T eta[NCH], a0[NCH], a[NCH][KOM], L1[NCH], ga0[NCH], ga[NCH][KOM];
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const T* th = th0 + (size_t)c * th_stride;
      eta[c] = th[0];
#pragma unroll
      for (int k = 0; k < KSM; ++k)
        if (k < ks) eta[c] = N::fma_(s.x[k], th[1 + k], eta[c]);
      a0[c] = th[ks + 1];
#pragma unroll
      for (int k = 0; k < KOM; ++k) {
        a[c][k] = (k < ko) ? th[ks + 2 + k] : T(0);
        ga[c][k] = T(0);
      }
      L1[c] = T(0);
      ga0[c] = T(0);
    }
    uint32_t yw = 0, mw = 0;
    const T* wrow = tile + p.L.off_w * kWarp + lane;
#pragma unroll 2
    for (int j = 0; j < J; ++j) {
      if ((j & 31) == 0) {
        yw = N::as_bits(tile[(p.L.off_y + (j >> 5)) * kWarp + lane]);
        mw = N::as_bits(tile[(p.L.off_m + (j >> 5)) * kWarp + lane]);
      }
      const bool m = (mw >> (j & 31)) & 1u;
      const bool y = (yw >> (j & 31)) & 1u;
      T w[KOM];
#pragma unroll
      for (int k = 0; k < KOM; ++k) w[k] = (k < ko) ? wrow[(j * ko + k) * kWarp] : T(0);
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        T nu = a0[c];
#pragma unroll
        for (int k = 0; k < KOM; ++k)
          if (k < ko) nu = N::fma_(w[k], a[c][k], nu);
        T term, g;
        if constexpr (kSfu) {
          const sfu::SoftSig ss = sfu::softsig<true>(nu);
          const float yf = y ? 1.f : 0.f;
          term = fmaf(yf, ss.xc, -ss.s);
          g = ss.inr ? (yf - ss.p) : 0.f;
        } else {
          const LogSig<T> ls = log_sigmoid_pair<T>(nu);
          term = y ? ls.lp : ls.l1mp;
          g = ls.inr ? (y ? ls.q : -ls.p) : T(0);
        }
        term = m ? term : T(0);
        g = m ? g : T(0);
        L1[c] += term;
        ga0[c] += g;
#pragma unroll
        for (int k = 0; k < KOM; ++k)
          if (k < ko) ga[c][k] = N::fma_(g, w[k], ga[c][k]);
      }
    }
    const T L0 = N::fma_(s.n1, N::log_tiny(), s.n0 * N::neg_tiny());
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      T ell, r, geta;
      if constexpr (kSfu) {
        const sfu::SoftSig se = sfu::softsig<true>(eta[c]);
        const float av = (se.xc - se.s) + L1[c];
        const float bv = L0 - se.s;
        const float dd = av - bv;
        const float td = sfu::ex2(-fabsf(dd) * sfu::kLog2e);
        const float ud = 1.0f + td;
        const float inv = sfu::rcp(ud);
        r = (dd >= 0.f) ? inv : td * inv;
        ell = fmaf(sfu::lg2(ud), sfu::kLn2, fmaxf(av, bv));
        geta = se.inr ? (r - se.p) : 0.f;
      } else {
        const LogSig<T> se = log_sigmoid_pair<T>(eta[c]);
        const T av = se.lp + L1[c];
        const T bv = se.l1mp + L0;
        const T dd = av - bv;
        const T td = N::exp_(-N::abs_(dd));
        const T inv = N::rcp_(T(1) + td);
        r = (dd >= T(0)) ? inv : td * inv;
        ell = N::max_(av, bv) + N::log1p_(td);
        geta = se.inr ? (r - se.p) : T(0);
      }
      q[c][0] = ell;
      q[c][1] = geta;
#pragma unroll
      for (int k = 0; k < KSM; ++k)
        if (k < ks) q[c][2 + k] = geta * s.x[k];
      q[c][2 + ks] = r * ga0[c];
#pragma unroll
      for (int k = 0; k < KOM; ++k)
        if (k < ko) q[c][3 + ks + k] = r * ga[c][k];
    }
  }
};

template <typename T, int KS, int KO, bool FP, bool STRICT, int MINB>
static cudaError_t launch_one(const EvalParams& p, dim3 grid, size_t smem, cudaStream_t stream, int* occ) {
  auto kern = eval_kernel<T, OccuModel<T, KS, KO, FP, STRICT>, MINB>;
  static bool configured = false;  // per instantiation
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  if (occ) return cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, kern, kBlockThreads, smem);
  kern<<<grid, kBlockThreads, smem, stream>>>(p);
  return cudaGetLastError();
}

// returns 1 if a register-specialised variant exists for (ks, ko)
int occu_has_specialisation(int ks, int ko, bool fp) {
  if (fp) return 0;
  return (ks == 1 && ko == 1) || (ks == 2 && ko == 1) || (ks == 5 && ko == 3);
}

template <typename T, bool STRICT>
static cudaError_t dispatch(const EvalParams& p, dim3 grid, size_t smem, cudaStream_t st, int* occ) {
  const int ks = p.L.ks, ko = p.L.ko;
  const bool fp = (p.flags & (BL_FLAG_FP_CONSTANT | BL_FLAG_FP_UNOCCUPIED)) != 0;
  constexpr int MB = 2;  // 3 blocks/SM (<= 80 registers) spills 16 words and is no faster: the ring is sized for 2
  if (fp) return launch_one<T, -1, -1, true, true, 2>(p, grid, smem, st, occ);
  if (ks == 1 && ko == 1) return launch_one<T, 1, 1, false, STRICT, MB>(p, grid, smem, st, occ);
  if (ks == 2 && ko == 1) return launch_one<T, 2, 1, false, STRICT, MB>(p, grid, smem, st, occ);
  if (ks == 5 && ko == 3) return launch_one<T, 5, 3, false, STRICT, MB>(p, grid, smem, st, occ);
  return launch_one<T, -1, -1, false, STRICT, 2>(p, grid, smem, st, occ);
}

cudaError_t launch_occu(const EvalParams& p, int dtype, dim3 grid, size_t smem, cudaStream_t stream, int* occ) {
  if (dtype == BL_F64) return dispatch<double, true>(p, grid, smem, stream, occ);
  return (p.flags & BL_FLAG_STRICT_MATH) ? dispatch<float, true>(p, grid, smem, stream, occ)
                                         : dispatch<float, false>(p, grid, smem, stream, occ);
}

cudaError_t launch_occu_summary(const EvalParams& p, int dtype, float* out, cudaStream_t st) {
  const bool fp = (p.flags & (BL_FLAG_FP_CONSTANT | BL_FLAG_FP_UNOCCUPIED)) != 0;
  if (dtype == BL_F64)
    return fp ? launch_summary<double, OccuModel<double, -1, -1, true, true>>(p, out, st)
              : launch_summary<double, OccuModel<double, -1, -1, false, true>>(p, out, st);
  return fp ? launch_summary<float, OccuModel<float, -1, -1, true, true>>(p, out, st)
            : launch_summary<float, OccuModel<float, -1, -1, false, true>>(p, out, st);
}

int occu_derived_slots(uint32_t flags) {
  return (flags & (BL_FLAG_FP_CONSTANT | BL_FLAG_FP_UNOCCUPIED)) ? 8 : 0;
}

}  // namespace bl
