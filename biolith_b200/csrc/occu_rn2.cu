// K2d: lane = chain Royle-Nichols kernel in PROBABILITY space (fp32, no false-positive constant, C >= 32, J <= 16).
//
// Reference: biolith/models/occu_rn.py:175-222 (N_i enumerated over 0..K, p_it = 1 - (1 - r_it)^N_i, masked
// Bernoulli), utils/distributions.py:31-40 (truncated-Poisson logits, normalised by CategoricalLogits); closed form
// and clamp semantics: oracle/occupancy.py:occu_rn_logp_grad.  Per unit and chain, with q_j = 1 - r_j:
//
//   weight of state k:   W_k = 2^(J_k) * prod_{j in det} clip(1 - q_j^k, tiny, 1 - eps),
//   J_k = k log2(lambda) - log2 Gamma(k+1) + sum_{j in nondet} max(k log2 q_j, log2 eps)       (clamp_probs semantics)
//   l = log sum_k W_k - log sum_k 2^(P_k),  P_k = k log2(lambda) - log2 Gamma(k+1)
//   dl/deta = E_W[k] - E_prior[k];   dl/dnu_j = +r_j sum_k k W_k q_j^k / (1 - q_j^k) [q_j^k > eps] / Z      (detection)
//                                    dl/dnu_j = -r_j sum_{k <= kc_j} k W_k / Z,  kc_j = last unclamped state   (non-det.)
//
// What makes it 3-4x cheaper than K2c (occu_rn.cu), which walked a shared-memory column of log-weights once per
// detection visit with a log and a reciprocal per (visit, state):
//   * visits are SORTED at pack time (detections first, then non-detections, then masked): lane = chain makes the
//     detection count n1 warp-uniform, so the state loop is instantiated for n1 = 0..12 with every per-detection
//     quantity (q, q^k, P_k by the all-positive recurrence P += q^k r, the gradient sum) in REGISTERS;
//   * the detection factors multiply in probability space: no log and no reciprocal per (visit, state) -- the
//     "weight without visit j" needed by the gradient is a prefix/suffix product, not a division;
//   * non-detections are linear in k until their clamp binds: the exponent J_k is built in a pre-pass (states up to
//     the first clamped one, warp-wide, cost 5 instructions; later ones 3 per visit), centred on the mode so that
//     the large terms k log2 lambda and log2 Gamma(k+1) cancel before rounding; its maximum M is the exact shift;
//   * ONE shared-memory column per thread: J_k is written by the pre-pass, read once by the main pass and
//     overwritten in place by the running sum S_k = sum_{k' <= k} k' W_k', from which every non-detection reads
//     its gradient with one load at k = kc_j.
// Exactness: a unit whose weights underflow (Z < 1e-25: only for absurd theta), or with more than 12 detections,
// takes rn2_site_exact (log space, per visit and state, the oracle's formulas) -- selected per lane / per site.
#include <cmath>
#include <cstdlib>

#include "engine.cuh"

namespace bl {

constexpr int kRn2MaxDet = 12;  // detections per unit handled in registers (more -> exact path)
constexpr int kRn2MaxKs = 8;
constexpr int kRn2MaxK = 1023;
__constant__ float c_lg2gamma_f[kRn2MaxK + 1];  // log2 Gamma(k + 1)
__constant__ float c_lgamma2_f[kRn2MaxK + 1];   // ln Gamma(k + 1) (exact path)

constexpr float kL2eLo = 1.925963033500011e-08f;  // log2(e) - (float)log2(e)
constexpr float kL2Eps = -23.0f;                  // log2(FLT_EPSILON)
constexpr float kOneMinusEps = 0.99999988079071044921875f;

// record of one unit, floats: [ X (XR) | n1, n_valid_visits, valid, 0 | JT visits x VR: w_1..w_Ko, 0.., flag ]
struct Rn2Layout {
  int ks, ko, J;
  int XR, VR, JT, R;
  int64_t n_units, n_tiles;
};

__host__ __device__ inline int rn2_jt(int J) { return J <= 8 ? 8 : (J <= 10 ? 10 : (J <= 12 ? 12 : (J <= 16 ? 16 : 0))); }

inline Rn2Layout make_rn2_layout(const Layout& L) {
  Rn2Layout s{};
  s.ks = L.ks; s.ko = L.ko; s.J = L.J;
  s.XR = (L.ks + 3) / 4 * 4;
  s.VR = L.ko + 1 <= 4 ? 4 : 8;
  s.JT = rn2_jt(L.J);
  s.R = s.XR + 4 + s.JT * s.VR;
  s.n_units = L.n_units;
  s.n_tiles = L.n_tiles;
  return s;
}

__global__ void repack_rn2_kernel(const float* __restrict__ packed, float* __restrict__ out, Layout L, Rn2Layout S) {
  const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= S.n_tiles * kWarp) return;
  float* rec = out + u * (int64_t)S.R;
  for (int i = 0; i < S.R; ++i) rec[i] = 0.f;
  if (u >= L.n_units) return;
  const float* base = packed + (u / kWarp) * (int64_t)L.F * kWarp + (u % kWarp);
  for (int k = 0; k < L.ks; ++k) rec[k] = base[k * kWarp];
  float* vis = rec + S.XR + 4;
  int slot = 0, n1 = 0;
  for (int pass = 0; pass < 2; ++pass) {  // detections first, then non-detections; masked visits stay zero records
    for (int j = 0; j < L.J; ++j) {
      const uint32_t yw = __float_as_uint(base[(L.off_y + (j >> 5)) * kWarp]);
      const uint32_t mw = __float_as_uint(base[(L.off_m + (j >> 5)) * kWarp]);
      const bool m = (mw >> (j & 31)) & 1u, y = (yw >> (j & 31)) & 1u;
      if (!m || (y != (pass == 0))) continue;
      for (int k = 0; k < L.ko; ++k) vis[slot * S.VR + k] = base[(L.off_w + j * L.ko + k) * kWarp];
      vis[slot * S.VR + S.VR - 1] = 1.f;
      ++slot;
      n1 += pass == 0;
    }
  }
  rec[S.XR] = (float)n1;
  rec[S.XR + 1] = (float)slot;
  rec[S.XR + 2] = 1.f;
}

cudaError_t launch_repack_rn2(const void* packed, void* out, const Layout& L, cudaStream_t st) {
  const Rn2Layout S = make_rn2_layout(L);
  const int64_t n = S.n_tiles * kWarp;
  if (n == 0) return cudaSuccess;
  repack_rn2_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>((const float*)packed, (float*)out, L, S);
  return cudaGetLastError();
}

size_t occu_rn2_bytes(const Layout& L) {
  const Rn2Layout S = make_rn2_layout(L);
  return (size_t)S.n_tiles * kWarp * S.R * sizeof(float);
}

template <int N> struct LoadV {
  static __device__ __forceinline__ void ld(const float* p, float (&o)[N]) {
#pragma unroll
    for (int i = 0; i < N / 4; ++i) {
      const float4 t = *reinterpret_cast<const float4*>(p + 4 * i);
      o[4 * i] = t.x; o[4 * i + 1] = t.y; o[4 * i + 2] = t.z; o[4 * i + 3] = t.w;
    }
  }
};

template <int KO> struct Rn2Out {
  float ell, geta;
  float iZ;  // 1 / Z of the fast form: the caller scales the non-detection sums S[kc_j] with it
  float ga[KO + 1];
};

// sigmoid(nu), sigmoid(-nu) and 1 + e^-|nu| with the bounded-error SFU forms (two-float log2 e, Newton on rcp)
__device__ __forceinline__ void sig_pair(float nu, float& r, float& q, float& u) {
  const float t0 = -fabsf(nu);
  const float t = sfu::ex2(fmaf(t0, kL2eLo, t0 * sfu::kLog2e));
  u = 1.0f + t;
  float inv = sfu::rcp(u);
  inv = fmaf(inv, fmaf(-u, inv, 1.0f), inv);
  const float ti = t * inv;
  r = nu >= 0.f ? inv : ti;
  q = nu >= 0.f ? ti : inv;
}

// ---- exact log-space form of one unit (rare: underflow of the probability-space weights, > 12 detections) ----
// Follows oracle/occupancy.py:occu_rn_logp_grad / _bern_terms_from_log1mP visit by visit and state by state.
template <int KO>
__device__ __noinline__ Rn2Out<KO> rn2_site_exact(const float* __restrict__ vis, int VR, int n1, int nv, float eta,
                                                 const float* __restrict__ alpha, int K, float* __restrict__ col,
                                                 int BT) {
  Rn2Out<KO> o;
  const float log_eps = Num<float>::log_eps(), log_tiny = Num<float>::log_tiny();
  // prior logits and their normaliser
  float mp = -Num<float>::inf();
  for (int k = 0; k <= K; ++k) {
    const float lk = fmaf((float)k, eta, -c_lgamma2_f[k]);
    col[(size_t)k * BT] = lk;
    mp = fmaxf(mp, lk);
  }
  float zp = 0.f, ep = 0.f;
  for (int k = 0; k <= K; ++k) {
    const float e = expf(col[(size_t)k * BT] - mp);
    zp += e;
    ep = fmaf((float)k, e, ep);
  }
  const float log_zp = mp + logf(zp);
  ep /= zp;
  for (int j = 0; j < nv; ++j) {
    const float* w = vis + j * VR;
    float nu = alpha[0];
#pragma unroll
    for (int k = 0; k < KO; ++k) nu = fmaf(w[k], alpha[1 + k], nu);
    const float uj = -(fmaxf(nu, 0.f) + log1pf(expf(-fabsf(nu))));  // log(1 - r)
    const bool det = j < n1;
    for (int k = 0; k <= K; ++k) {
      const float lq = (float)k * uj;
      const float P = -expm1f(lq);
      const bool inr = (lq > log_eps) && (P > FLT_MIN);
      float t;
      if (det) t = inr ? logf(P) : (P <= FLT_MIN ? log_tiny : Num<float>::log1m_eps());
      else t = inr ? lq : (P <= FLT_MIN ? Num<float>::neg_tiny() : log_eps);
      col[(size_t)k * BT] += t;
    }
  }
  float mx = -Num<float>::inf();
  for (int k = 0; k <= K; ++k) mx = fmaxf(mx, col[(size_t)k * BT]);
  float z = 0.f, eq = 0.f;
  for (int k = 0; k <= K; ++k) {
    const float e = expf(col[(size_t)k * BT] - mx);
    col[(size_t)k * BT] = e;
    z += e;
    eq = fmaf((float)k, e, eq);
  }
  const float iz = 1.0f / z;
  o.ell = (mx + logf(z)) - log_zp;
  o.geta = eq * iz - ep;
#pragma unroll
  for (int k = 0; k <= KO; ++k) o.ga[k] = 0.f;
  for (int j = 0; j < nv; ++j) {
    const float* w = vis + j * VR;
    float nu = alpha[0];
#pragma unroll
    for (int k = 0; k < KO; ++k) nu = fmaf(w[k], alpha[1 + k], nu);
    const float uj = -(fmaxf(nu, 0.f) + log1pf(expf(-fabsf(nu))));
    const float r = 1.0f / (1.0f + expf(-nu));
    const bool det = j < n1;
    float g = 0.f;
    for (int k = 1; k <= K; ++k) {
      const float lq = (float)k * uj;
      const float P = -expm1f(lq);
      const bool inr = (lq > log_eps) && (P > FLT_MIN);
      if (!inr) continue;
      const float dt = det ? -expf(lq) / P : 1.0f;          // dt/dlq
      g = fmaf(col[(size_t)k * BT] * (float)k, dt, g);     // dlq/dnu = -k r
    }
    const float gnu = -r * g * iz;
    o.ga[0] += gnu;
#pragma unroll
    for (int k = 0; k < KO; ++k) o.ga[1 + k] = fmaf(gnu, w[k], o.ga[1 + k]);
  }
  return o;
}

// ---- the fast form for a unit with N1 detections (compile time), JT visit slots --------------------------------
template <int KO, int JT, int N1>
__device__ __forceinline__ bool rn2_site_fast(const float* __restrict__ vis, float eta, const float (&a)[KO + 1], int K,
                                              float* __restrict__ col, int BT, const float* __restrict__ sG,
                                              Rn2Out<KO>& o) {
  constexpr int VR = KO + 1 <= 4 ? 4 : 8;
  constexpr int ND = N1 > 0 ? N1 : 1;
  float q[ND], r[ND], qk[ND], P0[ND], G[ND];
  float u2[JT];  // log2(1 - r_j) of the non-detections; 0 for detections and masked visits (neutral)
  float U2 = 0.f, umin = 0.f, rmin = 1.0f;
#pragma unroll
  for (int j = 0; j < JT; ++j) {
    float w[VR];
    LoadV<VR>::ld(vis + j * VR, w);
    float nu = a[0];
#pragma unroll
    for (int k = 0; k < KO; ++k) nu = fmaf(w[k], a[1 + k], nu);
    float rj, qj, uj;
    sig_pair(nu, rj, qj, uj);
    if (j < N1) {
      q[j] = qj; r[j] = rj; qk[j] = 1.0f; P0[j] = 0.f; G[j] = 0.f;
      u2[j] = 0.f;
      rmin = fminf(rmin, rj);
    } else {
      const float mx0 = fmaxf(nu, 0.f);
      float l2q = -(fmaf(mx0, kL2eLo, mx0 * sfu::kLog2e) + sfu::lg2(uj));  // log2(1 - r) = -softplus(nu) log2 e
      l2q = w[VR - 1] != 0.f ? l2q : 0.f;
      u2[j] = l2q;
      U2 += l2q;
      umin = fminf(umin, l2q);
    }
  }
  // rmin < 1e-30: a detection that is all but impossible -> the exact log-space form decides (checked at the end:
  // every lane has to reach the warp-wide reduction below)
  const float Kf = (float)K;
  const float eta2 = fmaf(eta, kL2eLo, eta * sfu::kLog2e);
  const float etaU = eta2 + U2;
  // centres: the (linear) joint exponent peaks near 2^etaU, the prior near lambda = 2^eta2
  const int k0 = (int)fminf(sfu::ex2(fminf(etaU, 30.f)), Kf);
  const int k0p = (int)fminf(sfu::ex2(fminf(eta2, 30.f)), Kf);
  const float k0f = (float)k0, k0pf = (float)k0p, Gk0 = sG[k0], Gk0p = sG[k0p];
  // states 0..kfree have no clamped non-detection in ANY lane of the warp
  int kfree = K;
  if (umin < 0.f) kfree = max(0, (int)fminf(kL2Eps * sfu::rcp(umin), Kf) - 1);
  kfree = __reduce_min_sync(0xffffffffu, kfree);

  // ---- pre-pass: centred joint exponent Jc_k = (k - k0) etaU - (G_k - G_k0) + clamp corrections, its maximum
  float M = -Num<float>::inf();
  {
    float kf = 0.f, dk = -k0f;
    float* ck = col;
    int k = 0;
    for (; k <= kfree; ++k) {
      const float jc = fmaf(dk, etaU, Gk0 - sG[k]);
      M = fmaxf(M, jc);
      *ck = jc;
      ck += BT; dk += 1.0f;
    }
    kf = (float)k;
    for (; k <= K; ++k) {
      float corr = 0.f;
#pragma unroll
      for (int j = N1; j < JT; ++j) corr += fmaxf(fmaf(-kf, u2[j], kL2Eps), 0.f);
      const float jc = fmaf(dk, etaU, Gk0 - sG[k]) + corr;
      M = fmaxf(M, jc);
      *ck = jc;
      ck += BT; dk += 1.0f; kf += 1.0f;
    }
  }
  // ---- main pass.  State 0 is peeled: every detection factor is clip(0) = tiny there and it carries no gradient
  // (k = 0); from k = 1 on P_k >= r_j > tiny (the caller's guard), so only the upper clamp remains in the loop.
  float Z, Zp, Ep = 0.f, S = 0.f;
  {
    float W = sfu::ex2(col[0] - M);
#pragma unroll
    for (int j = 0; j < N1; ++j) W *= FLT_MIN;
    Z = W;
    Zp = sfu::ex2(fmaf(-k0pf, eta2, Gk0p - sG[0]));
    col[0] = 0.f;
#pragma unroll
    for (int j = 0; j < N1; ++j) { P0[j] = r[j]; qk[j] = q[j]; }
    float kf = 1.0f, dkp = 1.0f - k0pf;
    float* ck = col + BT;
    for (int k = 1; k <= K; ++k) {
      const float W0 = sfu::ex2(*ck - M);
      const float pw = sfu::ex2(fmaf(dkp, eta2, Gk0p - sG[k]));
      Zp += pw;
      Ep = fmaf(kf, pw, Ep);
      W = W0;
      if constexpr (N1 > 0) {
        float F[ND], pre[ND];
#pragma unroll
        for (int j = 0; j < N1; ++j) F[j] = fminf(P0[j], kOneMinusEps);  // clip(1 - q^k, tiny, 1 - eps), k >= 1
        pre[0] = F[0];
#pragma unroll
        for (int j = 1; j < N1; ++j) pre[j] = pre[j - 1] * F[j];
        W = W0 * pre[N1 - 1];
        // weight without visit j: prefix * suffix (no division)
        float suf = kf * W0;
#pragma unroll
        for (int j = N1 - 1; j >= 0; --j) {
          const float others = j > 0 ? pre[j - 1] * suf : suf;
          const float sel = qk[j] > FLT_EPSILON ? qk[j] : 0.f;  // zero gradient where P is clamped at 1 - eps
          G[j] = fmaf(others, sel, G[j]);
          suf *= F[j];
        }
#pragma unroll
        for (int j = 0; j < N1; ++j) {
          P0[j] = fmaf(qk[j], r[j], P0[j]);  // P_{k+1} = P_k + q^k r (all positive: no 1 - q^k cancellation)
          qk[j] *= q[j];
        }
      }
      Z += W;
      S = fmaf(kf, W, S);
      *ck = S;
      ck += BT; kf += 1.0f; dkp += 1.0f;
    }
  }
  if (!(Z > 1e-25f) || !(Z < 1e30f) || !(Zp > 0.f) || rmin < 1e-30f) return false;
  const float iZ = sfu::rcp(Z), iZp = sfu::rcp(Zp);
  // constants of the two centred exponents (one double FMA pair per unit): cJ - cP
  const double cdiff = ((double)k0f * (double)etaU - (double)Gk0) - ((double)k0pf * (double)eta2 - (double)Gk0p);
  o.ell = sfu::kLn2 * (float)((double)(sfu::lg2(Z * iZp) + M) + cdiff);
  o.geta = S * iZ - Ep * iZp;
  o.iZ = iZ;
#pragma unroll
  for (int k = 0; k <= KO; ++k) o.ga[k] = 0.f;
  // detections: +r_j G_j / Z   (the non-detections read S[kc_j] from the column in the caller: rn2_nondet_grad)
#pragma unroll
  for (int j = 0; j < N1; ++j) {
    float w[VR];
    LoadV<VR>::ld(vis + j * VR, w);
    const float gnu = r[j] * G[j] * iZ;
    o.ga[0] += gnu;
#pragma unroll
    for (int k = 0; k < KO; ++k) o.ga[1 + k] = fmaf(gnu, w[k], o.ga[1 + k]);
  }
  return true;
}

// non-detections j = n1 .. nv-1 of a unit (runtime bounds, one copy of the code): dl/dnu_j = -r_j S[kc_j] / Z with
// kc_j the last state whose k log2(1 - r_j) is above log2 eps -- decided with the same fma as the pre-pass
template <int KO>
__device__ __forceinline__ void rn2_nondet_grad(const float* __restrict__ vis, int n1, int nv, const float (&a)[KO + 1],
                                                int K, const float* __restrict__ col, int BT, Rn2Out<KO>& o) {
  constexpr int VR = KO + 1 <= 4 ? 4 : 8;
  const float Kf = (float)K;
  for (int j = n1; j < nv; ++j) {
    float w[VR];
    LoadV<VR>::ld(vis + j * VR, w);
    float nu = a[0];
#pragma unroll
    for (int k = 0; k < KO; ++k) nu = fmaf(w[k], a[1 + k], nu);
    float rj, qj, uj;
    sig_pair(nu, rj, qj, uj);
    const float mx0 = fmaxf(nu, 0.f);
    const float l2q = -(fmaf(mx0, kL2eLo, mx0 * sfu::kLog2e) + sfu::lg2(uj));
    int kc = K;
    if (fmaf(-Kf, l2q, kL2Eps) >= 0.f) {  // clamped at K: find the last state with k log2 q > log2 eps
      kc = (int)fminf(kL2Eps * sfu::rcp(l2q), Kf);
      while (kc < K && fmaf(-(float)(kc + 1), l2q, kL2Eps) < 0.f) ++kc;
      while (kc > 0 && !(fmaf(-(float)kc, l2q, kL2Eps) < 0.f)) --kc;
    }
    const float gnu = -rj * col[(size_t)kc * BT] * o.iZ;
    o.ga[0] += gnu;
#pragma unroll
    for (int k = 0; k < KO; ++k) o.ga[1 + k] = fmaf(gnu, w[k], o.ga[1 + k]);
  }
}

template <int KO, int JT>
__device__ __forceinline__ bool rn2_dispatch(int n1, const float* __restrict__ vis, float eta, const float (&a)[KO + 1],
                                             int K, float* __restrict__ col, int BT, const float* __restrict__ sG,
                                             Rn2Out<KO>& o) {
  switch (n1) {  // warp-uniform
    case 0:
      if constexpr (0 <= JT) return rn2_site_fast<KO, JT, 0>(vis, eta, a, K, col, BT, sG, o);
      else return false;
    case 1:
      if constexpr (1 <= JT) return rn2_site_fast<KO, JT, 1>(vis, eta, a, K, col, BT, sG, o);
      else return false;
    case 2:
      if constexpr (2 <= JT) return rn2_site_fast<KO, JT, 2>(vis, eta, a, K, col, BT, sG, o);
      else return false;
    case 3:
      if constexpr (3 <= JT) return rn2_site_fast<KO, JT, 3>(vis, eta, a, K, col, BT, sG, o);
      else return false;
    case 4:
      if constexpr (4 <= JT) return rn2_site_fast<KO, JT, 4>(vis, eta, a, K, col, BT, sG, o);
      else return false;
    case 5:
      if constexpr (5 <= JT) return rn2_site_fast<KO, JT, 5>(vis, eta, a, K, col, BT, sG, o);
      else return false;
    case 6:
      if constexpr (6 <= JT) return rn2_site_fast<KO, JT, 6>(vis, eta, a, K, col, BT, sG, o);
      else return false;
    case 7:
      if constexpr (7 <= JT) return rn2_site_fast<KO, JT, 7>(vis, eta, a, K, col, BT, sG, o);
      else return false;
    case 8:
      if constexpr (8 <= JT) return rn2_site_fast<KO, JT, 8>(vis, eta, a, K, col, BT, sG, o);
      else return false;
    case 9:
      if constexpr (9 <= JT) return rn2_site_fast<KO, JT, 9>(vis, eta, a, K, col, BT, sG, o);
      else return false;
    case 10:
      if constexpr (10 <= JT) return rn2_site_fast<KO, JT, 10>(vis, eta, a, K, col, BT, sG, o);
      else return false;
    case 11:
      if constexpr (11 <= JT) return rn2_site_fast<KO, JT, 11>(vis, eta, a, K, col, BT, sG, o);
      else return false;
    case 12:
      if constexpr (12 <= JT) return rn2_site_fast<KO, JT, 12>(vis, eta, a, K, col, BT, sG, o);
      else return false;
    default: return false;
  }
}

template <int KO, int JT, int BT>
__global__ void __launch_bounds__(BT, BT == 128 ? 3 : 2) occu_rn2_kernel(const __grid_constant__ EvalParams p, const Rn2Layout S) {
  constexpr int KSM = kRn2MaxKs, KB = KSM + 1, KA = KO + 1, NQ = 1 + KB + KA;
  constexpr int VR = KO + 1 <= 4 ? 4 : 8;
  const int ks = S.ks, XR = S.XR, R = S.R, K = p.K;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
  float* stage0 = reinterpret_cast<float*>(smem_raw + 128);
  __shared__ int s_is_last;
  const uint32_t tile_elems = (uint32_t)R * kWarp;
  const uint32_t tile_bytes = tile_elems * sizeof(float);
  const int tid = threadIdx.x;
  double* g64 = reinterpret_cast<double*>(stage0 + (size_t)p.nstage * tile_elems) + tid;  // [NQ][BT]
  float* col = reinterpret_cast<float*>(g64 - tid + (size_t)NQ * BT) + tid;                // [K+1][BT]
  float* sG = col - tid + (size_t)(K + 1) * BT;                                            // [K+1]
  const int c0 = blockIdx.y * p.CB;
  const int ncb = min(p.CB, p.C - c0);
  const bool chain_ok = tid < ncb;
  const bool warp_on = (tid & ~31) < ncb;
  const int64_t nbt = p.n_block_tiles;
  const int64_t bt_begin = nbt * blockIdx.x / gridDim.x;
  const int64_t bt_end = nbt * (blockIdx.x + 1) / gridDim.x;
  const int n_it = (int)(bt_end - bt_begin);
  const float* packed = reinterpret_cast<const float*>(p.packed);

  if (tid == 0) {
    for (int s = 0; s < p.nstage; ++s) mbar_init(&bars[s], 1);
    fence_mbar_init();
  }
  for (int k = tid; k <= K; k += BT) sG[k] = c_lg2gamma_f[k];
  const float* th = reinterpret_cast<const float*>(p.theta) + (size_t)(c0 + (chain_ok ? tid : 0)) * p.D;
  float a[KA];
#pragma unroll
  for (int k = 0; k < KA; ++k) a[k] = th[ks + 1 + k];
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

  for (int it = 0; it < n_it; ++it) {
    const int s = it % p.nstage;
    mbar_wait(&bars[s], (uint32_t)((it / p.nstage) & 1));
    const float* tile = stage0 + (size_t)s * tile_elems;
    const int64_t unit0 = (bt_begin + it) * kWarp;
    const int n_valid = (int)max((int64_t)0, min((int64_t)kWarp, S.n_units - unit0));
    const int n_mine = warp_on ? n_valid : 0;
    for (int si = 0; si < n_mine; ++si) {
      const float* rec = tile + (size_t)si * R;
      const float4 h = *reinterpret_cast<const float4*>(rec + XR);
      const int n1 = (int)h.x, nv = (int)h.y;
      // the state loops need every register: beta is re-read (L1) and the sums go straight to the fp64 columns
      float eta = __ldg(th);
      for (int k = 0; k < ks; ++k) eta = fmaf(rec[k], __ldg(th + 1 + k), eta);
      Rn2Out<KO> o;
      const bool ok = rn2_dispatch<KO, JT>(n1, rec + XR + 4, eta, a, K, col, BT, sG, o);
      if (ok) rn2_nondet_grad<KO>(rec + XR + 4, n1, nv, a, K, col, BT, o);
      if (__any_sync(0xffffffffu, !ok)) {
        const Rn2Out<KO> ex = rn2_site_exact<KO>(rec + XR + 4, VR, n1, nv, eta, th + ks + 1, K, col, BT);
        if (!ok) o = ex;
      }
      g64[0] += (double)o.ell;
      g64[(size_t)BT] += (double)o.geta;
      for (int k = 0; k < ks; ++k) g64[(size_t)(2 + k) * BT] += (double)(o.geta * rec[k]);
#pragma unroll
      for (int k = 0; k < KA; ++k) g64[(size_t)(1 + KB + k) * BT] += (double)o.ga[k];
    }
    __syncthreads();
    if (tid == 0 && it + p.nstage < n_it) {
      mbar_expect_tx(&bars[s], tile_bytes);
      tma_load_bulk(stage0 + (size_t)s * tile_elems, packed + (size_t)(bt_begin + it + p.nstage) * tile_elems,
                    tile_bytes, &bars[s]);
    }
  }
  if (chain_ok) {
    double* my = p.partial + ((size_t)blockIdx.x * p.C + c0 + tid) * p.NQ;
    my[0] = g64[0];
#pragma unroll
    for (int k = 0; k < KB; ++k)
      if (k <= ks) my[1 + k] = g64[(size_t)(1 + k) * BT];
#pragma unroll
    for (int k = 0; k < KA; ++k) my[2 + ks + k] = g64[(size_t)(1 + KB + k) * BT];
  }
  finish_block<float>(p, c0, ncb, &s_is_last);
}

bool occu_rn2_supported(int dtype, int ks, int ko, int J, int K, uint32_t flags) {
  if (dtype != BL_F32 || (flags & (BL_FLAG_STRICT_MATH | BL_FLAG_FP_CONSTANT))) return false;
  if (const char* e = getenv("BL_RN_CHAIN_KERNEL"))  // tuning switch: 1 = keep K2c
    if (atoi(e) == 1) return false;
  return ks >= 0 && ks <= kRn2MaxKs && ko >= 1 && ko <= 4 && rn2_jt(J) > 0 && K <= kRn2MaxK;
}

size_t occu_rn2_smem(const Layout& L, int nstage, int K, int bt) {
  const Rn2Layout S = make_rn2_layout(L);
  size_t b = 128 + (size_t)nstage * S.R * kWarp * sizeof(float);
  b = (b + 15) & ~size_t(15);
  b += (size_t)(3 + kRn2MaxKs + L.ko) * bt * sizeof(double);  // fp64 sums [NQ][BT]
  b += (size_t)(K + 1) * bt * sizeof(float);                  // the per-thread state column
  return b + (size_t)(K + 1) * sizeof(float) + 16;            // log2 Gamma table
}

// threads (= chains) per block.  Measured on B200 (config 3, K = 50, 256 chains, ms per evaluation): 128 threads
// (167 registers, no spills, 3 blocks / SM) 9.08, 256 threads (128 registers, spills in the wide instantiations) 9.30
// (the 256-thread instantiations were a tuning switch only; dropped: they doubled this file's compile time)
int occu_rn2_block_threads(const Layout& L, int C, int K, size_t smem_limit) {
  (void)L; (void)C; (void)K; (void)smem_limit;
  return 128;
}

static cudaError_t ensure_rn2_tables() {
  static bool done[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 64 && done[dev]) return cudaSuccess;
  static float h2[kRn2MaxK + 1], hn[kRn2MaxK + 1];
  for (int k = 0; k <= kRn2MaxK; ++k) {
    const double lg = std::lgamma((double)k + 1.0);
    hn[k] = (float)lg;
    h2[k] = (float)(lg * 1.4426950408889634074);
  }
  cudaError_t e = cudaMemcpyToSymbol(c_lg2gamma_f, h2, sizeof(h2));
  if (e == cudaSuccess) e = cudaMemcpyToSymbol(c_lgamma2_f, hn, sizeof(hn));
  if (e == cudaSuccess && dev < 64) done[dev] = true;
  return e;
}

template <int KO, int JT, int BT>
static cudaError_t launch_rn2_one(const EvalParams& p, const Rn2Layout& S, dim3 grid, size_t smem, cudaStream_t st,
                                  int* occ) {
  auto kern = occu_rn2_kernel<KO, JT, BT>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  if (occ) return cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, kern, BT, smem);
  kern<<<grid, BT, smem, st>>>(p, S);
  return cudaGetLastError();
}

template <int KO, int JT>
static cudaError_t launch_rn2_bt(const EvalParams& p, const Rn2Layout& S, dim3 grid, size_t smem, cudaStream_t st,
                                 int* occ) {
  if (p.chain_bt != 128) return cudaErrorNotSupported;
  return launch_rn2_one<KO, JT, 128>(p, S, grid, smem, st, occ);
}

template <int KO>
static cudaError_t launch_rn2_jt(const EvalParams& p, const Rn2Layout& S, dim3 grid, size_t smem, cudaStream_t st,
                                 int* occ) {
  if (S.JT == 8) return launch_rn2_bt<KO, 8>(p, S, grid, smem, st, occ);
  if (S.JT == 10) return launch_rn2_bt<KO, 10>(p, S, grid, smem, st, occ);
  if (S.JT == 12) return launch_rn2_bt<KO, 12>(p, S, grid, smem, st, occ);
  if (S.JT == 16) return launch_rn2_bt<KO, 16>(p, S, grid, smem, st, occ);
  return cudaErrorNotSupported;
}

cudaError_t launch_occu_rn2(const EvalParams& p, dim3 grid, size_t smem, cudaStream_t st, int* occ) {
  if (!occ) {
    cudaError_t e = ensure_rn2_tables();
    if (e != cudaSuccess) return e;
  }
  const Rn2Layout S = make_rn2_layout(p.L);
  switch (p.L.ko) {
    case 1: return launch_rn2_jt<1>(p, S, grid, smem, st, occ);
    case 2: return launch_rn2_jt<2>(p, S, grid, smem, st, occ);
    case 3: return launch_rn2_jt<3>(p, S, grid, smem, st, occ);
    case 4: return launch_rn2_jt<4>(p, S, grid, smem, st, occ);
    default: return cudaErrorNotSupported;
  }
}

}  // namespace bl
