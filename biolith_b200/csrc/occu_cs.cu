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
#include <cstdlib>
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

// ------------------------------------------------------------------------------------------
// K8c: lane = chain kernel (fp32, bounded-error SFU math), the occu_cs counterpart of occu_chain.cu /
// occu_cop.cu's chain kernels: theta and the per-tile fp32 sums live in registers, the fp64 running sums
// in shared memory (one column per thread), no cross-lane reduction in the loop.  Sites are
// warp-broadcast: each thread reads a field of NS = 2 consecutive sites with one LDS.64 and runs them
// interleaved -- with three softplus / sigmoid pairs per visit that is six independent MUFU chains.  The
// mask bits of a tile are expanded to floats once per tile by the whole block.
// KS < 0: runtime number of site covariates (<= 8), accumulator slots laid out for the capacity.
// ------------------------------------------------------------------------------------------
template <int KS, int KO, int BT>
__global__ void __launch_bounds__(BT, BT == 128 ? 4 : 2) occu_cs_chain_kernel(const __grid_constant__ EvalParams p) {
  using N = Num<float>;
  constexpr int KSM = KS < 0 ? 8 : KS;
  constexpr int KB = KSM + 1, KA = KO + 1, NS = 2, NQM = 1 + KB + KA + 4, EX = 1 + KB + KA;
  const int ks = KS < 0 ? p.L.ks : KS;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
  float* stage0 = reinterpret_cast<float*>(smem_raw + 128);
  __shared__ int s_is_last;
  const int F = p.L.F, J = p.L.J, NQ = p.NQ, D = p.D;
  const uint32_t tile_elems = (uint32_t)F * kWarp;
  const uint32_t tile_bytes = tile_elems * sizeof(float);
  const int tid = threadIdx.x;
  float* mfx = stage0 + (size_t)p.nstage * tile_elems;                               // [J][32] expanded mask
  double* g64 = reinterpret_cast<double*>(mfx + (size_t)J * kWarp) + tid;             // [NQM][BT]
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
  float mu0, mu1, is0, is1, e1x, xs0, cdiff;
  {
    const float* th = reinterpret_cast<const float*>(p.theta) + (size_t)(c0 + (chain_ok ? tid : 0)) * D;
#pragma unroll
    for (int k = 0; k < KB; ++k) b[k] = (k <= ks) ? th[k] : 0.f;
#pragma unroll
    for (int k = 0; k < KA; ++k) a[k] = th[ks + 1 + k];
    const int i0 = ks + 1 + KA;
    mu0 = th[i0];
    e1x = expf(th[i0 + 1]);
    mu1 = mu0 + e1x;
    xs0 = th[i0 + 2];
    is0 = expf(-xs0);
    is1 = expf(-th[i0 + 3]);
    cdiff = xs0 - th[i0 + 3];
  }
  const float dz0 = N::log_tiny() - N::neg_tiny();
  const float cnm = xs0 + 0.91893853320467274178f;  // per unmasked visit: log sigma0 + log sqrt(2 pi)
  double logp64 = 0.0;
#pragma unroll
  for (int i = 1; i < NQM; ++i) g64[(size_t)i * BT] = 0.0;
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
    const int n_valid = (int)max((int64_t)0, min((int64_t)kWarp, p.L.n_units - unit0));
    for (int e = tid; e < J * kWarp; e += BT) {
      const int site = e & 31, j = e >> 5;
      const uint32_t mw = __float_as_uint(tile[(p.L.off_m + (j >> 5)) * kWarp + site]);
      mfx[e] = ((mw >> (j & 31)) & 1u) ? 1.f : 0.f;
    }
    __syncthreads();
    float acc[NQM];
#pragma unroll
    for (int i = 0; i < NQM; ++i) acc[i] = 0.f;
    const int n_mine = warp_on ? n_valid : 0;
    for (int g0 = 0; g0 < n_mine; g0 += NS) {
      float eta[NS], L1[NS], L0[NS], nm[NS], ga0[NS], ga[KO > 0 ? KO : 1][NS];
      float Ae0[NS], Aq0[NS], Ae1[NS], Aq1[NS], Be0[NS], Bq0[NS], Be1[NS], Bq1[NS];
#pragma unroll
      for (int i = 0; i < NS; ++i) {
        eta[i] = b[0];
        L1[i] = L0[i] = nm[i] = ga0[i] = 0.f;
        Ae0[i] = Aq0[i] = Ae1[i] = Aq1[i] = Be0[i] = Bq0[i] = Be1[i] = Bq1[i] = 0.f;
#pragma unroll
        for (int k = 0; k < KO; ++k) ga[k][i] = 0.f;
      }
#pragma unroll
      for (int k = 0; k < KSM; ++k) {
        if (k < ks) {  // re-read below for the gradient instead of held in registers
          const float2 v = *reinterpret_cast<const float2*>(tile + k * kWarp + g0);
          eta[0] = fmaf(v.x, b[1 + k], eta[0]);
          eta[1] = fmaf(v.y, b[1 + k], eta[1]);
        }
      }
#pragma unroll 2
      for (int j = 0; j < J; ++j) {
        float w[KO > 0 ? KO : 1][NS], nu[NS];
#pragma unroll
        for (int i = 0; i < NS; ++i) nu[i] = a[0];
#pragma unroll
        for (int k = 0; k < KO; ++k) {
          const float2 v = *reinterpret_cast<const float2*>(tile + (p.L.off_w + j * KO + k) * kWarp + g0);
          w[k][0] = v.x; w[k][1] = v.y;
#pragma unroll
          for (int i = 0; i < NS; ++i) nu[i] = fmaf(w[k][i], a[1 + k], nu[i]);
        }
        const float2 scv = *reinterpret_cast<const float2*>(tile + (p.L.off_y + j) * kWarp + g0);
        const float2 mfv = *reinterpret_cast<const float2*>(mfx + j * kWarp + g0);
        const float sc[NS] = {scv.x, scv.y}, mf[NS] = {mfv.x, mfv.y};
#pragma unroll
        for (int i = 0; i < NS; ++i) {
          const float e0 = (sc[i] - mu0) * is0, e1 = (sc[i] - mu1) * is1;
          const float h0 = -0.5f * e0 * e0, h1 = -0.5f * e1 * e1;
          const float delta = (h1 - h0) + cdiff;                    // n1_j - n0_j
          const sfu::SoftSig sq = sfu::softsig<true>(nu[i]);        // log q~ = xc - s, log(1-q~) = -s
          const sfu::SoftSig s1 = sfu::softsig<false>(sq.xc + delta);  // z = 1 branch
          // z = 0 branch: q~ = tiny, so softplus / sigmoid of (log tiny + delta) are below 1e-13 unless the
          // f = 1 density beats the f = 0 density by e^57 -- skipped when no lane of the warp needs them
          // (3 of the 9 MUFU per visit; the kernel is SFU-bound)
          const float d0 = dz0 + delta;
          sfu::SoftSig s0;
          s0.s = 0.f; s0.p = 0.f;
          if (__any_sync(0xffffffffu, d0 > -30.f)) s0 = sfu::softsig<false>(d0);
          L1[i] = fmaf(mf[i], (s1.s - sq.s) + h0, L1[i]);
          L0[i] = fmaf(mf[i], (N::neg_tiny() + s0.s) + h0, L0[i]);
          nm[i] += mf[i];
          const float g = sq.inr ? mf[i] * (s1.p - sq.p) : 0.f;
          ga0[i] += g;
#pragma unroll
          for (int k = 0; k < KO; ++k) ga[k][i] = fmaf(g, w[k][i], ga[k][i]);
          const float q0 = fmaf(e0, e0, -1.f), q1 = fmaf(e1, e1, -1.f);
          const float w1m = mf[i] * s1.p, w0m = mf[i] - w1m, v1m = mf[i] * s0.p, v0m = mf[i] - v1m;
          Ae0[i] = fmaf(w0m, e0, Ae0[i]); Aq0[i] = fmaf(w0m, q0, Aq0[i]);
          Ae1[i] = fmaf(w1m, e1, Ae1[i]); Aq1[i] = fmaf(w1m, q1, Aq1[i]);
          Be0[i] = fmaf(v0m, e0, Be0[i]); Bq0[i] = fmaf(v0m, q0, Bq0[i]);
          Be1[i] = fmaf(v1m, e1, Be1[i]); Bq1[i] = fmaf(v1m, q1, Bq1[i]);
        }
      }
      float geta[NS];
#pragma unroll
      for (int i = 0; i < NS; ++i) {
        const float vf = (g0 + i < n_valid) ? 1.f : 0.f;
        const sfu::SoftSig se = sfu::softsig<true>(eta[i]);
        const float av = (se.xc - se.s) + L1[i];
        const float bv = L0[i] - se.s;
        const sfu::SoftSig sd = sfu::softsig<false>(av - bv);
        const float ell = (bv + sd.s) - nm[i] * cnm;
        const float r = sd.p * vf, r0 = (1.f - sd.p) * vf;
        geta[i] = se.inr ? (sd.p - se.p) * vf : 0.f;
        logp64 += (double)(ell * vf);
        acc[1] += geta[i];
        acc[1 + KB] = fmaf(r, ga0[i], acc[1 + KB]);
#pragma unroll
        for (int k = 0; k < KO; ++k) acc[2 + KB + k] = fmaf(r, ga[k][i], acc[2 + KB + k]);
        const float g_mu0 = fmaf(r, Ae0[i], r0 * Be0[i]) * is0;
        const float g_mu1 = fmaf(r, Ae1[i], r0 * Be1[i]) * is1;
        acc[EX + 0] += g_mu0 + g_mu1;              // mu1 = mu0 + exp(x1)
        acc[EX + 1] = fmaf(g_mu1, e1x, acc[EX + 1]);
        acc[EX + 2] += fmaf(r, Aq0[i], r0 * Bq0[i]);
        acc[EX + 3] += fmaf(r, Aq1[i], r0 * Bq1[i]);
      }
#pragma unroll
      for (int k = 0; k < KSM; ++k) {
        if (k < ks) {
          const float2 v = *reinterpret_cast<const float2*>(tile + k * kWarp + g0);
          acc[2 + k] = fmaf(geta[0], v.x, acc[2 + k]);
          acc[2 + k] = fmaf(geta[1], v.y, acc[2 + k]);
        }
      }
    }
#pragma unroll
    for (int i = 1; i < NQM; ++i) g64[(size_t)i * BT] += (double)acc[i];
    __syncthreads();
    if (tid == 0 && it + p.nstage < n_it) {
      mbar_expect_tx(&bars[s], tile_bytes);
      tma_load_bulk(stage0 + (size_t)s * tile_elems, packed + (size_t)(bt_begin + it + p.nstage) * tile_elems,
                    tile_bytes, &bars[s]);
    }
  }
  if (chain_ok) {
    double* my = p.partial + ((size_t)blockIdx.x * p.C + c0 + tid) * NQ;
    my[0] = logp64;
#pragma unroll
    for (int kk = 0; kk < KB; ++kk)
      if (kk <= ks) my[1 + kk] = g64[(size_t)(1 + kk) * BT];
#pragma unroll
    for (int kk = 0; kk < KA; ++kk) my[2 + ks + kk] = g64[(size_t)(1 + KB + kk) * BT];
#pragma unroll
    for (int e = 0; e < 4; ++e) my[2 + ks + KA + e] = g64[(size_t)(EX + e) * BT];
  }
  finish_block<float>(p, c0, ncb, &s_is_last);
}

bool occu_cs_chain_supported(int dtype, int ks, int ko, uint32_t flags) {
  if (dtype != BL_F32 || (flags & BL_FLAG_STRICT_MATH)) return false;
  if (const char* e = getenv("BL_CS_CHAIN"))
    if (atoi(e) == 0) return false;  // tuning switch: 0 forces the site-parallel engine
  return ks >= 0 && ks <= 8 && ko >= 1 && ko <= 4;
}

// threads (= chains) per block for a batch of C chains; see occu_chain.cu:occu_chain_block_threads
int occu_cs_chain_block_threads(int C) { return (C > 0 && C % 256 == 0) ? 256 : 128; }

size_t occu_cs_chain_smem(const Layout& L, int nstage, int bt) {
  size_t bts = 128 + (size_t)nstage * L.F * kWarp * sizeof(float) + (size_t)L.J * kWarp * sizeof(float);
  bts = (bts + 15) & ~size_t(15);
  return bts + (size_t)(1 + 9 + L.ko + 1 + 4) * bt * sizeof(double);
}

template <int KS, int KO, int BT>
static cudaError_t launch_cs_chain_bt(const EvalParams& p, dim3 grid, size_t smem, cudaStream_t st, int* occ) {
  auto kern = occu_cs_chain_kernel<KS, KO, BT>;
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

template <int KS, int KO>
static cudaError_t launch_cs_chain_one(const EvalParams& p, dim3 grid, size_t smem, cudaStream_t st, int* occ) {
  if (p.chain_bt == 128) return launch_cs_chain_bt<KS, KO, 128>(p, grid, smem, st, occ);
  return launch_cs_chain_bt<KS, KO, 256>(p, grid, smem, st, occ);
}

cudaError_t launch_occu_cs_chain(const EvalParams& p, dim3 grid, size_t smem, cudaStream_t st, int* occ) {
  if (p.L.ks == 5 && p.L.ko == 3) return launch_cs_chain_one<5, 3>(p, grid, smem, st, occ);
  if (p.L.ks >= 0 && p.L.ks <= 8) {  // runtime Ks
    if (p.L.ko == 1) return launch_cs_chain_one<-1, 1>(p, grid, smem, st, occ);
    if (p.L.ko == 2) return launch_cs_chain_one<-1, 2>(p, grid, smem, st, occ);
    if (p.L.ko == 3) return launch_cs_chain_one<-1, 3>(p, grid, smem, st, occ);
    if (p.L.ko == 4) return launch_cs_chain_one<-1, 4>(p, grid, smem, st, occ);
  }
  return cudaErrorNotSupported;
}

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
