// K1d: lane = chain occu kernel over SIGNED site records (fp32, no false-positive extras, C >= 32).
//
// Same mapping as K1c (occu_chain.cu): lane = chain, every lane reads the same site (broadcast LDS.128),
// block = <= 256 chains x a contiguous range of 32-site tiles staged by TMA (cp.async.bulk) through an
// mbarrier ring, fp32 inside a tile, fp64 across tiles / blocks, ticketed last-block reduction.
// What changes is the arithmetic per visit (reference: biolith/models/occu.py:221-242 with p_fp = z p):
//
//   sgn_j = +1 (detection), -1 (non-detection), 0 (masked);  v_j = sgn_j * [1, W_j]  (packed once per fit)
//   x'_j  = v_j . alpha                      the Bernoulli log-lik of visit j is  log sigmoid(x'_j) = -log(1 + e_j),
//   e_j   = exp(-x'_j)                       and d/d alpha = q_j v_j with q_j = e_j / (1 + e_j)
//
//   * ONE ex2 per visit: the covariate slots of a record hold sgn W * (-log2 e), rounded per element at pack time, and
//     the intercept enters as sgn * (A_hi + A_lo) with A = -log2(e) alpha_0 split into two floats per chain, so
//     x2 = -log2(e) x' feeds ex2 directly.  (Near the posterior mode |g| << |H| |theta|: evaluating the density at a
//     theta that is off by a FIXED relative 1e-8 -- a rounded log2 e, a rounded alpha_0 log2 e -- is by itself a
//     gradient error of 4e-6 .. 1.5e-5 of |g|inf at config 2, measured; per-element rounding of the data is
//     unbiased noise and does not add up.)
//   * product-log:  sum_j log(1 + e_j) = log prod_j (1 + e_j)   -> one lg2 per 4 visits;
//   * batch inversion: 1 / (1 + e_j) for 4 visits from ONE rcp of their product (+ Newton) and 8 multiplies;
//   * masked / padded visits have v = 0 -> e = 1, 1 + e = 2 exactly -> they add exactly 1 to the lg2 sum, which the
//     per-site count `cnt` removes again, and nothing to the gradient: no mask handling in the loop.
//   => 17 MUFU and ~190 instructions per (site, chain) at J = 8, Ko = 3 instead of 30 and 270 (K1c).
//
// numpyro's clamp_probs (p~ = clip(p, tiny, 1 - eps), zero gradient outside) is kept EXACTLY: the fast form is only
// valid while no visit is clamped at the low side, i.e. while every 1 + e_j < 2^23 (x > log((1-eps)/eps) for a
// non-detection; the bound is conservative for detections, whose clamp sits at log tiny).  Each lane tracks the
// largest pair product (>= every 1 + e_j of the pair); lanes over the bound take the per-visit clamped form of
// K1c for that site (slow_visits: bit-identical arithmetic to sfu::softsig<true>), selected per lane, so a chain's
// result never depends on its neighbours.  The high-side clamp changes a visit by < 1.2e-7 absolute in value and
// gradient (log(1 - eps) vs -log(1 + e), e < eps) -- below the fp32 rounding of the O(1) terms it is added to.
// The site-level terms (psi, logaddexp over z) keep sfu::softsig<true> as in K1c.
#include <cstdlib>

#include "engine.cuh"

namespace bl {

constexpr int kSignedMaxKs = 8;

constexpr double kLog2eD = 1.4426950408889634074, kLn2D = 0.6931471805599453094;
constexpr float kLog2eLo = 1.925963033500011e-08f;  // log2(e) - (float)log2(e): second term of the two-float constant
constexpr float kLn2Lo = -1.9046542121259336e-09f;  // ln(2)   - (float)ln(2)

// MUFU.EX2 on B200 has a mean relative error of -5e-8 for negative arguments (profiles/r02_mufu_error.txt), and near
// the posterior mode that bias IS most of the gradient error (scripts/ex2_bias_emulation.py).  Shifting the argument by
// +5e-8 / ln 2 compensates it where it matters; BL_SIGNED_EX2_SHIFT=0 at build time (-DBL_SIGNED_EX2_SHIFT=0) removes it.
#ifndef BL_SIGNED_EX2_SHIFT
#define BL_SIGNED_EX2_SHIFT 1
#endif
constexpr float kEx2Shift = BL_SIGNED_EX2_SHIFT ? 7.2134752e-08f : 0.0f;

// Elementary functions of K1d.  Default: the bounded-error SFU forms (ex2.approx, lg2.approx, rcp.approx + Newton).
// STRICT (BL_FLAG_STRICT_MATH): an FMA-pipe exp2 (below), libm log2f and IEEE division -- north_star's "fast-math-free" clause -- in the
// SAME formulation (one exponential per visit, product-log, batch inversion are algebra, not approximations), so the
// conformant path costs ~100 instructions per (site, chain) more instead of running the 4 x slower engine.
// 2^x on the FMA pipe: n = rint(x), 2^(x - n) by its degree-7 Taylor polynomial on [-1/2, 1/2] (truncation 5e-9,
// seven FMAs: ~1 ulp, and -- unlike MUFU.EX2, which libm's exp2f also ends in -- no systematic error: near the
// posterior mode a 5e-8 mean relative error of the exponential IS the gradient error, scripts/ex2_bias_emulation.py).
__device__ __forceinline__ float exp2_fma(float x) {
  x = fminf(fmaxf(x, -126.0f), 126.0f);
  const float n = rintf(x);
  const float f = x - n;
  float p = 1.525273380405984e-05f;
  p = fmaf(p, f, 1.5403530393381608e-04f);
  p = fmaf(p, f, 1.3333558146428443e-03f);
  p = fmaf(p, f, 9.618129107628477e-03f);
  p = fmaf(p, f, 5.550410866482158e-02f);
  p = fmaf(p, f, 2.402265069591007e-01f);
  p = fmaf(p, f, 6.931471805599453e-01f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + ((int)n << 23));  // p in [0.70, 1.42]: the exponent add cannot carry out
}

template <bool STRICT> struct SMath {
  static __device__ __forceinline__ float ex2(float x) {
    if constexpr (STRICT) return exp2_fma(x); else return sfu::ex2(x);
  }
  static __device__ __forceinline__ float lg2(float x) {
    if constexpr (STRICT) return log2f(x); else return sfu::lg2(x);
  }
  static __device__ __forceinline__ float inv(float x) {
    if constexpr (STRICT) {
      return 1.0f / x;
    } else {
      const float r = sfu::rcp(x);
      return fmaf(r, fmaf(-x, r, 1.0f), r);  // Newton: MUFU.RCP's bias would add up over 10^7 visits
    }
  }
};

// 2^(t log2 e) with the two-float constant: the argument carries no fixed relative error
template <bool STRICT>
__device__ __forceinline__ float exp_neg_abs(float x) {
  const float t = -fabsf(x);
  return SMath<STRICT>::ex2(fmaf(t, kLog2eLo, t * sfu::kLog2e));
}


// record of one unit, floats:  [ X (XR = roundup(Ks,4)) | n1, cnt, valid, 0 | NQ quads x 4 visits x VR ]
struct SignedLayout {
  int ks, ko, J;
  int XR, VR, nq, R;
  int G;  // warp-tiles (32 sites each) staged per TMA / ring slot: fewer block barriers per site
  int64_t n_units, n_tiles;
};

// Sites staged per ring slot.  Measured on B200 (config 2, ms per 1024-chain evaluation; block barrier per slot):
// 32 sites 7.57 | 64: 7.39 | 128: 7.32 | 256: 8.30 (ring too shallow) -> the largest group whose slot stays <= 24 KB.
static int signed_group(int R) {
  if (const char* e = getenv("BL_SIGNED_G")) {  // tuning switch, read when a launch is configured
    const int g = atoi(e);
    if (g == 1 || g == 2 || g == 4 || g == 8) return g;
  }
  int g = 4;
  while (g > 1 && (size_t)R * kWarp * g * sizeof(float) > 24576) g /= 2;
  return g;
}

__host__ __device__ inline int signed_vr(int ko) { return ko + 1 <= 2 ? 2 : (ko + 1 <= 4 ? 4 : 8); }

inline SignedLayout make_signed_layout(const Layout& L) {
  SignedLayout s{};
  s.ks = L.ks; s.ko = L.ko; s.J = L.J;
  s.XR = (L.ks + 3) / 4 * 4;
  s.VR = signed_vr(L.ko);
  s.nq = (L.J + 3) / 4;
  s.R = s.XR + 4 + s.nq * 4 * s.VR;
  s.n_units = L.n_units;
  s.n_tiles = L.n_tiles_padded;  // padded to 8 warp-tiles, so any group size up to 8 stays in bounds
  s.G = signed_group(s.R);
  return s;
}

__global__ void repack_signed_kernel(const float* __restrict__ packed, float* __restrict__ out, Layout L,
                                     SignedLayout S) {
  const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= S.n_tiles * kWarp) return;
  float* rec = out + u * (int64_t)S.R;
  for (int i = 0; i < S.R; ++i) rec[i] = 0.f;
  float* hdr = rec + S.XR;
  if (u >= L.n_units) {  // padding unit of the last tile: weight 0, every visit "masked"
    hdr[1] = (float)(4 * S.nq);
    return;
  }
  const float* base = packed + (u / kWarp) * (int64_t)L.F * kWarp + (u % kWarp);
  for (int k = 0; k < L.ks; ++k) rec[k] = base[k * kWarp];
  int cnt = 4 * S.nq - L.J;
  float* vis = hdr + 4;
  for (int j = 0; j < L.J; ++j) {
    const uint32_t yw = __float_as_uint(base[(L.off_y + (j >> 5)) * kWarp]);
    const uint32_t mw = __float_as_uint(base[(L.off_m + (j >> 5)) * kWarp]);
    const bool m = (mw >> (j & 31)) & 1u, y = (yw >> (j & 31)) & 1u;
    const float sgn = m ? (y ? 1.f : -1.f) : 0.f;
    cnt += m ? 0 : 1;
    // slot 0: sgn (exact); covariates: sgn W * (-log2 e), rounded once per element (unbiased)
    vis[j * S.VR] = sgn;
    for (int k = 0; k < L.ko; ++k)
      vis[j * S.VR + 1 + k] = (float)(-kLog2eD * (double)(sgn * base[(L.off_w + j * L.ko + k) * kWarp]));
  }
  hdr[0] = base[L.off_n1 * kWarp];
  hdr[1] = (float)cnt;
  hdr[2] = 1.f;
}

cudaError_t launch_repack_signed(const void* packed, void* out, const Layout& L, cudaStream_t st) {
  const SignedLayout S = make_signed_layout(L);
  const int64_t n = S.n_tiles * kWarp;
  if (n == 0) return cudaSuccess;
  repack_signed_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>((const float*)packed, (float*)out, L, S);
  return cudaGetLastError();
}

size_t occu_signed_bytes(const Layout& L) {
  const SignedLayout S = make_signed_layout(L);
  return (size_t)S.n_tiles * kWarp * S.R * sizeof(float);
}

// ---- exact per-visit form for lanes with a clamped visit (rare): K1c's arithmetic on the signed record ----------
template <int KO> struct SlowOut { float L1; float ga[KO + 1]; };

template <int KO, bool STRICT>
__device__ __noinline__ SlowOut<KO> slow_visits(const float* __restrict__ vis, int nvis,
                                               const float* __restrict__ alpha) {
  constexpr int VR = KO + 1 <= 2 ? 2 : (KO + 1 <= 4 ? 4 : 8);
  SlowOut<KO> o;
  o.L1 = 0.f;
#pragma unroll
  for (int k = 0; k <= KO; ++k) o.ga[k] = 0.f;
  for (int j = 0; j < nvis; ++j) {
    const float* v = vis + j * VR;  // = [sgn, -log2(e) sgn W_j]
    const float sgn = v[0];
    if (sgn == 0.f) continue;
    float x2 = 0.f;
#pragma unroll
    for (int k = 0; k < KO; ++k) x2 = fmaf(v[1 + k], alpha[1 + k], x2);
    const float xp = fmaf(sgn, alpha[0], -fmaf(x2, kLn2Lo, x2 * sfu::kLn2));  // x' = sgn (alpha0 + W . alpha)
    const float x = sgn * xp;
    const float yf = sgn > 0.f ? 1.f : 0.f;
    float g;
    if constexpr (STRICT) {  // the engine's libm form (common.cuh log_sigmoid_pair)
      const LogSig<float> ls = log_sigmoid_pair<float>(x);
      o.L1 += sgn > 0.f ? ls.lp : ls.l1mp;
      g = ls.inr ? (sgn > 0.f ? ls.q : -ls.p) : 0.f;
    } else {
      const sfu::SoftSig ss = sfu::softsig<true>(x);
      o.L1 += fmaf(yf, ss.xc, -ss.s);
      g = ss.inr ? (yf - ss.p) : 0.f;
    }
    const float gs = g * sgn;  // d/d alpha_k = g W_k; kept in the record's units: (g sgn) v_k = -log2(e) g W_k, k >= 1
#pragma unroll
    for (int k = 0; k <= KO; ++k) o.ga[k] = fmaf(gs, v[k], o.ga[k]);
  }
  return o;
}

constexpr float kClampProduct = 8388608.0f;  // 2^23 > (1 - eps) / eps: a pair product below it has no clamped visit

template <int N> struct LoadVec {
  static __device__ __forceinline__ void ld(const float* p, float (&o)[N]) {
    static_assert(N % 4 == 0 || N == 2, "records are float4 / float2 multiples");
    if constexpr (N == 2) {
      const float2 t = *reinterpret_cast<const float2*>(p);
      o[0] = t.x; o[1] = t.y;
    } else {
#pragma unroll
      for (int i = 0; i < N / 4; ++i) {
        const float4 t = *reinterpret_cast<const float4*>(p + 4 * i);
        o[4 * i] = t.x; o[4 * i + 1] = t.y; o[4 * i + 2] = t.z; o[4 * i + 3] = t.w;
      }
    }
  }
};

// NV = 4 or 8 visits of one site: x2 = v . a2, e = 2^x2, u = 1 + e; ONE lg2 and ONE rcp (+ Newton) for the product of
// all NV factors (batch inversion through the pair-product tree), q_j = e_j / u_j, ga += q_j v_j.  `mx` tracks the
// largest pair product: while it stays below 2^23 no visit is clamped and the product of 8 factors is < 2^92.
template <int KO, int NV, bool STRICT>
__device__ __forceinline__ void visit_block(const float* __restrict__ vp, const float (&a2)[KO + 2], float& lgsum,
                                            float& mx, float (&ga)[KO + 1]) {
  using M = SMath<STRICT>;
  constexpr int VR = KO + 1 <= 2 ? 2 : (KO + 1 <= 4 ? 4 : 8);
  static_assert(NV == 4 || NV == 8, "visits are processed in quads or octets");
  float v[NV * VR];
  LoadVec<NV * VR>::ld(vp, v);
  float e[NV], u[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    // sgn * A_lo (+ a sub-ulp offset that cancels MUFU.EX2's mean error at the negative arguments that dominate; it
    // survives the roundings of the chain below statistically, like A_lo itself -- see kEx2Shift)
    float x2 = STRICT ? v[j * VR] * a2[KO + 1] : fmaf(v[j * VR], a2[KO + 1], kEx2Shift);
#pragma unroll
    for (int k = 0; k < KO; ++k) x2 = fmaf(v[j * VR + 1 + k], a2[1 + k], x2);
    x2 = fmaf(v[j * VR], a2[0], x2);    // + sgn * A_hi
    e[j] = M::ex2(x2);
    u[j] = 1.0f + e[j];
  }
  float pr[NV / 2], rp[NV / 2];  // pair products and their reciprocals
#pragma unroll
  for (int i = 0; i < NV / 2; ++i) pr[i] = u[2 * i] * u[2 * i + 1];
  if constexpr (NV == 4) {
    const float pp = pr[0] * pr[1];
    mx = fmaxf(mx, fmaxf(pr[0], pr[1]));
    const float rinv = M::inv(pp);
    lgsum += M::lg2(pp);
    rp[0] = rinv * pr[1];
    rp[1] = rinv * pr[0];
  } else {
    const float pa = pr[0] * pr[1], pb = pr[2] * pr[3], pp = pa * pb;
    mx = fmaxf(fmaxf(mx, fmaxf(pr[0], pr[1])), fmaxf(pr[2], pr[3]));
    const float rinv = M::inv(pp);
    lgsum += M::lg2(pp);
    const float ra = rinv * pb, rb = rinv * pa;  // 1 / (u0 u1 u2 u3), 1 / (u4 u5 u6 u7)
    rp[0] = ra * pr[1]; rp[1] = ra * pr[0];
    rp[2] = rb * pr[3]; rp[3] = rb * pr[2];
  }
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const float qj = e[j] * (rp[j / 2] * u[j ^ 1]);  // e_j / u_j
#pragma unroll
    for (int k = 0; k <= KO; ++k) ga[k] = fmaf(qj, v[j * VR + k], ga[k]);
  }
}

// KS < 0: runtime Ks (<= kSignedMaxKs); NQD = 0: runtime number of visit quads.
// Tried and rejected (measured, config 2): a dedicated producer warp with full / empty mbarrier pairs instead of
// the block barrier per ring slot (288 threads -> 96..112 registers): 7.76 ms against 7.32 ms; three resident
// blocks per SM (72..80 registers, spills): 8.2 ms.
template <int KS, int KO, int NQD, int NS, int MINB, int BT, bool STRICT>
__global__ void __launch_bounds__(BT, MINB) occu_signed_kernel(const __grid_constant__ EvalParams p, const SignedLayout S) {
  using M = SMath<STRICT>;
  constexpr int KSM = KS < 0 ? kSignedMaxKs : KS;
  constexpr int KB = KSM + 1, KA = KO + 1, NQ = 1 + KB + KA;
  constexpr int VR = KO + 1 <= 2 ? 2 : (KO + 1 <= 4 ? 4 : 8);
  constexpr int XRC = (KSM + 3) / 4 * 4;
  const int ks = KS < 0 ? S.ks : KS;
  const int XR = KS < 0 ? S.XR : XRC;
  const int nq = NQD > 0 ? NQD : S.nq;
  const int R = (KS >= 0 && NQD > 0) ? (XRC + 4 + NQD * 4 * VR) : S.R;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
  float* stage0 = reinterpret_cast<float*>(smem_raw + 128);
  __shared__ int s_is_last;
  const int TS = S.G * kWarp;  // sites per stage
  const uint32_t tile_elems = (uint32_t)R * TS;
  const uint32_t tile_bytes = tile_elems * sizeof(float);
  const int tid = threadIdx.x;
  const int c0 = blockIdx.y * p.CB;
  const int ncb = min(p.CB, p.C - c0);
  const bool chain_ok = tid < ncb;
  const bool warp_on = (tid & ~31) < ncb;
  const int64_t nbt = p.n_block_tiles;
  const int64_t bt_begin = nbt * blockIdx.x / gridDim.x;
  const int64_t bt_end = nbt * (blockIdx.x + 1) / gridDim.x;
  const int n_it = (int)(bt_end - bt_begin);
  const float* packed = reinterpret_cast<const float*>(p.packed);

  // ring_mode 1: no block barrier per slot -- every warp ARRIVES on the slot's `empty` mbarrier when it is done with
  // it and runs on; thread 0 re-arms the slot two iterations later (it only waits if a warp lags two slots behind)
  uint64_t* empty = bars + kMaxStages;
  const bool free_ring = p.ring_mode == 1 && p.nstage >= 3;
  if (tid == 0) {
    for (int s = 0; s < p.nstage; ++s) {
      mbar_init(&bars[s], 1);
      mbar_init(&empty[s], BT / kWarp);
    }
    fence_mbar_init();
  }
  const float* th = reinterpret_cast<const float*>(p.theta) + (size_t)(c0 + (chain_ok ? tid : 0)) * p.D;
  float b[KB], a2[KA + 1];
#pragma unroll
  for (int k = 0; k < KB; ++k) b[k] = (k <= ks) ? th[k] : 0.f;
#pragma unroll
  for (int k = 1; k < KA; ++k) a2[k] = th[ks + 1 + k];  // covariate slots of the records are pre-scaled by -log2(e)
  {
    const double A = -kLog2eD * (double)th[ks + 1];  // intercept: two floats, no per-chain rounding of -log2(e) alpha_0
    a2[0] = (float)A;
    a2[KA] = (float)(A - (double)a2[0]);
  }
  const float* th_alpha = th + ks + 1;
  double* g64 = reinterpret_cast<double*>(stage0 + (size_t)p.nstage * tile_elems) + tid;
  double logp64 = 0.0;
#pragma unroll
  for (int i = 1; i < NQ; ++i) g64[(size_t)i * BT] = 0.0;
  __syncthreads();
  if (tid == 0) {
    const int pre = min(p.nstage, n_it);
    for (int s = 0; s < pre; ++s) {
      mbar_expect_tx(&bars[s], tile_bytes);
      tma_load_bulk(stage0 + (size_t)s * tile_elems, packed + (size_t)(bt_begin + s) * tile_elems, tile_bytes,
                    &bars[s]);
    }
  }
  const float log_tiny = Num<float>::log_tiny();

  for (int it = 0; it < n_it; ++it) {
    const int s = it % p.nstage;
    if (free_ring && tid == 0 && it >= 2 && it - 2 + p.nstage < n_it) {
      const int ps = (it - 2) % p.nstage;
      mbar_wait(&empty[ps], (uint32_t)(((it - 2) / p.nstage) & 1));
      mbar_expect_tx(&bars[ps], tile_bytes);
      tma_load_bulk(stage0 + (size_t)ps * tile_elems, packed + (size_t)(bt_begin + it - 2 + p.nstage) * tile_elems,
                    tile_bytes, &bars[ps]);
    }
    mbar_wait(&bars[s], (uint32_t)((it / p.nstage) & 1));
    const float* tile = stage0 + (size_t)s * tile_elems;
    const int64_t unit0 = (bt_begin + it) * TS;
    const int n_valid = (int)max((int64_t)0, min((int64_t)TS, S.n_units - unit0));
    float acc[NQ];
#pragma unroll
    for (int i = 0; i < NQ; ++i) acc[i] = 0.f;
    if (warp_on) {
      for (int g0 = 0; g0 < n_valid; g0 += NS) {
        const float* rec = tile + (size_t)g0 * R;
        float lgsum[NS], mx[NS], ga[NS][KA];
#pragma unroll
        for (int i = 0; i < NS; ++i) {
          lgsum[i] = 0.f; mx[i] = 0.f;
#pragma unroll
          for (int k = 0; k < KA; ++k) ga[i][k] = 0.f;
        }
        if constexpr (NQD == 2) {  // 8 visits: one octet per site
#pragma unroll
          for (int i = 0; i < NS; ++i)
            visit_block<KO, 8, STRICT>(rec + (size_t)i * R + XR + 4, a2, lgsum[i], mx[i], ga[i]);
        } else {
          int q = 0;
          for (; q + 1 < nq; q += 2) {
#pragma unroll
            for (int i = 0; i < NS; ++i)
              visit_block<KO, 8, STRICT>(rec + (size_t)i * R + XR + 4 + q * 4 * VR, a2, lgsum[i], mx[i], ga[i]);
          }
          if (q < nq) {
#pragma unroll
            for (int i = 0; i < NS; ++i)
              visit_block<KO, 4, STRICT>(rec + (size_t)i * R + XR + 4 + q * 4 * VR, a2, lgsum[i], mx[i], ga[i]);
          }
        }
        float L1[NS], n1[NS], vf[NS];
        bool slow = false;
#pragma unroll
        for (int i = 0; i < NS; ++i) {
          const float4 h = *reinterpret_cast<const float4*>(rec + (size_t)i * R + XR);
          n1[i] = h.x; vf[i] = h.z;
          L1[i] = -sfu::kLn2 * (lgsum[i] - h.y);
          slow |= mx[i] >= kClampProduct;
        }
        if (__any_sync(0xffffffffu, slow)) {
#pragma unroll
          for (int i = 0; i < NS; ++i) {
            const SlowOut<KO> so = slow_visits<KO, STRICT>(rec + (size_t)i * R + XR + 4, 4 * nq, th_alpha);
            if (mx[i] >= kClampProduct) {
              L1[i] = so.L1;
#pragma unroll
              for (int k = 0; k < KA; ++k) ga[i][k] = so.ga[k];
            }
          }
        }
        float eta[NS];
#pragma unroll
        for (int i = 0; i < NS; ++i) eta[i] = b[0];
#pragma unroll
        for (int k4 = 0; k4 < XRC; k4 += 4) {
          if (k4 < ks) {
#pragma unroll
            for (int i = 0; i < NS; ++i) {
              const float4 x = *reinterpret_cast<const float4*>(rec + (size_t)i * R + k4);
              eta[i] = fmaf(x.x, b[1 + k4], eta[i]);
              if (k4 + 1 < KSM) eta[i] = fmaf(x.y, b[2 + k4 < KB ? 2 + k4 : 0], eta[i]);
              if (k4 + 2 < KSM) eta[i] = fmaf(x.z, b[3 + k4 < KB ? 3 + k4 : 0], eta[i]);
              if (k4 + 3 < KSM) eta[i] = fmaf(x.w, b[4 + k4 < KB ? 4 + k4 : 0], eta[i]);
            }
          }
        }
        float geta[NS];
#pragma unroll
        for (int i = 0; i < NS; ++i) {
          // psi~ = sigmoid(xc) with xc the clamped eta (numpyro's clamp_probs in logit space, see sfu::softsig);
          // a = log psi~ + L1, b = log(1 - psi~) + n1 log tiny share the term -softplus(xc) = -max(xc,0) - log u_e, so
          //   d = a - b = xc + L1 - n1 log tiny,   logaddexp(a, b) = max(d, 0) + n1 log tiny - max(xc, 0) + log(u_d / u_e)
          const float xc = fminf(fmaxf(eta[i], sfu::kXLo), sfu::kXHi);
          const bool inr = xc == eta[i];
          const float te = exp_neg_abs<STRICT>(xc);
          const float ue = 1.0f + te;
          const float inve = M::inv(ue);
          const float psi = (xc >= 0.f) ? inve : te * inve;
          const float bl = n1[i] * log_tiny;
          const float al = xc + L1[i];
          const float d = al - bl;
          const float td = exp_neg_abs<STRICT>(d);
          const float ud = 1.0f + td;
          const float invd = M::inv(ud);
          const float rr = (d >= 0.f) ? invd : td * invd;  // P(z = 1 | y)
          const float r = rr * vf[i];
          // max(a, b) picked by select, not as b + max(d, 0): n1 log tiny ~ -87 n1 would cost the sum its low bits
          const float ell = (fmaf(M::lg2(ud * inve), sfu::kLn2, (d >= 0.f ? al : bl) - fmaxf(xc, 0.f))) * vf[i];
          geta[i] = inr ? (rr - psi) * vf[i] : 0.f;
          logp64 += (double)ell;  // fp64 per unit: NUTS needs energy *differences* of a ~1e6-sized sum
          acc[1] += geta[i];
#pragma unroll
          for (int k = 0; k < KA; ++k) acc[1 + KB + k] = fmaf(r, ga[i][k], acc[1 + KB + k]);
        }
#pragma unroll
        for (int k4 = 0; k4 < XRC; k4 += 4) {
          if (k4 < ks) {
#pragma unroll
            for (int i = 0; i < NS; ++i) {
              const float4 x = *reinterpret_cast<const float4*>(rec + (size_t)i * R + k4);
              acc[2 + k4] = fmaf(geta[i], x.x, acc[2 + k4]);
              if (k4 + 1 < KSM) acc[3 + k4 < NQ ? 3 + k4 : 0] = fmaf(geta[i], x.y, acc[3 + k4 < NQ ? 3 + k4 : 0]);
              if (k4 + 2 < KSM) acc[4 + k4 < NQ ? 4 + k4 : 0] = fmaf(geta[i], x.z, acc[4 + k4 < NQ ? 4 + k4 : 0]);
              if (k4 + 3 < KSM) acc[5 + k4 < NQ ? 5 + k4 : 0] = fmaf(geta[i], x.w, acc[5 + k4 < NQ ? 5 + k4 : 0]);
            }
          }
        }
      }
    }
#pragma unroll
    for (int i = 1; i <= KB + 1; ++i) g64[(size_t)i * BT] += (double)acc[i];
#pragma unroll
    for (int i = 2 + KB; i < NQ; ++i) g64[(size_t)i * BT] += (double)acc[i] * -kLn2D;  // covariate slots carry -log2(e)
    if (free_ring) {
      __syncwarp();
      if ((tid & 31) == 0) mbar_arrive(&empty[s]);
      continue;
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
    my[0] = logp64;
#pragma unroll
    for (int k = 0; k < KB; ++k)
      if (k <= ks) my[1 + k] = g64[(size_t)(1 + k) * BT];
#pragma unroll
    for (int k = 0; k < KA; ++k) my[2 + ks + k] = g64[(size_t)(1 + KB + k) * BT];
  }
  finish_block<float>(p, c0, ncb, &s_is_last);
}

template <int KS, int KO, int NQD, int NS, int MINB, int BT, bool STRICT = false>
static cudaError_t launch_signed_one(const EvalParams& p, const SignedLayout& S, dim3 grid, size_t smem,
                                     cudaStream_t st, int* occ) {
  auto kern = occu_signed_kernel<KS, KO, NQD, NS, MINB, BT, STRICT>;
  constexpr int NT = BT;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  if (occ) return cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, kern, NT, smem);
  kern<<<grid, NT, smem, st>>>(p, S);
  return cudaGetLastError();
}

// BL_FLAG_STRICT_MATH is supported (the STRICT instantiations); BL_STRICT_ENGINE=1 sends it back to the engine (A/B)
bool occu_signed_supported(int dtype, int ks, int ko, uint32_t flags) {
  if (dtype != BL_F32) return false;
  if (flags & BL_FLAG_STRICT_MATH)
    if (const char* e = getenv("BL_STRICT_ENGINE"))
      if (atoi(e) != 0) return false;
  if (flags & (BL_FLAG_FP_CONSTANT | BL_FLAG_FP_UNOCCUPIED)) return false;
  return ks >= 0 && ks <= kSignedMaxKs && ko >= 1 && ko <= 4;
}

int64_t occu_signed_block_tiles(const Layout& L) {
  const SignedLayout S = make_signed_layout(L);
  return (L.n_tiles + S.G - 1) / S.G;
}

size_t occu_signed_smem(const Layout& L, int nstage, int block_threads) {
  const SignedLayout S = make_signed_layout(L);
  size_t b = 128 + (size_t)nstage * S.R * kWarp * S.G * sizeof(float);
  b = (b + 15) & ~size_t(15);
  return b + (size_t)(3 + kSignedMaxKs + L.ko) * block_threads * sizeof(double);  // fp64 gradient columns
}

// same rule as K1c (occu_chain_block_threads): 256-thread blocks on whole multiples of 256 chains, else 128
int occu_signed_block_threads(int C) {
  if (const char* e = getenv("BL_SIGNED_BT")) return atoi(e) == 128 ? 128 : 256;
  return (C > 0 && C % 256 == 0) ? 256 : 128;
}

// Two sites interleaved per thread (measured, config 2 near the mode, ms: two 7.34, one 7.75; the one-site
// instantiations were a tuning switch only and are no longer built).
template <int KS, int KO, int NQD>
static cudaError_t launch_signed_bt(const EvalParams& p, const SignedLayout& S, dim3 grid, size_t smem,
                                    cudaStream_t st, int* occ) {
  if (p.flags & BL_FLAG_STRICT_MATH) {  // FMA-pipe exp2, libm log2f, IEEE division
    if (p.chain_bt == 128) return launch_signed_one<KS, KO, NQD, 2, 4, 128, true>(p, S, grid, smem, st, occ);
    return launch_signed_one<KS, KO, NQD, 2, 2, 256, true>(p, S, grid, smem, st, occ);
  }
  if (p.chain_bt == 128) return launch_signed_one<KS, KO, NQD, 2, 4, 128>(p, S, grid, smem, st, occ);
  return launch_signed_one<KS, KO, NQD, 2, 2, 256>(p, S, grid, smem, st, occ);
}

cudaError_t launch_occu_signed(const EvalParams& p, dim3 grid, size_t smem, cudaStream_t st, int* occ) {
  const SignedLayout S = make_signed_layout(p.L);
  const int ks = p.L.ks, ko = p.L.ko;
  if (ks == 5 && ko == 3 && S.nq == 2) return launch_signed_bt<5, 3, 2>(p, S, grid, smem, st, occ);
  if (ks == 1 && ko == 1) return launch_signed_bt<1, 1, 0>(p, S, grid, smem, st, occ);
  if (ko == 1) return launch_signed_bt<-1, 1, 0>(p, S, grid, smem, st, occ);
  if (ko == 2) return launch_signed_bt<-1, 2, 0>(p, S, grid, smem, st, occ);
  if (ko == 3) return launch_signed_bt<-1, 3, 0>(p, S, grid, smem, st, occ);
  if (ko == 4) return launch_signed_bt<-1, 4, 0>(p, S, grid, smem, st, occ);
  return cudaErrorNotSupported;
}

}  // namespace bl
