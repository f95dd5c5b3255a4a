// K1s: lane = site occu kernel for SMALL chain batches (fp32, no false-positive extras, C < 32: what `fit` with the
// reference's default num_chains = 5 evaluates, biolith/utils/fit.py:24, and the tail of a NUTS run).
//
// The site-parallel engine (engine.cuh + occu.cu) pays, per (warp-tile, chain), a 16-shuffle transposed butterfly into
// fp64 shared-memory accumulators and 3 MUFU per visit.  This kernel keeps the engine's packed "SoA in tile" dataset
// (128 B / site at config 2 -- a second packing would cost the HBM-bound C = 1 case its bandwidth), staged by TMA
// through one mbarrier ring PER WARP (no block barrier in the loop), but
//   * holds the per-lane sums of a chunk of <= 8 chains in REGISTERS over every tile a thread walks and reduces them
//     across lanes ONCE per block (fp32 over the <= ~20 sites of a lane, fp64 from there on): nothing per (tile, chain);
//   * uses K1d's visit arithmetic (occu_signed.cu; reference: biolith/models/occu.py:221-242 with p_fp = z p):
//       s_j = +1 detection, -1 non-detection, 0 masked;   x'_j = s_j (alpha_0 + W_j . alpha),  e_j = exp(-x'_j)
//       log-lik of visit j = -log(1 + e_j),  d/d alpha = q_j s_j [1, W_j],  q_j = e_j / (1 + e_j)
//     one ex2 per visit (two-float log2 e: no fixed relative error in the argument), ONE lg2 and ONE rcp (+ Newton) per
//     8 visits (product-log, batch inversion through the pair-product tree); masked visits have x' = 0 -> 1 + e = 2
//     exactly and are removed by a count; the sign / mask decode and s_j W_j are done once per site and reused by every
//     chain of the chunk;
//   * numpyro's clamp_probs is kept exactly: a (site, chain) whose pair product reaches 2^23 (some visit at the low
//     clamp) takes the per-visit clamped form (small_slow_site: the engine's softsig<true> arithmetic) instead.
// Chains beyond 8 are cut into chunks on grid.y (the dataset is re-read per chunk, mostly from L2).
// Same outputs and ticketed last-block reduction as every other kernel (finish_block): one launch per evaluation.
#pragma once
#include <cstdlib>

#include "engine.cuh"

namespace bl {

// ONE block per SM, as many warps as the register file carries for NC chains (<= 2: 128 registers x 512 threads,
// <= 5: 168 x 384, <= 8: 255 x 256): with a ring per warp the block size costs nothing in the loop, and the last-block
// reduction sums 148 partial rows instead of 444..592 (measured with 128-thread blocks, 3..4 per SM: C = 1 34.4 us of
// which 21.6 us streaming; the two-level sum of 592 rows was most of the rest).
__host__ __device__ constexpr int small_bt(int nc) { return nc <= 2 ? 512 : (nc <= 5 ? 384 : 256); }
constexpr int kSmallHeader = 512;                  // 16 warps x kMaxStages mbarriers
constexpr int kSmallMaxNC = 8;                     // chains per block (register-resident sums)
constexpr int kSmallMaxKs = 8;
constexpr int kSmallThetaStride = 16;              // staged theta row: [beta (KBM) | alpha_0, alpha_1..KO], padded
constexpr float kSmLog2eLo = 1.925963033500011e-08f;  // log2(e) - (float)log2(e)
constexpr float kSmClampProduct = 8388608.0f;         // 2^23 > (1 - eps) / eps
constexpr float kSmEx2Shift = 7.2134752e-08f;         // 5e-8 / ln 2

// Elementary functions: bounded-error SFU forms, or (STRICT = BL_FLAG_STRICT_MATH) libm exp2f / log2f and IEEE
// division in the same formulation -- as K1d's SMath (occu_signed.cu).
template <bool STRICT> struct SmMath {
  static __device__ __forceinline__ float ex2(float x) {
    if constexpr (STRICT) return exp2f(x); else return sfu::ex2(x);
  }
  static __device__ __forceinline__ float lg2(float x) {
    if constexpr (STRICT) return log2f(x); else return sfu::lg2(x);
  }
  static __device__ __forceinline__ float inv(float x) {
    if constexpr (STRICT) {
      return 1.0f / x;
    } else {
      const float r = sfu::rcp(x);
      return fmaf(r, fmaf(-x, r, 1.0f), r);
    }
  }
};

template <bool STRICT>
__device__ __forceinline__ float sm_exp_neg_abs(float x) {
  const float t = -fabsf(x);
  return SmMath<STRICT>::ex2(fmaf(t, kSmLog2eLo, t * sfu::kLog2e));
}

#ifdef BL_TRACE
// profiling aid (built only with -DBL_TRACE): five %globaltimer stamps per block -- entry, first tile landed (warp 0),
// warp 0 done with its tiles, block partial published, block exit
static __device__ unsigned long long g_small_trace[8 * 2048];  // one copy per translation unit
__device__ __forceinline__ unsigned long long small_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define BL_STAMP(i) do { if (threadIdx.x == 0 && blockIdx.y == 0 && blockIdx.x < 2048) g_small_trace[blockIdx.x * 8 + (i)] = small_now(); } while (0)
#else
#define BL_STAMP(i) do { } while (0)
#endif

template <int KO> struct SmallSlow { float L1; float ga[KO + 1]; };

// exact per-visit form of one site for one chain (rare): numpyro's clamps through sfu::softsig<true>, natural units
template <int KO, bool STRICT>
__device__ __noinline__ SmallSlow<KO> small_slow_site(const float* __restrict__ tile, int lane, int off_w, int off_y,
                                                      int off_m, int J, const float* __restrict__ al) {
  SmallSlow<KO> o;
  o.L1 = 0.f;
#pragma unroll
  for (int k = 0; k <= KO; ++k) o.ga[k] = 0.f;
  uint32_t yw = 0, mw = 0;
  for (int j = 0; j < J; ++j) {
    if ((j & 31) == 0) {
      yw = __float_as_uint(tile[(off_y + (j >> 5)) * kWarp + lane]);
      mw = __float_as_uint(tile[(off_m + (j >> 5)) * kWarp + lane]);
    }
    if (!((mw >> (j & 31)) & 1u)) continue;
    const float yf = ((yw >> (j & 31)) & 1u) ? 1.f : 0.f;
    float w[KO];
    float nu = al[0];
#pragma unroll
    for (int k = 0; k < KO; ++k) {
      w[k] = tile[(off_w + j * KO + k) * kWarp + lane];
      nu = fmaf(w[k], al[1 + k], nu);
    }
    float g;
    if constexpr (STRICT) {  // the engine's libm form (common.cuh log_sigmoid_pair)
      const LogSig<float> ls = log_sigmoid_pair<float>(nu);
      o.L1 += yf > 0.f ? ls.lp : ls.l1mp;
      g = ls.inr ? (yf > 0.f ? ls.q : -ls.p) : 0.f;
    } else {
      const sfu::SoftSig ss = sfu::softsig<true>(nu);
      o.L1 += fmaf(yf, ss.xc, -ss.s);
      g = ss.inr ? (yf - ss.p) : 0.f;
    }
    o.ga[0] += g;
#pragma unroll
    for (int k = 0; k < KO; ++k) o.ga[1 + k] = fmaf(g, w[k], o.ga[1 + k]);
  }
  return o;
}

// NV = 4 or 8 visits of one site for one chain.  s[j] = sign, sw[j][k] = s_j W_jk (exact), al = [alpha_0, alpha_1..KO].
template <int KO, int NV, bool STRICT>
__device__ __forceinline__ void small_visits(const float (&s)[NV], const float (&sw)[NV][KO], const float (&al)[KO + 1],
                                             float& lgsum, float& mx, float (&ga)[KO + 1]) {
  static_assert(NV == 4 || NV == 8, "visits are processed in quads or octets");
  using M = SmMath<STRICT>;
  float e[NV], u[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    float xp = s[j] * al[0];
#pragma unroll
    for (int k = 0; k < KO; ++k) xp = fmaf(sw[j][k], al[1 + k], xp);
    // exp(-x'); the SFU form shifts the argument by +5e-8 / ln 2: MUFU.EX2's mean error at negative arguments,
    // compensated where it matters (K1d's kEx2Shift, occu_signed.cu; scripts/ex2_bias_emulation.py)
    e[j] = M::ex2(fmaf(xp, -kSmLog2eLo, STRICT ? xp * -sfu::kLog2e : fmaf(xp, -sfu::kLog2e, kSmEx2Shift)));
    u[j] = 1.0f + e[j];
  }
  float pr[NV / 2], rp[NV / 2];
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
    const float ra = rinv * pb, rb = rinv * pa;
    rp[0] = ra * pr[1]; rp[1] = ra * pr[0];
    rp[2] = rb * pr[3]; rp[3] = rb * pr[2];
  }
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const float qj = e[j] * (rp[j / 2] * u[j ^ 1]);  // e_j / u_j
    ga[0] = fmaf(qj, s[j], ga[0]);
#pragma unroll
    for (int k = 0; k < KO; ++k) ga[1 + k] = fmaf(qj, sw[j][k], ga[1 + k]);
  }
}

// sign and s W of visit j of this lane's site, read from the staged tile (j >= J: a masked visit)
template <int KO>
__device__ __forceinline__ void small_decode(const float* __restrict__ tile, int lane, int off_w, uint32_t yw,
                                             uint32_t mw, int j, int J, float& s, float (&sw)[KO]) {
  const bool in = j < J;
  const bool m = in && ((mw >> (j & 31)) & 1u);
  const bool y = (yw >> (j & 31)) & 1u;
  s = m ? (y ? 1.f : -1.f) : 0.f;
#pragma unroll
  for (int k = 0; k < KO; ++k) {
    const float w = in ? tile[(off_w + j * KO + k) * kWarp + lane] : 0.f;
    sw[k] = s * w;
  }
}

// KS < 0: runtime Ks (<= kSmallMaxKs).  J8: the site's 8 visits are decoded once into registers (J == 8); otherwise
// visits are decoded per chain from shared memory in quads (any J).
template <int KS, int KO, bool J8, int NC, bool STRICT>
__global__ void __launch_bounds__(small_bt(NC), 1) occu_small_kernel(const __grid_constant__ EvalParams p) {
  using M = SmMath<STRICT>;
  constexpr int kSmallBT = small_bt(NC), kSmallWarpsMax = kSmallBT / kWarp;
  const int kSmallWarps = (int)blockDim.x / kWarp;  // fewer than the maximum when a warp-tile is wide (ring budget)
  constexpr int KSM = KS < 0 ? kSmallMaxKs : KS;
  constexpr int KBM = KSM + 1, KA = KO + 1, NG = KBM + KA;  // gradient slots in this kernel's (padded) order
  static_assert(NG <= 16 && KBM + KA <= kSmallThetaStride, "one 16-wide butterfly / theta row");
  const int ks = KS < 0 ? p.L.ks : KS;
  const int J = J8 ? 8 : p.L.J;
  const int F = p.L.F, off_w = p.L.off_w, off_y = p.L.off_y, off_m = p.L.off_m;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
  float* stage0 = reinterpret_cast<float*>(smem_raw + kSmallHeader);
  __shared__ int s_is_last;
  __shared__ __align__(16) float s_th[NC * kSmallThetaStride];
  __shared__ double s_red[kSmallWarpsMax][NC][1 + 16];
  __shared__ double s_scr[kSmallBT];  // scratch of the cooperative last-block reduction
  // one TMA ring PER WARP (warp-tile = 32 sites = F x 128 B, one cp.async.bulk each): no block barrier in the loop,
  // the warps drift freely (measured with a block-wide ring and one barrier per 4 warp-tiles: barrier stalls 0.65
  // per issue at C = 5)
  const uint32_t tile_elems = (uint32_t)F * kWarp;
  const uint32_t tile_bytes = tile_elems * sizeof(float);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int c0 = blockIdx.y * p.CB;
  const int ncb = min(p.CB, p.C - c0);
  const int64_t nwt = p.L.n_tiles;
  const int64_t wt_begin = nwt * blockIdx.x / gridDim.x;
  const int64_t wt_end = nwt * (blockIdx.x + 1) / gridDim.x;
  const int64_t wt_mine = wt_end - wt_begin - warp;  // this warp takes wt_begin + warp, + kSmallWarps, ...
  const int n_it = wt_mine > 0 ? (int)((wt_mine + kSmallWarps - 1) / kSmallWarps) : 0;
  const float* packed = reinterpret_cast<const float*>(p.packed) + (size_t)(wt_begin + warp) * tile_elems;
  const size_t kStep = (size_t)kSmallWarps;  // warp-tiles between two iterations of a warp
  uint64_t* wbars = bars + warp * kMaxStages;
  float* wstage0 = stage0 + (size_t)warp * p.nstage * tile_elems;

  BL_STAMP(0);
  if (lane == 0) {  // the data does not depend on theta: start the copies before anything else
    for (int s = 0; s < p.nstage; ++s) mbar_init(&wbars[s], 1);
    fence_mbar_init();
    if (n_it > 0) {  // stage 0 of every warp first: the first tiles land ~2 us earlier than behind a full ring
      mbar_expect_tx(&wbars[0], tile_bytes);
      tma_load_bulk(wstage0, packed, tile_bytes, &wbars[0]);
    }
  }
  for (int i = tid; i < NC * kSmallThetaStride; i += (int)blockDim.x) {
    const int c = i / kSmallThetaStride, e = i % kSmallThetaStride;
    float v = 0.f;
    if (c < ncb) {
      const float* th = reinterpret_cast<const float*>(p.theta) + (size_t)(c0 + c) * p.D;
      if (e < KBM) v = (e <= ks) ? th[e] : 0.f;
      else if (e - KBM < KA) v = th[ks + 1 + (e - KBM)];
    }
    s_th[i] = v;
  }
  __syncthreads();  // theta staged, every warp's barriers initialised
  if (lane == 0) {
    const int pre = min(p.nstage, n_it);
    for (int s = 1; s < pre; ++s) {
      mbar_expect_tx(&wbars[s], tile_bytes);
      tma_load_bulk(wstage0 + (size_t)s * tile_elems, packed + (size_t)s * kStep * tile_elems, tile_bytes, &wbars[s]);
    }
  }
  const float log_tiny = Num<float>::log_tiny();
  const int nq = (J + 3) / 4;

  double lp64[NC];
  float acc[NC][NG];
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    lp64[c] = 0.0;
#pragma unroll
    for (int i = 0; i < NG; ++i) acc[c][i] = 0.f;
  }

  for (int it = 0; it < n_it; ++it) {
    const int st = it % p.nstage;
    mbar_wait(&wbars[st], (uint32_t)((it / p.nstage) & 1));
    if (it == 0) BL_STAMP(1);
    const float* tile = wstage0 + (size_t)st * tile_elems;
    const int64_t unit = (wt_begin + warp + (int64_t)it * kSmallWarps) * kWarp + lane;
    const float vf = unit < p.L.n_units ? 1.f : 0.f;

    float x[KSM];
#pragma unroll
    for (int k = 0; k < KSM; ++k) x[k] = (k < ks) ? tile[k * kWarp + lane] : 0.f;
    const uint32_t yw0 = __float_as_uint(tile[off_y * kWarp + lane]);
    const uint32_t mw0 = __float_as_uint(tile[off_m * kWarp + lane]);
    int n1i = 0, nmask = 4 * nq;
    if constexpr (J8) {
      n1i = __popc(yw0 & mw0);
      nmask -= __popc(mw0);
    } else {
      for (int w = 0; w < p.L.nw; ++w) {
        const uint32_t yw = __float_as_uint(tile[(off_y + w) * kWarp + lane]);
        const uint32_t mw = __float_as_uint(tile[(off_m + w) * kWarp + lane]);
        n1i += __popc(yw & mw);
        nmask -= __popc(mw);
      }
    }
    const float bl0 = (float)n1i * log_tiny;  // z = 0 branch: n1 log(tiny) (data only)
    const float cnt = (float)nmask;          // masked + padded visits: each adds exactly 1 to the lg2 sum

    float s8[J8 ? 8 : 1], sw8[J8 ? 8 : 1][KO];
    if constexpr (J8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) small_decode<KO>(tile, lane, off_w, yw0, mw0, j, 8, s8[j], sw8[j]);
    }

#pragma unroll
    for (int c = 0; c < NC; ++c) {
      // rows c >= ncb of a ragged last chunk hold theta = 0: evaluated like any chain (finite), never published
      float tt[kSmallThetaStride];
#pragma unroll
      for (int i = 0; i < kSmallThetaStride / 4; ++i) {
        const float4 t = *reinterpret_cast<const float4*>(s_th + c * kSmallThetaStride + 4 * i);
        tt[4 * i] = t.x; tt[4 * i + 1] = t.y; tt[4 * i + 2] = t.z; tt[4 * i + 3] = t.w;
      }
      float al[KA];
#pragma unroll
      for (int k = 0; k < KA; ++k) al[k] = tt[KBM + k];
      float lgsum = 0.f, mx = 0.f, ga[KA];
#pragma unroll
      for (int k = 0; k < KA; ++k) ga[k] = 0.f;
      if constexpr (J8) {
        small_visits<KO, 8, STRICT>(s8, sw8, al, lgsum, mx, ga);
      } else {
        uint32_t yw = yw0, mw = mw0;
        for (int q = 0; q < nq; ++q) {
          if (q > 0 && (q & 7) == 0) {
            yw = __float_as_uint(tile[(off_y + (q >> 3)) * kWarp + lane]);
            mw = __float_as_uint(tile[(off_m + (q >> 3)) * kWarp + lane]);
          }
          float s4[4], sw4[4][KO];
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) small_decode<KO>(tile, lane, off_w, yw, mw, q * 4 + jj, J, s4[jj], sw4[jj]);
          small_visits<KO, 4, STRICT>(s4, sw4, al, lgsum, mx, ga);
        }
      }
      float L1 = -sfu::kLn2 * (lgsum - cnt);
      const bool slow = mx >= kSmClampProduct;
      if (__any_sync(0xffffffffu, slow)) {
        const SmallSlow<KO> so = small_slow_site<KO, STRICT>(tile, lane, off_w, off_y, off_m, J, s_th + c * kSmallThetaStride + KBM);
        if (slow) {
          L1 = so.L1;
#pragma unroll
          for (int k = 0; k < KA; ++k) ga[k] = so.ga[k];
        }
      }
      float eta = tt[0];
#pragma unroll
      for (int k = 0; k < KSM; ++k) eta = fmaf(x[k], tt[1 + k], eta);
      // site level as in K1d: psi~ = sigmoid(xc), a = log psi~ + L1, b = log(1 - psi~) + n1 log tiny,
      //   d = a - b = xc + L1 - n1 log tiny,  logaddexp(a, b) = max(al, bl) - max(xc, 0) + log(u_d / u_e)
      const float xc = fminf(fmaxf(eta, sfu::kXLo), sfu::kXHi);
      const bool inr = xc == eta;
      const float te = sm_exp_neg_abs<STRICT>(xc);
      const float ue = 1.0f + te;
      const float inve = M::inv(ue);
      const float psi = (xc >= 0.f) ? inve : te * inve;
      const float av = xc + L1;
      const float d = av - bl0;
      const float td = sm_exp_neg_abs<STRICT>(d);
      const float ud = 1.0f + td;
      const float invd = M::inv(ud);
      const float rr = (d >= 0.f) ? invd : td * invd;  // P(z = 1 | y)
      const float r = rr * vf;
      const float ell = fmaf(M::lg2(ud * inve), sfu::kLn2, (d >= 0.f ? av : bl0) - fmaxf(xc, 0.f)) * vf;
      const float geta = inr ? (rr - psi) * vf : 0.f;
      lp64[c] += (double)ell;
      acc[c][0] += geta;
#pragma unroll
      for (int k = 0; k < KSM; ++k) acc[c][1 + k] = fmaf(geta, x[k], acc[c][1 + k]);
#pragma unroll
      for (int k = 0; k < KA; ++k) acc[c][KBM + k] = fmaf(r, ga[k], acc[c][KBM + k]);
    }
    __syncwarp();  // every lane is done reading stage st
    if (lane == 0 && it + p.nstage < n_it) {
      mbar_expect_tx(&wbars[st], tile_bytes);
      tma_load_bulk(wstage0 + (size_t)st * tile_elems, packed + (size_t)(it + p.nstage) * kStep * tile_elems, tile_bytes,
                    &wbars[st]);
    }
  }

  BL_STAMP(2);
  // once per block: lanes -> warp (fixed butterfly), warps -> block (fixed order), publish [bx][c][q]
  const int NQ = p.NQ;
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    if (c < ncb) {
      const double lp = warp_sum<double>(lp64[c]);
      float v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = i < NG ? acc[c][i < NG ? i : 0] : 0.f;
      const float r = transpose_reduce<float, 16>(v, lane);  // lane l: sum of slot l >> 1
      if (lane == 0) s_red[warp][c][0] = lp;
      if ((lane & 1) == 0) s_red[warp][c][1 + (lane >> 1)] = (double)r;
    }
  }
  __syncthreads();
  double* my_partial = p.partial + ((size_t)blockIdx.x * p.C + c0) * NQ;
  for (int i = tid; i < ncb * NQ; i += (int)blockDim.x) {
    const int c = i / NQ, q = i % NQ;
    // output order [logp | beta_0..ks | alpha_0..ko] from this kernel's padded slots [beta (KBM) | alpha (KA)]
    const int slot = q == 0 ? 0 : (q - 1 <= ks ? q : 1 + KBM + (q - 1 - (ks + 1)));
    double v = 0.0;
    for (int w = 0; w < kSmallWarps; ++w) v += s_red[w][c][slot];
    my_partial[i] = v;
  }
  BL_STAMP(3);
  finish_block<float>(p, c0, ncb, &s_is_last, s_scr);
  BL_STAMP(4);
}

// ---- launch ----------------------------------------------------------------------------------------------------
template <int KS, int KO, bool J8, int NC, bool STRICT>
static cudaError_t launch_small_strict(const EvalParams& p, dim3 grid, size_t smem, cudaStream_t st, int* occ) {
  auto kern = occu_small_kernel<KS, KO, J8, NC, STRICT>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  if (occ) return cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, kern, p.chain_bt, smem);
  kern<<<grid, p.chain_bt, smem, st>>>(p);
  return cudaGetLastError();
}

template <int KS, int KO, bool J8, bool STRICT>
static cudaError_t launch_small_nc(const EvalParams& p, dim3 grid, size_t smem, cudaStream_t st, int* occ) {
  switch (p.CB) {
    case 1: return launch_small_strict<KS, KO, J8, 1, STRICT>(p, grid, smem, st, occ);
    case 2: return launch_small_strict<KS, KO, J8, 2, STRICT>(p, grid, smem, st, occ);
    case 3: return launch_small_strict<KS, KO, J8, 3, STRICT>(p, grid, smem, st, occ);
    case 4: return launch_small_strict<KS, KO, J8, 4, STRICT>(p, grid, smem, st, occ);
    case 5: return launch_small_strict<KS, KO, J8, 5, STRICT>(p, grid, smem, st, occ);
    case 6: return launch_small_strict<KS, KO, J8, 6, STRICT>(p, grid, smem, st, occ);
    case 7: return launch_small_strict<KS, KO, J8, 7, STRICT>(p, grid, smem, st, occ);
    case 8: return launch_small_strict<KS, KO, J8, 8, STRICT>(p, grid, smem, st, occ);
    default: return cudaErrorNotSupported;
  }
}

}  // namespace bl
