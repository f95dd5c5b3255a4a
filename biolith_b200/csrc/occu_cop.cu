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
#include <cstdlib>
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
  static constexpr int kMultiChain = 1;  // chains per pass over a warp-tile (engine.cuh)

  struct Site {
    T x[KSM];
    T sy, st;
  };
  static __device__ __forceinline__ T unit_const(const EvalParams& p, const T* tile, int lane) {
    return tile[(p.L.off_sy + 2) * kWarp + lane];
  }

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
                                                    const Site& s, const T* __restrict__ th, T* __restrict__ q,
                                                    T* __restrict__ extra = nullptr) {
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
    int i = 3 + ks + ko;
    const T w0 = (r < T(1)) ? (T(1) - r) * d0 : T(0);
    if (p.flags & BL_FLAG_FP_CONSTANT) q[i++] = (r * s1 + w0) * c;  // x = log c: dc/dx = c
    if (p.flags & BL_FLAG_FP_UNOCCUPIED) q[i++] = w0 * u;
  }
};

// ------------------------------------------------------------------------------------------------
// K3c: chain-parallel count-detection kernel (fp32, SFU math, C >= 64).  Same mapping as K1c
// (occu_chain.cu): lane = chain with theta and the running sums in registers, sites warp-broadcast, 4
// consecutive sites per LDS.128.  Masked visits are packed as (y, T) = (0, 0) and vanish on their own,
// so the loop has no mask handling at all; per visit: exp (ex2), log (lg2) and reciprocal (rcp) of the
// rate -> 3 MUFU like the Bernoulli kernel.
// ------------------------------------------------------------------------------------------------
template <int KS, int KO, int BT, int JT>
__global__ void __launch_bounds__(BT, BT == 128 ? 4 : 2) occu_cop_chain_kernel(const __grid_constant__ EvalParams p) {
  // KS < 0: runtime number of site covariates (<= 8); accumulator slots are laid out for the capacity
  constexpr int KSM = KS < 0 ? 8 : KS;
  constexpr int KB = KSM + 1, KA = KO + 1, NS = 4, NQM = 1 + KB + KA + 2;
  const int ks = KS < 0 ? p.L.ks : KS;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
  float* stage0 = reinterpret_cast<float*>(smem_raw + 128);
  __shared__ int s_is_last;
  const int F = p.L.F, J = JT > 0 ? JT : p.L.J, NQ = p.NQ, D = p.D;
  const uint32_t tile_elems = (uint32_t)F * kWarp;
  const uint32_t tile_bytes = tile_elems * sizeof(float);
  const int tid = threadIdx.x;
  double* g64 = reinterpret_cast<double*>(stage0 + (size_t)p.nstage * tile_elems) + tid;  // [NQM][BT]
  const int c0 = blockIdx.y * p.CB;
  const int ncb = min(p.CB, p.C - c0);
  const bool chain_ok = tid < ncb;
  const bool warp_on = (tid & ~31) < ncb;  // warps past the end of the batch only help stage the tiles
  const int64_t nbt = p.n_block_tiles;
  const int64_t bt_begin = nbt * blockIdx.x / gridDim.x;
  const int64_t bt_end = nbt * (blockIdx.x + 1) / gridDim.x;
  const int n_it = (int)(bt_end - bt_begin);
  const float* packed = reinterpret_cast<const float*>(p.packed);
  const bool fpc = (p.flags & BL_FLAG_FP_CONSTANT) != 0, fpu = (p.flags & BL_FLAG_FP_UNOCCUPIED) != 0;

  if (tid == 0) {
    for (int s = 0; s < p.nstage; ++s) mbar_init(&bars[s], 1);
    fence_mbar_init();
  }
  float b[KB], a[KA];
  float c = 0.f, u = 0.f;
  {
    const float* th = reinterpret_cast<const float*>(p.theta) + (size_t)(c0 + (chain_ok ? tid : 0)) * D;
#pragma unroll
    for (int k = 0; k < KB; ++k) b[k] = (k <= ks) ? th[k] : 0.f;
#pragma unroll
    for (int k = 0; k < KA; ++k) a[k] = th[ks + 1 + k];
    int i = ks + 1 + KA;
    if (fpc) c = expf(th[i++]);
    if (fpu) u = expf(th[i++]);
  }
  const float rho0 = u + c;
  const float lrho0 = rho0 > 0.f ? logf(rho0) : -Num<float>::inf();
  const float irho0 = rho0 > 0.f ? 1.f / rho0 : 0.f;
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
    float acc[NQM];
#pragma unroll
    for (int i = 0; i < NQM; ++i) acc[i] = 0.f;
    const int n_mine = warp_on ? n_valid : 0;
    for (int g0 = 0; g0 < n_mine; g0 += NS) {
      float eta[NS], geta[NS];
#pragma unroll
      for (int i = 0; i < NS; ++i) eta[i] = b[0];
#pragma unroll
      for (int k = 0; k < KSM; ++k) {
        if (k < ks) {  // re-read below for the gradient instead of held in registers
          const float4 v = *reinterpret_cast<const float4*>(tile + k * kWarp + g0);
          const float xk[NS] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int i = 0; i < NS; ++i) eta[i] = fmaf(xk[i], b[1 + k], eta[i]);
        }
      }
      float L1[NS], s1[NS], ga0[NS], ga[KO > 0 ? KO : 1][NS];
#pragma unroll
      for (int i = 0; i < NS; ++i) {
        L1[i] = 0.f; s1[i] = 0.f; ga0[i] = 0.f;
#pragma unroll
        for (int k = 0; k < KO; ++k) ga[k][i] = 0.f;
      }
#pragma unroll(JT > 0 ? JT : 2)
      for (int j = 0; j < J; ++j) {
        float w[KO > 0 ? KO : 1][NS], nu[NS];
#pragma unroll
        for (int i = 0; i < NS; ++i) nu[i] = a[0];
#pragma unroll
        for (int k = 0; k < KO; ++k) {
          const float4 v = *reinterpret_cast<const float4*>(tile + (p.L.off_w + j * KO + k) * kWarp + g0);
          w[k][0] = v.x; w[k][1] = v.y; w[k][2] = v.z; w[k][3] = v.w;
#pragma unroll
          for (int i = 0; i < NS; ++i) nu[i] = fmaf(w[k][i], a[1 + k], nu[i]);
        }
        const float4 yv = *reinterpret_cast<const float4*>(tile + (p.L.off_y + j) * kWarp + g0);
        const float4 tv = *reinterpret_cast<const float4*>(tile + (p.L.off_t + j) * kWarp + g0);
        const float y[NS] = {yv.x, yv.y, yv.z, yv.w}, T[NS] = {tv.x, tv.y, tv.z, tv.w};
#pragma unroll
        for (int i = 0; i < NS; ++i) {
          const float mu = sfu::ex2(nu[i] * sfu::kLog2e);
          const float rho1 = mu + c;
          const bool ypos = y[i] > 0.f;
          const float ylog = ypos ? (y[i] * sfu::kLn2) * sfu::lg2(rho1) : 0.f;  // xlogy(y, rho1)
          const float yinv = ypos ? y[i] * sfu::rcp(rho1) : 0.f;
          L1[i] += fmaf(-T[i], rho1, ylog);
          const float d1 = yinv - T[i];  // dt1/drho1
          s1[i] += d1;
          const float g = d1 * mu;       // drho1/dnu = mu
          ga0[i] += g;
#pragma unroll
          for (int k = 0; k < KO; ++k) ga[k][i] = fmaf(g, w[k][i], ga[k][i]);
        }
      }
      const float4 syv = *reinterpret_cast<const float4*>(tile + p.L.off_sy * kWarp + g0);
      const float4 stv = *reinterpret_cast<const float4*>(tile + (p.L.off_sy + 1) * kWarp + g0);
      const float sy[NS] = {syv.x, syv.y, syv.z, syv.w}, st[NS] = {stv.x, stv.y, stv.z, stv.w};
#pragma unroll
      for (int i = 0; i < NS; ++i) {
        const float vf = (g0 + i < n_valid) ? 1.f : 0.f;
        const float L0 = (sy[i] > 0.f ? sy[i] * lrho0 : 0.f) - st[i] * rho0;
        const float d0 = (sy[i] > 0.f ? sy[i] * irho0 : 0.f) - st[i];
        const sfu::SoftSig se = sfu::softsig<true>(eta[i]);
        const float av = (se.xc - se.s) + L1[i];
        const float bv = L0 - se.s;
        float rr, ell;
        if (bv == -Num<float>::inf()) {  // Poisson(0) saw a count: occupied with certainty
          rr = 1.f;
          ell = av;
        } else {
          const float d = av - bv;
          const float td = sfu::ex2(-fabsf(d) * sfu::kLog2e);
          const float ud = 1.0f + td;
          const float invd = sfu::rcp(ud);
          rr = (d >= 0.f) ? invd : td * invd;
          ell = fmaf(sfu::lg2(ud), sfu::kLn2, fmaxf(av, bv));
        }
        const float r = rr * vf;
        geta[i] = se.inr ? (rr - se.p) * vf : 0.f;
        const float w0 = (rr < 1.f) ? (1.f - rr) * d0 * vf : 0.f;
        logp64 += (double)(ell * vf);
        acc[1] += geta[i];
        acc[1 + KB] = fmaf(r, ga0[i], acc[1 + KB]);
#pragma unroll
        for (int k = 0; k < KO; ++k) acc[2 + KB + k] = fmaf(r, ga[k][i], acc[2 + KB + k]);
        acc[1 + KB + KA] += fmaf(r, s1[i], w0);  // dl/dc (constant fp): both branches
        acc[2 + KB + KA] += w0;                  // dl/du (unoccupied fp): z = 0 branch only
      }
#pragma unroll
      for (int k = 0; k < KSM; ++k) {
        if (k < ks) {
          const float4 v = *reinterpret_cast<const float4*>(tile + k * kWarp + g0);
          const float xk[NS] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int i = 0; i < NS; ++i) acc[2 + k] = fmaf(geta[i], xk[i], acc[2 + k]);
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
    int i = 2 + ks + KA;
    const double gc = g64[(size_t)(1 + KB + KA) * BT], gu = g64[(size_t)(2 + KB + KA) * BT];
    if (fpc) my[i++] = gc * (double)c;  // x = log c: dc/dx = c
    if (fpu) my[i++] = gu * (double)u;
  }
  finish_block<float>(p, c0, ncb, &s_is_last);
}

static int cop_chain_variant() {
  const char* e = getenv("BL_CHAIN_VARIANT");  // tuning switch: 2 = 128-thread blocks, 3 = 256-thread blocks
  return e ? atoi(e) : 0;
}

bool occu_cop_chain_supported(int dtype, int ks, int ko, uint32_t flags) {
  if (dtype != BL_F32 || (flags & BL_FLAG_STRICT_MATH)) return false;
  return (ks == 5 && ko == 3) || (ks >= 0 && ks <= 8 && ko >= 1 && ko <= 4);
}

// threads (= chains) per block for a batch of C chains; see occu_chain.cu:occu_chain_block_threads.
// Measured on B200 (config 4, ms per evaluation, 128 / 256 threads): C=32 0.36 / 0.59, 64 0.58 / 0.91,
// 96 0.94 / 0.94, 128 1.11 / 0.96, 192 1.59 / 1.77, 256 1.84 / 1.82, 384 2.64 / 3.38, 512 3.46 / 3.41,
// 1024 6.77 / 6.70 (site-parallel engine: 32 0.45, 64 0.87, 128 1.70).
int occu_cop_chain_block_threads(int C) {
  if (cop_chain_variant() == 2) return 128;
  if (cop_chain_variant() == 3) return 256;
  if (C > 64 && C <= 128) return 256;
  return (C > 0 && C % 256 == 0) ? 256 : 128;
}

size_t occu_cop_chain_smem(const Layout& L, int nstage, int bt) {
  size_t bts = 128 + (size_t)nstage * L.F * kWarp * sizeof(float);
  bts = (bts + 15) & ~size_t(15);
  return bts + (size_t)(5 + 8 + L.ko) * bt * sizeof(double);
}

template <int KS, int KO, int JT, int BT>
static cudaError_t launch_cop_chain_bt(const EvalParams& p, dim3 grid, size_t smem, cudaStream_t st, int* occ) {
  auto kern = occu_cop_chain_kernel<KS, KO, BT, JT>;
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

template <int KS, int KO, int JT>
static cudaError_t launch_cop_chain_one(const EvalParams& p, dim3 grid, size_t smem, cudaStream_t st, int* occ) {
  if (p.chain_bt == 128) return launch_cop_chain_bt<KS, KO, JT, 128>(p, grid, smem, st, occ);
  return launch_cop_chain_bt<KS, KO, JT, 256>(p, grid, smem, st, occ);
}

cudaError_t launch_occu_cop_chain(const EvalParams& p, dim3 grid, size_t smem, cudaStream_t st, int* occ) {
  if (p.L.ks == 5 && p.L.ko == 3) {
    if (p.L.J == 12) return launch_cop_chain_one<5, 3, 12>(p, grid, smem, st, occ);
    return launch_cop_chain_one<5, 3, 0>(p, grid, smem, st, occ);
  }
  if (p.L.ks >= 0 && p.L.ks <= 8) {  // runtime Ks
    if (p.L.ko == 1) return launch_cop_chain_one<-1, 1, 0>(p, grid, smem, st, occ);
    if (p.L.ko == 2) return launch_cop_chain_one<-1, 2, 0>(p, grid, smem, st, occ);
    if (p.L.ko == 3) return launch_cop_chain_one<-1, 3, 0>(p, grid, smem, st, occ);
    if (p.L.ko == 4) return launch_cop_chain_one<-1, 4, 0>(p, grid, smem, st, occ);
  }
  return cudaErrorNotSupported;
}

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

cudaError_t launch_occu_cop_summary(const EvalParams& p, int dtype, float* out, cudaStream_t st) {
  if (dtype == BL_F64) return launch_summary<double, OccuCopModel<double, -1, -1, true>>(p, out, st);
  return launch_summary<float, OccuCopModel<float, -1, -1, true>>(p, out, st);
}

int occu_cop_derived_slots(uint32_t) { return 6; }

}  // namespace bl
