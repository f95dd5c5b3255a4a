// K1c: chain-parallel variant of the occu kernel for large chain batches (C >= 64).
//
//   lane  = chain: every thread keeps ONE chain's beta/alpha in registers and its 1+D running sums
//           (fp32 inside a tile, fp64 across tiles) -> no cross-lane reduction anywhere in the loop;
//   site  = warp-broadcast: all lanes read the same packed fields (conflict-free broadcast LDS.128
//           covering NS = 4 consecutive sites of the "SoA in tile" layout at once -> 4-way ILP);
//   block = 256 chains x a contiguous range of site tiles, staged by TMA (cp.async.bulk) through the
//           same mbarrier ring as the site-parallel engine; grid = (site splits, chain chunks).
// The arithmetic is the closed form of oracle/occupancy.py:occu_logp_grad (reference:
// biolith/models/occu.py:182-242, regression/linear.py:59-66) with bounded-error SFU math
// (common.cuh sfu::softsig).  fp32, no false-positive extras; everything else uses the engine path.
#include "engine.cuh"

namespace bl {

constexpr int kNS = 4;  // sites processed together by one thread (one LDS.128 per field)

template <int KS, int KO, int MINB>
__global__ void __launch_bounds__(kBlockThreads, MINB) occu_chain_kernel(const EvalParams p) {
  constexpr int KB = KS + 1, KA = KO + 1, NQ = 1 + KB + KA;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
  float* stage0 = reinterpret_cast<float*>(smem_raw + 128);
  __shared__ int s_is_last;
  const int F = p.L.F, J = p.L.J;
  const int WS = p.WS;  // warp-tiles per stage
  const uint32_t tile_elems = (uint32_t)WS * F * kWarp;
  const uint32_t tile_bytes = tile_elems * sizeof(float);
  const int tid = threadIdx.x;
  const int c0 = blockIdx.y * p.CB;
  const int ncb = min(p.CB, p.C - c0);
  const bool chain_ok = tid < ncb;
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
  {
    const float* th = reinterpret_cast<const float*>(p.theta) + (size_t)(c0 + (chain_ok ? tid : 0)) * p.D;
#pragma unroll
    for (int k = 0; k < KB; ++k) b[k] = th[k];
#pragma unroll
    for (int k = 0; k < KA; ++k) a[k] = th[KB + k];
  }
  double acc64[NQ];
#pragma unroll
  for (int i = 0; i < NQ; ++i) acc64[i] = 0.0;
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
    mbar_wait(&bars[s], (uint32_t)((it / p.nstage) & 1));
    for (int wt = 0; wt < WS; ++wt) {
      const float* tile = stage0 + (size_t)s * tile_elems + (size_t)wt * F * kWarp;
      const int64_t unit0 = ((bt_begin + it) * WS + wt) * kWarp;
      const int n_valid = (int)max((int64_t)0, min((int64_t)kWarp, p.L.n_units - unit0));
      float acc[NQ];
#pragma unroll
      for (int i = 0; i < NQ; ++i) acc[i] = 0.f;
      for (int g0 = 0; g0 < n_valid; g0 += kNS) {
        // ---- site-level linear predictor for NS sites
        float x[KS > 0 ? KS : 1][kNS], eta[kNS];
#pragma unroll
        for (int i = 0; i < kNS; ++i) eta[i] = b[0];
#pragma unroll
        for (int k = 0; k < KS; ++k) {
          const float4 v = *reinterpret_cast<const float4*>(tile + k * kWarp + g0);
          x[k][0] = v.x; x[k][1] = v.y; x[k][2] = v.z; x[k][3] = v.w;
#pragma unroll
          for (int i = 0; i < kNS; ++i) eta[i] = fmaf(x[k][i], b[1 + k], eta[i]);
        }
        float L1[kNS], ga0[kNS], ga[KO > 0 ? KO : 1][kNS];
#pragma unroll
        for (int i = 0; i < kNS; ++i) {
          L1[i] = 0.f; ga0[i] = 0.f;
#pragma unroll
          for (int k = 0; k < KO; ++k) ga[k][i] = 0.f;
        }
        uint32_t yw[kNS], mw[kNS];
        // ---- visits
#pragma unroll 2
        for (int j = 0; j < J; ++j) {
          if ((j & 31) == 0) {
            const uint4 yv = *reinterpret_cast<const uint4*>(tile + (p.L.off_y + (j >> 5)) * kWarp + g0);
            const uint4 mv = *reinterpret_cast<const uint4*>(tile + (p.L.off_m + (j >> 5)) * kWarp + g0);
            yw[0] = yv.x; yw[1] = yv.y; yw[2] = yv.z; yw[3] = yv.w;
            mw[0] = mv.x; mw[1] = mv.y; mw[2] = mv.z; mw[3] = mv.w;
          }
          float w[KO > 0 ? KO : 1][kNS], nu[kNS];
#pragma unroll
          for (int i = 0; i < kNS; ++i) nu[i] = a[0];
#pragma unroll
          for (int k = 0; k < KO; ++k) {
            const float4 v = *reinterpret_cast<const float4*>(tile + (p.L.off_w + j * KO + k) * kWarp + g0);
            w[k][0] = v.x; w[k][1] = v.y; w[k][2] = v.z; w[k][3] = v.w;
#pragma unroll
            for (int i = 0; i < kNS; ++i) nu[i] = fmaf(w[k][i], a[1 + k], nu[i]);
          }
          const uint32_t bit = 1u << (j & 31);
#pragma unroll
          for (int i = 0; i < kNS; ++i) {
            const sfu::SoftSig ss = sfu::softsig<true>(nu[i]);
            const bool y = (yw[i] & bit) != 0, m = (mw[i] & bit) != 0;
            // log-lik term  y*xc - softplus(xc);  d/dnu = y - p  (zero outside the clamp range)
            float term = (y ? ss.xc : 0.f) - ss.s;
            float g = (y ? 1.f : 0.f) - ss.p;
            g = (ss.inr && m) ? g : 0.f;
            term = m ? term : 0.f;
            L1[i] += term;
            ga0[i] += g;
#pragma unroll
            for (int k = 0; k < KO; ++k) ga[k][i] = fmaf(g, w[k][i], ga[k][i]);
          }
        }
        // ---- marginalise z, accumulate this chain's sums
        const float4 n1v = *reinterpret_cast<const float4*>(tile + p.L.off_n1 * kWarp + g0);
        const float n1[kNS] = {n1v.x, n1v.y, n1v.z, n1v.w};
#pragma unroll
        for (int i = 0; i < kNS; ++i) {
          const float vf = (g0 + i < n_valid) ? 1.f : 0.f;
          const sfu::SoftSig se = sfu::softsig<true>(eta[i]);
          const float av = (se.xc - se.s) + L1[i];       // log psi~ + L1
          const float bv = n1[i] * log_tiny - se.s;      // log1p(-psi~) + L0
          // logaddexp(av, bv) and r = sigmoid(av - bv), no clamp on d
          const float d = av - bv;
          const float td = sfu::ex2(-fabsf(d) * sfu::kLog2e);
          const float ud = 1.0f + td;
          const float invd = sfu::rcp(ud);
          const float rr = (d >= 0.f) ? invd : td * invd;            // P(z = 1 | y)
          const float r = rr * vf;
          const float ell = fmaf(sfu::lg2(ud), sfu::kLn2, fmaxf(av, bv)) * vf;
          const float geta = se.inr ? (rr - se.p) * vf : 0.f;
          acc[0] += ell;
          acc[1] += geta;
#pragma unroll
          for (int k = 0; k < KS; ++k) acc[2 + k] = fmaf(geta, x[k][i], acc[2 + k]);
          acc[1 + KB] = fmaf(r, ga0[i], acc[1 + KB]);
#pragma unroll
          for (int k = 0; k < KO; ++k) acc[2 + KB + k] = fmaf(r, ga[k][i], acc[2 + KB + k]);
        }
      }
#pragma unroll
      for (int i = 0; i < NQ; ++i) acc64[i] += (double)acc[i];
    }
    __syncthreads();
    if (tid == 0 && it + p.nstage < n_it) {
      mbar_expect_tx(&bars[s], tile_bytes);
      tma_load_bulk(stage0 + (size_t)s * tile_elems, packed + (size_t)(bt_begin + it + p.nstage) * tile_elems,
                    tile_bytes, &bars[s]);
    }
  }
  if (chain_ok) {
    double* my = p.partial + ((size_t)blockIdx.x * p.C + c0 + tid) * NQ;
#pragma unroll
    for (int i = 0; i < NQ; ++i) my[i] = acc64[i];
  }
  finish_block<float>(p, c0, ncb, &s_is_last);
}

template <int KS, int KO, int MINB>
static cudaError_t launch_chain_one(const EvalParams& p, dim3 grid, size_t smem, cudaStream_t st, int* occ) {
  auto kern = occu_chain_kernel<KS, KO, MINB>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  if (occ) return cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, kern, kBlockThreads, smem);
  kern<<<grid, kBlockThreads, smem, st>>>(p);
  return cudaGetLastError();
}

// chain-parallel path exists for fp32, no extras, these (Ks, Ko) and J a multiple of... any J
bool occu_chain_supported(int dtype, int ks, int ko, uint32_t flags) {
  if (dtype != BL_F32) return false;
  if (flags & (BL_FLAG_FP_CONSTANT | BL_FLAG_FP_UNOCCUPIED)) return false;
  return (ks == 1 && ko == 1) || (ks == 2 && ko == 1) || (ks == 5 && ko == 3);
}

cudaError_t launch_occu_chain(const EvalParams& p, dim3 grid, size_t smem, cudaStream_t st, int* occ) {
  const int ks = p.L.ks, ko = p.L.ko;
  if (ks == 1 && ko == 1) return launch_chain_one<1, 1, 2>(p, grid, smem, st, occ);
  if (ks == 2 && ko == 1) return launch_chain_one<2, 1, 2>(p, grid, smem, st, occ);
  if (ks == 5 && ko == 3) return launch_chain_one<5, 3, 2>(p, grid, smem, st, occ);
  return cudaErrorNotSupported;
}

}  // namespace bl
