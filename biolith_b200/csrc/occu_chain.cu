// K1c: chain-parallel variant of the occu kernel for large chain batches (C >= 64).
//
//   lane  = chain: every thread keeps ONE chain's beta/alpha in registers and its 1+D running sums
//           (fp32 inside a tile, fp64 across tiles) -> no cross-lane reduction anywhere in the loop;
//   site  = warp-broadcast: all lanes read the same packed fields (conflict-free broadcast LDS.128
//           covering NS consecutive sites of the "SoA in tile" layout at once -> NS-way ILP);
//   block = 256 chains x a contiguous range of site tiles, staged by TMA (cp.async.bulk) through the
//           same mbarrier ring as the site-parallel engine; grid = (site splits, chain chunks).
//   Per tile the y / mask bit words are expanded once, cooperatively, into 0/1 floats in shared
//   memory (2 values per thread), and a block-uniform flag selects a mask-free inner loop when the
//   tile has no missing visit.
// The arithmetic is the closed form of oracle/occupancy.py:occu_logp_grad (reference:
// biolith/models/occu.py:182-242, regression/linear.py:59-66) with bounded-error SFU math
// (common.cuh sfu::softsig).  fp32, no false-positive extras; everything else uses the engine path.
#include <cstdlib>

#include "engine.cuh"

namespace bl {

constexpr int kChainMaxKs = 8;  // runtime-Ks variants of the chain kernel hold up to 8 site coefficients

template <int NS> struct VecLoad;
template <> struct VecLoad<4> {
  static __device__ __forceinline__ void ld(const float* p, float (&o)[4]) {
    const float4 v = *reinterpret_cast<const float4*>(p);
    o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
  }
};
template <> struct VecLoad<2> {
  static __device__ __forceinline__ void ld(const float* p, float (&o)[2]) {
    const float2 v = *reinterpret_cast<const float2*>(p);
    o[0] = v.x; o[1] = v.y;
  }
};

// One warp-tile (<= 32 sites) for this thread's chain: adds into acc[1 + KB + KA].
template <int KS, int KO, int NS, bool MASKED, int JT>
__device__ __forceinline__ void chain_tile(const float* __restrict__ tile, const float* __restrict__ yfx,
                                           const float* __restrict__ mfx, const Layout& L, int n_valid,
                                           const float (&b)[(KS < 0 ? kChainMaxKs : KS) + 1],
                                           const float (&a)[KO + 1],
                                           float (&acc)[3 + (KS < 0 ? kChainMaxKs : KS) + KO], double& logp64) {
  // KS < 0: runtime number of site covariates (<= kChainMaxKs); the site-level loops are predicated on
  // the warp-uniform k < ks and cost nothing in the visit loop
  constexpr int KSM = KS < 0 ? kChainMaxKs : KS;
  constexpr int KB = KSM + 1;
  const int ks = KS < 0 ? L.ks : KS;
  const int J = JT > 0 ? JT : L.J;
  const float log_tiny = Num<float>::log_tiny();
  for (int g0 = 0; g0 < n_valid; g0 += NS) {
    float eta[NS];
#pragma unroll
    for (int i = 0; i < NS; ++i) eta[i] = b[0];
#pragma unroll
    for (int k = 0; k < KSM; ++k) {
      if (k < ks) {
        float xk[NS];  // re-read (one broadcast LDS.128) for the gradient below instead of held in registers
        VecLoad<NS>::ld(tile + k * kWarp + g0, xk);
#pragma unroll
        for (int i = 0; i < NS; ++i) eta[i] = fmaf(xk[i], b[1 + k], eta[i]);
      }
    }
    float L1[NS], ga0[NS], ga[KO > 0 ? KO : 1][NS];
#pragma unroll
    for (int i = 0; i < NS; ++i) {
      L1[i] = 0.f; ga0[i] = 0.f;
#pragma unroll
      for (int k = 0; k < KO; ++k) ga[k][i] = 0.f;
    }
#pragma unroll(JT > 0 ? JT : (JT < 0 ? -JT : 2))
    for (int j = 0; j < J; ++j) {
      float w[KO > 0 ? KO : 1][NS], nu[NS], yf[NS], mf[NS];
#pragma unroll
      for (int i = 0; i < NS; ++i) nu[i] = a[0];
#pragma unroll
      for (int k = 0; k < KO; ++k) {
        VecLoad<NS>::ld(tile + (L.off_w + j * KO + k) * kWarp + g0, w[k]);
#pragma unroll
        for (int i = 0; i < NS; ++i) nu[i] = fmaf(w[k][i], a[1 + k], nu[i]);
      }
      VecLoad<NS>::ld(yfx + j * kWarp + g0, yf);
      if (MASKED) VecLoad<NS>::ld(mfx + j * kWarp + g0, mf);
#pragma unroll
      for (int i = 0; i < NS; ++i) {
        const sfu::SoftSig ss = sfu::softsig<true>(nu[i]);
        // log-lik term  y*xc - softplus(xc);  d/dnu = y - p  (zero outside the clamp range)
        float term = fmaf(yf[i], ss.xc, -ss.s);
        float g = ss.inr ? (yf[i] - ss.p) : 0.f;
        if (MASKED) { term *= mf[i]; g *= mf[i]; }
        L1[i] += term;
        ga0[i] += g;
#pragma unroll
        for (int k = 0; k < KO; ++k) ga[k][i] = fmaf(g, w[k][i], ga[k][i]);
      }
    }
    float n1[NS];
    VecLoad<NS>::ld(tile + L.off_n1 * kWarp + g0, n1);
    float geta[NS];
#pragma unroll
    for (int i = 0; i < NS; ++i) {
      const float vf = (g0 + i < n_valid) ? 1.f : 0.f;
      const sfu::SoftSig se = sfu::softsig<true>(eta[i]);
      const float av = (se.xc - se.s) + L1[i];   // log psi~ + L1
      const float bv = n1[i] * log_tiny - se.s;  // log1p(-psi~) + L0   (L0 = n1 log tiny)
      // logaddexp(av, bv) and r = sigmoid(av - bv); no clamp on d
      const float d = av - bv;
      const float td = sfu::ex2(-fabsf(d) * sfu::kLog2e);
      const float ud = 1.0f + td;
      const float invd = sfu::rcp(ud);
      const float rr = (d >= 0.f) ? invd : td * invd;  // P(z = 1 | y)
      const float r = rr * vf;
      const float ell = fmaf(sfu::lg2(ud), sfu::kLn2, fmaxf(av, bv)) * vf;
      geta[i] = se.inr ? (rr - se.p) * vf : 0.f;
      logp64 += (double)ell;  // fp64 per unit: NUTS needs energy *differences* of a ~1e6-sized sum
      acc[1] += geta[i];
      acc[1 + KB] = fmaf(r, ga0[i], acc[1 + KB]);
#pragma unroll
      for (int k = 0; k < KO; ++k) acc[2 + KB + k] = fmaf(r, ga[k][i], acc[2 + KB + k]);
    }
#pragma unroll
    for (int k = 0; k < KSM; ++k) {
      if (k < ks) {
        float xk[NS];
        VecLoad<NS>::ld(tile + k * kWarp + g0, xk);
#pragma unroll
        for (int i = 0; i < NS; ++i) acc[2 + k] = fmaf(geta[i], xk[i], acc[2 + k]);
      }
    }
  }
}

template <int KS, int KO, int NS, int MINB, int BT, int JT>
__global__ void __launch_bounds__(BT, MINB) occu_chain_kernel(const __grid_constant__ EvalParams p) {
  // accumulator slots: [0] unused, [1 .. KB] beta (KB = KSM + 1 slots, ks + 1 used), then KA alpha slots
  constexpr int KSM = KS < 0 ? kChainMaxKs : KS;
  constexpr int KB = KSM + 1, KA = KO + 1, NQ = 1 + KB + KA;
  const int ks = KS < 0 ? p.L.ks : KS;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
  float* stage0 = reinterpret_cast<float*>(smem_raw + 128);
  __shared__ int s_is_last;
  const int F = p.L.F, J = p.L.J;
  const uint32_t tile_elems = (uint32_t)F * kWarp;  // one warp-tile per stage
  const uint32_t tile_bytes = tile_elems * sizeof(float);
  float* yfx = stage0 + (size_t)p.nstage * tile_elems;  // [J][32] expanded detections
  float* mfx = yfx + (size_t)J * kWarp;                 // [J][32] expanded mask
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

  if (tid == 0) {
    for (int s = 0; s < p.nstage; ++s) mbar_init(&bars[s], 1);
    fence_mbar_init();
  }
  float b[KB], a[KA];
  {
    const float* th = reinterpret_cast<const float*>(p.theta) + (size_t)(c0 + (chain_ok ? tid : 0)) * p.D;
#pragma unroll
    for (int k = 0; k < KB; ++k) b[k] = (k <= ks) ? th[k] : 0.f;
#pragma unroll
    for (int k = 0; k < KA; ++k) a[k] = th[ks + 1 + k];
  }
  // fp64 running sums of the gradient live in shared memory ([q][tid], one column per thread): frees
  // 2 x (NQ-1) registers (measured: 10.4 -> 9.7 ms); the log-marginal stays in a register pair
  double* g64 = reinterpret_cast<double*>(mfx + (size_t)J * kWarp) + tid;
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

  for (int it = 0; it < n_it; ++it) {
    const int s = it % p.nstage;
    mbar_wait(&bars[s], (uint32_t)((it / p.nstage) & 1));
    const float* tile = stage0 + (size_t)s * tile_elems;
    const int64_t unit0 = (bt_begin + it) * kWarp;
    const int n_valid = (int)max((int64_t)0, min((int64_t)kWarp, p.L.n_units - unit0));
    // expand y / mask bits of this tile to floats, once for the whole block
    int any_masked = 0;
    for (int e = tid; e < J * kWarp; e += BT) {
      const int site = e & 31, j = e >> 5;
      const uint32_t yw = __float_as_uint(tile[(p.L.off_y + (j >> 5)) * kWarp + site]);
      const uint32_t mw = __float_as_uint(tile[(p.L.off_m + (j >> 5)) * kWarp + site]);
      const bool mb = (mw >> (j & 31)) & 1u;
      yfx[e] = ((yw >> (j & 31)) & 1u) ? 1.f : 0.f;
      mfx[e] = mb ? 1.f : 0.f;
      any_masked |= (!mb && site < n_valid);
    }
    any_masked = __syncthreads_or(any_masked);
    float acc[NQ];
#pragma unroll
    for (int i = 0; i < NQ; ++i) acc[i] = 0.f;
    if (warp_on) {  // warps whose 32 lanes all lie past the end of the batch only help stage and expand
      if (any_masked) chain_tile<KS, KO, NS, true, JT>(tile, yfx, mfx, p.L, n_valid, b, a, acc, logp64);
      else chain_tile<KS, KO, NS, false, JT>(tile, yfx, mfx, p.L, n_valid, b, a, acc, logp64);
    }
#pragma unroll
    for (int i = 1; i < NQ; ++i) g64[(size_t)i * BT] += (double)acc[i];
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

template <int KS, int KO, int NS, int MINB, int BT, int JT = 0>
static cudaError_t launch_chain_one(const EvalParams& p, dim3 grid, size_t smem, cudaStream_t st, int* occ) {
  auto kern = occu_chain_kernel<KS, KO, NS, MINB, BT, JT>;
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

// chain-parallel path exists for fp32, no extras, and these (Ks, Ko); any J
bool occu_chain_supported(int dtype, int ks, int ko, uint32_t flags) {
  if (dtype != BL_F32) return false;
  if (flags & (BL_FLAG_FP_CONSTANT | BL_FLAG_FP_UNOCCUPIED)) return false;
  // specialised (Ks, Ko) pairs, plus any Ks <= 8 with Ko in 1..4 through the runtime-Ks variants
  return (ks == 1 && ko == 1) || (ks == 2 && ko == 1) || (ks == 5 && ko == 3) ||
         (ks >= 0 && ks <= kChainMaxKs && ko >= 1 && ko <= 4);
}

// (NS, min blocks/SM) variants of the headline shape; BL_CHAIN_VARIANT picks one for tuning runs
static int chain_variant() {
  const char* e = getenv("BL_CHAIN_VARIANT");  // tuning switch, read when a plan is made (never per launch)
  return e ? atoi(e) : 0;
}

int occu_chain_variant() { return chain_variant(); }

size_t occu_chain_smem(const Layout& L, int nstage, int block_threads) {
  size_t b = 128 + (size_t)nstage * L.F * kWarp * sizeof(float) + 2 * (size_t)L.J * kWarp * sizeof(float);
  b = (b + 15) & ~size_t(15);
  return b + (size_t)(3 + kChainMaxKs + L.ko) * block_threads * sizeof(double);  // fp64 gradient columns
}

// threads per block (= chains per block) for a batch of C chains.  A batch is cut into equal chunks of at
// most one block and warps whose lanes all lie past the end of their chunk skip the arithmetic, so narrow
// blocks waste fewer lanes on ragged batch sizes -- exactly the sizes the NUTS driver compacts to at the
// tail of a run.  Measured on B200 (config 2, ms per evaluation, 128 / 256 threads): C=32 0.47 / 0.95,
// 64 0.79 / 1.44, 128 1.42 / 1.49, 192 2.21 / 2.46, 384 3.67 / 5.24, 640 6.01 / 7.88; on whole multiples
// of 256 the two agree within 1 % (1024: 9.16 for 256 threads with J compile-time) -> 256 only there.
int occu_chain_block_threads(int ks, int ko, int C) {
  (void)ks; (void)ko;
  if (chain_variant() == 2) return 128;
  if (chain_variant() == 3) return 256;
  return (C > 0 && C % 256 == 0) ? 256 : 128;
}

template <int KS, int KO, int JT = 0>
static cudaError_t launch_chain_bt(const EvalParams& p, dim3 grid, size_t smem, cudaStream_t st, int* occ) {
  if (p.chain_bt == 128) return launch_chain_one<KS, KO, 4, 4, 128, JT>(p, grid, smem, st, occ);
  return launch_chain_one<KS, KO, 4, 2, 256, JT>(p, grid, smem, st, occ);
}

cudaError_t launch_occu_chain(const EvalParams& p, dim3 grid, size_t smem, cudaStream_t st, int* occ) {
  const int ks = p.L.ks, ko = p.L.ko;
  if (ks == 1 && ko == 1) return launch_chain_bt<1, 1>(p, grid, smem, st, occ);
  if (ks == 2 && ko == 1) return launch_chain_bt<2, 1>(p, grid, smem, st, occ);
  if (ks == 5 && ko == 3) {
    // measured on B200 (config 2, ms per 1024-chain eval): J compile-time (8 visits fully unrolled) 9.16 |
    // runtime J unroll 2: 9.70, unroll 4: 10.35, unroll 1: 10.19 | 128 thr x 4 blocks: 9.79 |
    // 256 thr x 3 blocks (80 regs): 10.8 | 128 thr x 5 blocks (96 regs): 11.4 -> fewer, fatter warps win
    if (p.L.J == 8 && p.chain_variant != 3) return launch_chain_bt<5, 3, 8>(p, grid, smem, st, occ);
    return launch_chain_bt<5, 3>(p, grid, smem, st, occ);
  }
  if (ks >= 0 && ks <= kChainMaxKs) {  // runtime Ks
    if (ko == 1) return launch_chain_bt<-1, 1>(p, grid, smem, st, occ);
    if (ko == 2) return launch_chain_bt<-1, 2>(p, grid, smem, st, occ);
    if (ko == 3) return launch_chain_bt<-1, 3>(p, grid, smem, st, occ);
    if (ko == 4) return launch_chain_bt<-1, 4>(p, grid, smem, st, occ);
  }
  return cudaErrorNotSupported;
}

}  // namespace bl
