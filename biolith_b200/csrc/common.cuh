// Shared device helpers: TMA bulk-copy / mbarrier PTX, math traits, warp reductions, and the
// dataset handle.  sm_100a only.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <cfloat>
#include <cmath>

#include "../../include/biolith_b200.h"

namespace bl {

constexpr int kWarp = 32;
constexpr int kBlockThreads = 256;            // 8 warps, arranged WS site-groups x WC chain-groups
constexpr int kWarpsPerBlock = kBlockThreads / kWarp;
constexpr int kMaxChainsPerBlock = 256;       // chain chunk handled by one block (grid.y splits the rest)
constexpr int kMaxCov = 16;                   // generic kernel: Ks, Ko <= 16
constexpr int kMaxStages = 4;

// ------------------------------------------------------------------------------------------
// packed dataset layout (built once by pack.cu, consumed by every eval kernel)
//
//   unit u = s * P + p                      (site-major, period minor; each unit is an independent
//                                            marginalisation over z / N, occu.py:204-210)
//   warp-tile t = u / 32, lane = u % 32
//   tile t occupies F * 32 elements:  field f of lane l at  t*F*32 + f*32 + l     ("SoA in tile")
//   fields:  [0, Ks)                 X_k               (NaN -> 0, +-inf -> +-max)
//            [Ks, Ks + J*Ko)         W_{j,k} at Ks + j*Ko + k
//            occu / occu_rn:  NW = ceil(J/32) words of y bits, NW words of mask bits, then n1 = number of
//                             unmasked detections as a float (data-only; z=0 branch / k=0 state)
//            occu_cop:        J floats y (0 if masked), J floats T (0 if masked), Sy, ST, the per-unit
//                             data constant sum m (y log T - lgamma(y+1)), NW mask words
//            nmixture:        J floats y (0 if masked), NW mask words, max count, Y0 = sum m y,
//                             Yw_k = sum_j m y W_jk (Ko floats; data-only parts of the alpha gradient)
//            occu_cs:         J floats score (0 if masked), NW mask words
//   every warp-wide read of one field is one coalesced 128-byte (fp32) line, and a whole tile is a
//   contiguous, 16-byte aligned chunk -> one cp.async.bulk (TMA) per block-tile.
// ------------------------------------------------------------------------------------------
struct Layout {
  int ks, ko, J, P;
  int nw;          // bit words per unit
  int F;           // fields per unit
  int off_w;       // first W field
  int off_y;       // y bits (occu/rn) or y floats (cop)
  int off_m;       // mask bit words
  int off_t;       // cop: T floats
  int off_sy;      // cop: sum of masked y; +1 = sum of masked T; +2 = sum m (y log T - lgamma(y+1))
  int off_n1;      // occu/rn: float count of unmasked detections
  int64_t n_units; // S*P
  int64_t n_tiles; // ceil(n_units/32)
  int64_t n_tiles_padded;  // multiple of kWarpsPerBlock so any block-tile TMA stays in bounds
};

inline Layout make_layout(int model, int64_t S, int P, int J, int ks, int ko) {
  Layout L{};
  L.ks = ks; L.ko = ko; L.J = J; L.P = P;
  L.nw = (J + 31) / 32;
  L.off_w = ks;
  int f = ks + J * ko;
  if (model == BL_MODEL_NMIXTURE) {
    L.off_y = f; f += J;
    L.off_m = f; f += L.nw;
    L.off_sy = f; f += 2 + ko;  // max count, Y0, Yw[ko]
    L.off_t = -1; L.off_n1 = -1;
  } else if (model == BL_MODEL_OCCU_CS) {
    L.off_y = f; f += J;
    L.off_m = f; f += L.nw;
    L.off_t = -1; L.off_sy = -1; L.off_n1 = -1;
  } else if (model == BL_MODEL_OCCU_COP) {
    L.off_y = f; f += J;
    L.off_t = f; f += J;
    L.off_sy = f; f += 3;
    L.off_m = f; f += L.nw;
    L.off_n1 = -1;
  } else {
    L.off_y = f; f += L.nw;
    L.off_m = f; f += L.nw;
    L.off_n1 = f; f += 1;
    L.off_t = -1; L.off_sy = -1;
  }
  L.F = f;
  L.n_units = S * (int64_t)P;
  L.n_tiles = (L.n_units + 31) / 32;
  L.n_tiles_padded = (L.n_tiles + kWarpsPerBlock - 1) / kWarpsPerBlock * kWarpsPerBlock;
  return L;
}

// ------------------------------------------------------------------------------------------
// math traits: constants follow numpyro's clamp_probs(finfo.tiny, 1 - finfo.eps) per dtype
// ------------------------------------------------------------------------------------------
template <typename T> struct Num;

template <> struct Num<float> {
  using bits_t = uint32_t;
  static __device__ __forceinline__ float log_tiny() { return -87.33654475055310898657f; }   // log(FLT_MIN)
  static __device__ __forceinline__ float log_eps() { return -15.94238515333216719f; }       // log(FLT_EPSILON)
  static __device__ __forceinline__ float log1m_eps() { return -1.1920929665620861e-07f; }   // log1p(-eps)
  static __device__ __forceinline__ float neg_tiny() { return -FLT_MIN; }                    // log1p(-tiny)
  static __device__ __forceinline__ float exp_(float x) { return expf(x); }
  static __device__ __forceinline__ float log_(float x) { return logf(x); }
  static __device__ __forceinline__ float log1p_(float x) { return log1pf(x); }
  static __device__ __forceinline__ float expm1_(float x) { return expm1f(x); }
  static __device__ __forceinline__ float abs_(float x) { return fabsf(x); }
  static __device__ __forceinline__ float max_(float a, float b) { return fmaxf(a, b); }
  static __device__ __forceinline__ float min_(float a, float b) { return fminf(a, b); }
  static __device__ __forceinline__ float fma_(float a, float b, float c) { return fmaf(a, b, c); }
  static __device__ __forceinline__ float rcp_(float x) { return 1.0f / x; }
  static __device__ __forceinline__ uint32_t as_bits(float x) { return __float_as_uint(x); }
  static __device__ __forceinline__ float inf() { return __int_as_float(0x7f800000); }
};

template <> struct Num<double> {
  using bits_t = uint64_t;
  static __device__ __forceinline__ double log_tiny() { return -708.3964185322641; }          // log(DBL_MIN)
  static __device__ __forceinline__ double log_eps() { return -36.04365338911715; }           // log(DBL_EPSILON)
  static __device__ __forceinline__ double log1m_eps() { return -2.2204460492503136e-16; }
  static __device__ __forceinline__ double neg_tiny() { return -DBL_MIN; }
  static __device__ __forceinline__ double exp_(double x) { return exp(x); }
  static __device__ __forceinline__ double log_(double x) { return log(x); }
  static __device__ __forceinline__ double log1p_(double x) { return log1p(x); }
  static __device__ __forceinline__ double expm1_(double x) { return expm1(x); }
  static __device__ __forceinline__ double abs_(double x) { return fabs(x); }
  static __device__ __forceinline__ double max_(double a, double b) { return fmax(a, b); }
  static __device__ __forceinline__ double min_(double a, double b) { return fmin(a, b); }
  static __device__ __forceinline__ double fma_(double a, double b, double c) { return fma(a, b, c); }
  static __device__ __forceinline__ double rcp_(double x) { return 1.0 / x; }
  static __device__ __forceinline__ uint32_t as_bits(double x) { return (uint32_t)__double_as_longlong(x); }
  static __device__ __forceinline__ double inf() { return __longlong_as_double(0x7ff0000000000000LL); }
};

// Clamped log-sigmoid pair in log space (oracle/occupancy.py:_clamped_log_sigmoid_pair):
//   p = sigmoid(x), p~ = clip(p, tiny, 1-eps);  lp = log p~, l1mp = log1p(-p~),
//   in-range flag `inr` (derivatives are zero outside, like jnp.clip).
template <typename T>
struct LogSig {
  T p, q;      // sigmoid(x), sigmoid(-x)
  T lp, l1mp;  // clamped logs
  bool inr;
};

template <typename T>
__device__ __forceinline__ LogSig<T> log_sigmoid_pair(T x) {
  using N = Num<T>;
  LogSig<T> r;
  const T t = N::exp_(-N::abs_(x));   // in (0, 1]
  const T l = N::log1p_(t);
  const T inv = N::rcp_(T(1) + t);
  const T ti = t * inv;
  const bool pos = x >= T(0);
  r.p = pos ? inv : ti;
  r.q = pos ? ti : inv;
  T lp = N::min_(x, T(0)) - l;         // log sigmoid(x)
  T l1 = -N::max_(x, T(0)) - l;        // log sigmoid(-x)
  const bool lo = lp <= N::log_tiny();
  const bool hi = l1 <= N::log_eps();
  r.inr = !(lo || hi);
  r.lp = lo ? N::log_tiny() : (hi ? N::log1m_eps() : lp);
  r.l1mp = lo ? N::neg_tiny() : (hi ? N::log_eps() : l1);
  return r;
}

// ------------------------------------------------------------------------------------------
// Bounded-error SFU math (fp32).  NOT -use_fast_math: three explicitly chosen MUFU approximations
// whose errors are *absolute* and below the fp32 rounding already present in the sums they feed
// (DESIGN.md "numerics"): ex2.approx (rel 2^-22), lg2.approx on (1,2] (abs 2^-22), rcp.approx on
// (1,2] (rel 2^-23).  softsig() returns softplus / sigmoid of the CLAMPED argument: clamping x to
// [log tiny, log((1-eps)/eps)] reproduces numpyro's clamp_probs values exactly (log p~ = xc - s,
// log1p(-p~) = -s) and `inr` carries its zero-derivative region.
// ------------------------------------------------------------------------------------------
namespace sfu {
__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr float kXHi = 15.9423847f;     // log((1-eps)/eps): p~ hits 1-eps
constexpr float kXLo = -87.33654475f;   // log(tiny):        p~ hits tiny

struct SoftSig {
  float xc;   // clamped argument
  float s;    // softplus(xc)  = -log1p(-p~)
  float p;    // sigmoid(xc)
  bool inr;   // x inside the clamp range (derivative not zeroed)
};

template <bool CLAMP>
__device__ __forceinline__ SoftSig softsig(float x) {
  SoftSig r;
  r.xc = CLAMP ? fminf(fmaxf(x, kXLo), kXHi) : x;
  r.inr = CLAMP ? (r.xc == x) : true;
  const float t = ex2(-fabsf(r.xc) * kLog2e);   // e^{-|x|} in (0,1]
  const float u = 1.0f + t;
  float inv = rcp(u);
  // one Newton step: MUFU.RCP's error is slightly biased and a bias adds up linearly over 10^7 visits
  // (measured at config 2 near the mode: gradient error 9.0e-6 -> 4.2e-6 of |g|inf for +3 % time)
  inv = fmaf(inv, fmaf(-u, inv, 1.0f), inv);
  r.s = fmaf(lg2(u), kLn2, fmaxf(r.xc, 0.0f));
  r.p = (r.xc >= 0.0f) ? inv : t * inv;
  return r;
}
}  // namespace sfu

// Scalar math selected by (type, SFU): SFU forms are only ever instantiated for float.
template <typename T, bool SFU> struct Mth {
  static __device__ __forceinline__ T exp_(T x) { return Num<T>::exp_(x); }
  static __device__ __forceinline__ T log_(T x) { return Num<T>::log_(x); }
  static __device__ __forceinline__ T rcp_(T x) { return T(1) / x; }
  // softplus(x) and sigmoid(x), libm-accurate
  static __device__ __forceinline__ void softsig(T x, T& s, T& p) {
    const T t = Num<T>::exp_(-Num<T>::abs_(x));
    const T inv = T(1) / (T(1) + t);
    s = Num<T>::max_(x, T(0)) + Num<T>::log1p_(t);
    p = x >= T(0) ? inv : t * inv;
  }
};
template <> struct Mth<float, true> {
  static __device__ __forceinline__ float exp_(float x) { return sfu::ex2(x * sfu::kLog2e); }
  static __device__ __forceinline__ float log_(float x) { return sfu::lg2(x) * sfu::kLn2; }
  static __device__ __forceinline__ float rcp_(float x) { return sfu::rcp(x); }
  static __device__ __forceinline__ void softsig(float x, float& s, float& p) {
    const sfu::SoftSig r = sfu::softsig<false>(x);
    s = r.s;
    p = r.p;
  }
};

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Transposed butterfly: every lane holds P values v[0..P); afterwards the return value of lane l is the
// sum over the 32 lanes of element (l >> (5 - log2 P)), replicated over the low lane bits.  At each of
// the first log2 P steps a lane hands the half of its values it does not keep to its partner, so P
// values cost P - 1 + (5 - log2 P) shuffles instead of 5 P (the SHFL pipe issues one warp-instruction
// per clock, a quarter of the FMA rate: for the 11 sums of the headline shape, 16 instead of 55).
template <int P> struct Log2 { static constexpr int v = 1 + Log2<P / 2>::v; };
template <> struct Log2<1> { static constexpr int v = 0; };

template <typename T, int P>
__device__ __forceinline__ T transpose_reduce(T (&v)[P], int lane) {
  constexpr int LG = Log2<P>::v;
  static_assert(P >= 1 && P <= 32 && (1 << LG) == P, "P must be a power of two <= 32");
#pragma unroll
  for (int s = 0; s < LG; ++s) {
    const int half = P >> (s + 1), o = 16 >> s;
    const bool upper = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const T send = upper ? v[i] : v[i + half];
      const T keep = upper ? v[i + half] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  T r = v[0];
#pragma unroll
  for (int o = 16 >> LG; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
  return r;
}

// ------------------------------------------------------------------------------------------
// TMA (cp.async.bulk) + mbarrier
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void tma_load_bulk(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                              uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  const uint32_t a = smem_u32(bar);
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(a), "r"(parity)
        : "memory");
  } while (!ok);
}

// ------------------------------------------------------------------------------------------
// kernel parameters shared by the three likelihood kernels
// ------------------------------------------------------------------------------------------
struct EvalParams {
  const void* packed;      // packed dataset (element type = compute type)
  const void* theta;       // [C][D]
  void* logp;              // [C]
  double* logp64;          // optional [C] fp64 copy of logp (NUTS energies), may be NULL
  void* grad;              // [C][D]
  double* partial;         // [nsplit][C][NQ] fp64 block partials
  unsigned int* counters;  // [n_chunks] "blocks done" tickets (self-resetting)
  Layout L;
  int model;    // bl_model
  int DS;       // stride of a staged theta row in shared memory (D + derived per-chain slots)
  int C;        // chains in this call
  int D;        // theta dim
  int NQ;       // 1 + D
  int CB;       // chains per block (chunk)
  int WC, WS;   // warp arrangement: WC chain-groups x WS site-groups = 8 warps
  int nstage;
  int nsplit;   // grid.x
  int64_t n_block_tiles;  // ceil(n_tiles / WS)
  uint32_t flags;
  int K;        // occu_rn: max_abundance
  uint32_t rn_scratch_off;  // occu_rn: byte offset of the per-thread A_k scratch in dynamic smem
  void* rn_scratch_global;  // occu_rn: non-NULL -> A_k scratch in global memory (too big for smem)
  double cop_const;  // occu_cop: sum_s sum_j m (y log T - lgamma(y+1)), data-only
  double prior_beta_loc, prior_beta_scale, prior_alpha_loc, prior_alpha_scale;
  double prior_beta_norm, prior_alpha_norm;      // log(scale) + 0.5 log(2 pi), computed on the host
  double prior_beta_iscale, prior_alpha_iscale;  // 1 / scale
  double prior_fp_a, prior_fp_b, prior_fp_rate;
  int chain_variant;  // BL_CHAIN_VARIANT tuning switch as read at plan time (3 = runtime-J kernel)
  int chain_bt;   // lane = chain kernels: threads (= chains) per block of the selected variant
  int nch;        // engine: chains a warp interleaves per pass over a warp-tile (1, 2 or 4)
  int coop_reduce;  // last-block reduction: 1 = one warp per (chain, quantity) when there are <= 64 of them
  int ring_mode;    // K1d: 0 = block barrier per ring slot; 1 = consumers release slots through `empty` mbarriers
  int hier_reduce;  // 1 = two-level (group, then chunk) reduction of the site splits when there are >= 64 of them
  int allreduce;  // 0 none; 1 = leave raw sums in `sums` for a collective, finalize separately
  double* sums;   // [C][NQ] raw (un-prior'd) sums when allreduce != 0
};

}  // namespace bl
