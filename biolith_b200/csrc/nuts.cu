// Device-resident, chain-batched NUTS around bl_eval  (SURVEY.md section 8 row f1).
//
// What it replaces: biolith/utils/fit.py:92-130 hands the model to numpyro's MCMC(NUTS(...)); every
// leapfrog there is value_and_grad(potential_fn).  Here the sampler state of all C chains lives in
// HBM and one "global step" is  [bl_eval on the C pending positions] + [nuts_advance_kernel],
// both asynchronous on one stream -- no host round trip per leapfrog.
//
// Algorithm = numpyro's iterative NUTS (numpyro/infer/hmc_util.py: build_tree, _double_tree,
// _iterative_build_subtree, _combine_tree, _is_turning, _leaf_idx_to_ckpt_idxs; hmc.py sample_kernel;
// warmup_adapter with dual averaging (t0=10, kappa=0.75, gamma=0.05), Stan's window schedule and
// regularised Welford diagonal mass matrix), restated as a per-chain state machine: thread = chain,
// each advance consumes exactly one (logp, grad) leaf.  Chains are NOT kept in lock-step by draw:
// every chain spends every global step on a useful leapfrog of whatever tree it is building
// (numpyro's vectorized chain_method makes all chains wait for the deepest tree).
// Random numbers: Philox4x32-10 keyed by (seed, chain); statistically, not bitwise, equal to
// numpyro's threefry streams.  Energies and the tree weights are fp64 (a 1M-site log-density is
// ~1e6 in magnitude: fp32 would leave delta-energy a resolution of 0.5).
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "engine.cuh"
#include "handle.h"

namespace bl {

constexpr int kNutsMaxDepth = 12;
constexpr int kNutsMaxWindows = 32;

struct NutsParams {
  int C, D, max_depth, num_warmup, num_samples;
  double target_accept, max_delta_energy, init_step_size;
  unsigned long long seed;
  int adapt_step_size, adapt_mass;
  int find_heuristic_step_size;  // numpyro HMC(find_heuristic_step_size=True): search at initialisation
  int n_windows;
  int win_end[kNutsMaxWindows];
  // device buffers
  double* vec;        // [n_vec_fields][D][C]
  double* sc;         // [n_scalar_fields][C]
  int* isc;           // [n_int_fields][C]
  double* ckpt;       // [2][max_depth][D][C]
  void* theta;        // [C][D]  eval input  (dataset dtype)
  const void* grad;   // [C][D]  eval output
  const double* logp; // [C]     eval output (fp64)
  float* samples;     // [num_samples][C][D]
  float* stat_accept; // [num_samples][C]
  int* stat_steps;    // [num_samples][C]
  unsigned char* stat_div;  // [num_samples][C]
  double* stat_pe;    // [num_samples][C]
  int* n_done;        // [1]
  int* slot;          // [C] row of chain c in the theta / logp / grad batch (finished chains are compacted away)
};

// vector fields ([D] each)
enum VecField {
  V_Z, V_G, V_IMM, V_WMEAN, V_WM2,                                   // current state, inverse mass, Welford
  V_ZL, V_RL, V_GL, V_ZR, V_RR, V_GR, V_ZP, V_GP, V_RSUM,             // main tree
  V_SZL, V_SRL, V_SGL, V_SZR, V_SRR, V_SGR, V_SZP, V_SGP, V_SRSUM,    // subtree under construction
  V_RHALF, V_ZNEW,                                                    // pending leaf
  V_COUNT
};
enum ScField {
  S_U, S_EPS, S_E0, S_WEIGHT, S_SUMACC, S_UP, S_SWEIGHT, S_SSUMACC, S_SUP, S_DIREPS,
  S_SS_XT, S_SS_XAVG, S_SS_GAVG, S_SS_PROX, S_WN, S_COUNT
};
enum IntField {
  I_IT, I_SAVED, I_DEPTH, I_NUM, I_TURN, I_DIV, I_SNUM, I_STURN, I_SDIV, I_RIGHT, I_DONE, I_WINDOW, I_SS_T,
  I_CTR_LO, I_CTR_HI, I_LEAPS, I_LEAPS_WARM, I_SEARCH, I_SEARCH_DIR, I_COUNT
};

// ---- Philox4x32-10 -----------------------------------------------------------------------
struct Philox {
  unsigned int k0, k1, c2, c3;
  unsigned long long ctr;
  __device__ void next4(unsigned int (&out)[4]) {
    unsigned int c[4] = {(unsigned int)ctr, (unsigned int)(ctr >> 32), c2, c3};
    unsigned int ka = k0, kb = k1;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      const unsigned int hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
      const unsigned int hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
      const unsigned int n0 = hi1 ^ c[1] ^ ka, n2 = hi0 ^ c[3] ^ kb;
      c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
      ka += 0x9E3779B9u; kb += 0xBB67AE85u;
    }
    ++ctr;
    out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
  }
  __device__ double uniform() {  // (0, 1)
    unsigned int o[4];
    next4(o);
    const unsigned long long m = (((unsigned long long)o[0]) << 21) ^ (unsigned long long)(o[1] >> 11);
    return ((double)(m & ((1ull << 53) - 1)) + 0.5) * (1.0 / 9007199254740992.0);
  }
  __device__ void normal2(double& a, double& b) {
    unsigned int o[4];
    next4(o);
    const double u1 = ((double)o[0] + 0.5) * (1.0 / 4294967296.0);
    const double u2 = ((double)o[1] + 0.5) * (1.0 / 4294967296.0);
    const double rad = sqrt(-2.0 * log(u1));
    double s, c;
    sincospi(2.0 * u2, &s, &c);
    a = rad * c;
    b = rad * s;
  }
};

// The state of chain c: element e of a field lives at base[e * stride].  Global view: base = buffer + c, stride = C
// (coalesced across chains); staged view (nuts_advance_staged_kernel): the chain's state copied to shared memory,
// stride = 1.
struct ChainView {
  const NutsParams& p;
  int c;
  double* vec; double* sc; int* isc; double* ckpt;
  size_t stride;
  __device__ ChainView(const NutsParams& p_, int c_)
      : p(p_), c(c_), vec(p_.vec + c_), sc(p_.sc + c_), isc(p_.isc + c_), ckpt(p_.ckpt + c_), stride((size_t)p_.C) {}
  __device__ ChainView(const NutsParams& p_, int c_, double* vec_, double* sc_, int* isc_, double* ckpt_)
      : p(p_), c(c_), vec(vec_), sc(sc_), isc(isc_), ckpt(ckpt_), stride(1) {}
  __device__ double& v(int f, int d) const { return vec[((size_t)f * p.D + d) * stride]; }
  __device__ double& s(int f) const { return sc[(size_t)f * stride]; }
  __device__ int& i(int f) const { return isc[(size_t)f * stride]; }
  __device__ double& ck(int which, int level, int d) const {
    return ckpt[(((size_t)which * p.max_depth + level) * p.D + d) * stride];
  }
  __device__ void copy(int dst, int src) const {
    for (int d = 0; d < p.D; ++d) v(dst, d) = v(src, d);
  }
};

__device__ __forceinline__ double logaddexp_d(double a, double b) {
  const double m = fmax(a, b);
  if (m == -INFINITY) return -INFINITY;
  return m + log1p(exp(-fabs(a - b)));
}

// numpyro _is_turning for a diagonal inverse mass matrix; r_left / r_right / r_sum given as fields
__device__ bool is_turning(const ChainView& cv, int f_left, int f_right, int f_sum) {
  double dl = 0.0, dr = 0.0;
  for (int d = 0; d < cv.p.D; ++d) {
    const double rl = cv.v(f_left, d), rr = cv.v(f_right, d);
    const double rs = cv.v(f_sum, d) - 0.5 * (rl + rr);
    const double im = cv.v(V_IMM, d);
    dl += im * rl * rs;
    dr += im * rr * rs;
  }
  return (dl <= 0.0) || (dr <= 0.0);
}

template <typename T>
__device__ void set_next_leaf(const ChainView& cv) {
  // leapfrog half step from the outer edge of (subtree if it has leaves, else main tree)
  const NutsParams& p = cv.p;
  const bool right = cv.i(I_RIGHT) != 0;
  const bool sub = cv.i(I_SNUM) > 0;
  const int fz = sub ? (right ? V_SZR : V_SZL) : (right ? V_ZR : V_ZL);
  const int fr = sub ? (right ? V_SRR : V_SRL) : (right ? V_RR : V_RL);
  const int fg = sub ? (right ? V_SGR : V_SGL) : (right ? V_GR : V_GL);
  const double de = right ? cv.s(S_EPS) : -cv.s(S_EPS);
  cv.s(S_DIREPS) = de;
  T* th = reinterpret_cast<T*>(p.theta) + (size_t)p.slot[cv.c] * p.D;
  for (int d = 0; d < p.D; ++d) {
    const double rh = cv.v(fr, d) + 0.5 * de * cv.v(fg, d);  // r - (eps/2) dU/dz, dU = -dlogp
    const double zn = cv.v(fz, d) + de * cv.v(V_IMM, d) * rh;
    cv.v(V_RHALF, d) = rh;
    cv.v(V_ZNEW, d) = zn;
    th[d] = (T)zn;
  }
}

__device__ void choose_direction(const ChainView& cv, Philox& rng) {
  cv.i(I_RIGHT) = rng.uniform() < 0.5 ? 1 : 0;
  cv.i(I_SNUM) = 0;
  cv.i(I_STURN) = 0;
  cv.i(I_SDIV) = 0;
}

__device__ void start_transition(const ChainView& cv, Philox& rng) {
  const NutsParams& p = cv.p;
  double kin = 0.0;
  for (int d = 0; d < p.D; d += 2) {
    double n0, n1;
    rng.normal2(n0, n1);
    const double r0 = n0 * rsqrt(cv.v(V_IMM, d));
    cv.v(V_RL, d) = r0;
    kin += 0.5 * cv.v(V_IMM, d) * r0 * r0;
    if (d + 1 < p.D) {
      const double r1 = n1 * rsqrt(cv.v(V_IMM, d + 1));
      cv.v(V_RL, d + 1) = r1;
      kin += 0.5 * cv.v(V_IMM, d + 1) * r1 * r1;
    }
  }
  for (int d = 0; d < p.D; ++d) {
    const double z = cv.v(V_Z, d), g = cv.v(V_G, d), r = cv.v(V_RL, d);
    cv.v(V_ZL, d) = z; cv.v(V_ZR, d) = z; cv.v(V_ZP, d) = z;
    cv.v(V_GL, d) = g; cv.v(V_GR, d) = g; cv.v(V_GP, d) = g;
    cv.v(V_RR, d) = r; cv.v(V_RSUM, d) = r;
  }
  cv.s(S_E0) = cv.s(S_U) + kin;
  cv.s(S_UP) = cv.s(S_U);
  cv.s(S_WEIGHT) = 0.0;
  cv.s(S_SUMACC) = 0.0;
  cv.i(I_DEPTH) = 0; cv.i(I_NUM) = 0; cv.i(I_TURN) = 0; cv.i(I_DIV) = 0;
  choose_direction(cv, rng);
}

// numpyro warmup_adapter.update for transition index t (0-based) with the just-accepted z
__device__ void adapt(const ChainView& cv, int t, double accept_prob) {
  const NutsParams& p = cv.p;
  if (p.adapt_step_size) {
    const double g = p.target_accept - accept_prob;
    const int tt = ++cv.i(I_SS_T);
    const double t0 = 10.0, kappa = 0.75, gamma = 0.05;
    double g_avg = cv.s(S_SS_GAVG);
    g_avg = (1.0 - 1.0 / (tt + t0)) * g_avg + g / (tt + t0);
    const double x_t = cv.s(S_SS_PROX) - sqrt((double)tt) / gamma * g_avg;
    const double w = pow((double)tt, -kappa);
    const double x_avg = (1.0 - w) * cv.s(S_SS_XAVG) + w * x_t;
    cv.s(S_SS_GAVG) = g_avg; cv.s(S_SS_XT) = x_t; cv.s(S_SS_XAVG) = x_avg;
    double eps = (t == p.num_warmup - 1) ? exp(x_avg) : exp(x_t);
    eps = fmin(fmax(eps, 1e-300), 1e300);
    cv.s(S_EPS) = eps;
  }
  const int win = cv.i(I_WINDOW);
  const bool middle = (win > 0) && (win < p.n_windows - 1);
  if (p.adapt_mass && middle) {  // Welford
    const double n = cv.s(S_WN) + 1.0;
    cv.s(S_WN) = n;
    for (int d = 0; d < p.D; ++d) {
      const double x = cv.v(V_Z, d);
      const double pre = x - cv.v(V_WMEAN, d);
      const double mean = cv.v(V_WMEAN, d) + pre / n;
      cv.v(V_WMEAN, d) = mean;
      cv.v(V_WM2, d) += pre * (x - mean);
    }
  }
  const bool at_end = (t == p.win_end[win]);
  if (at_end) cv.i(I_WINDOW) = win + 1;
  if (at_end && middle) {
    if (p.adapt_mass) {
      const double n = cv.s(S_WN);
      for (int d = 0; d < p.D; ++d) {
        double var = cv.v(V_WM2, d) / (n - 1.0);
        var = (n / (n + 5.0)) * var + 1e-3 * (5.0 / (n + 5.0));
        cv.v(V_IMM, d) = var;
        cv.v(V_WMEAN, d) = 0.0;
        cv.v(V_WM2, d) = 0.0;
      }
      cv.s(S_WN) = 0.0;
    }
    if (p.adapt_step_size) {  // ss_init(log(10 * step_size))
      cv.s(S_SS_PROX) = log(10.0 * cv.s(S_EPS));
      cv.s(S_SS_XT) = 0.0; cv.s(S_SS_XAVG) = 0.0; cv.s(S_SS_GAVG) = 0.0;
      cv.i(I_SS_T) = 0;
    }
  }
}

// numpyro find_reasonable_step_size (hmc_util.py): single leapfrogs from the current point with fresh
// momentum, doubling / halving the step size until the direction of "accept_prob vs target" flips.
// I_SEARCH = 1 while searching; I_SEARCH_DIR = last direction (0 = none yet).
template <typename T>
__device__ void search_next_leaf(const ChainView& cv, Philox& rng) {
  const NutsParams& p = cv.p;
  double kin = 0.0;
  for (int d = 0; d < p.D; d += 2) {
    double n0, n1;
    rng.normal2(n0, n1);
    const double r0 = n0 * rsqrt(cv.v(V_IMM, d));
    cv.v(V_RL, d) = r0;
    kin += 0.5 * cv.v(V_IMM, d) * r0 * r0;
    if (d + 1 < p.D) {
      const double r1 = n1 * rsqrt(cv.v(V_IMM, d + 1));
      cv.v(V_RL, d + 1) = r1;
      kin += 0.5 * cv.v(V_IMM, d + 1) * r1 * r1;
    }
  }
  cv.s(S_E0) = cv.s(S_U) + kin;
  const double de = cv.s(S_EPS);
  cv.s(S_DIREPS) = de;
  T* th = reinterpret_cast<T*>(p.theta) + (size_t)p.slot[cv.c] * p.D;
  for (int d = 0; d < p.D; ++d) {
    const double rh = cv.v(V_RL, d) + 0.5 * de * cv.v(V_G, d);
    const double zn = cv.v(V_Z, d) + de * cv.v(V_IMM, d) * rh;
    cv.v(V_RHALF, d) = rh;
    cv.v(V_ZNEW, d) = zn;
    th[d] = (T)zn;
  }
}

// First call: (logp, grad) at the initial positions are in place; set up every chain.
template <typename T>
__global__ void nuts_start_kernel(NutsParams p) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= p.C) return;
  ChainView cv(p, c);
  p.slot[c] = c;
  Philox rng{(unsigned int)p.seed, (unsigned int)(p.seed >> 32), (unsigned int)c, 0x6e757473u, 0ull};
  const T* th = reinterpret_cast<const T*>(p.theta) + (size_t)c * p.D;
  const T* gr = reinterpret_cast<const T*>(p.grad) + (size_t)c * p.D;
  for (int d = 0; d < p.D; ++d) {
    cv.v(V_Z, d) = (double)th[d];
    cv.v(V_G, d) = (double)gr[d];
    cv.v(V_IMM, d) = 1.0;
    cv.v(V_WMEAN, d) = 0.0;
    cv.v(V_WM2, d) = 0.0;
  }
  cv.s(S_U) = -p.logp[c];
  cv.s(S_EPS) = p.init_step_size;
  cv.s(S_SS_PROX) = log(10.0 * p.init_step_size);
  cv.s(S_SS_XT) = 0.0; cv.s(S_SS_XAVG) = 0.0; cv.s(S_SS_GAVG) = 0.0; cv.s(S_WN) = 0.0;
  cv.i(I_IT) = 0; cv.i(I_SAVED) = 0; cv.i(I_DONE) = 0; cv.i(I_WINDOW) = 0; cv.i(I_SS_T) = 0;
  cv.i(I_LEAPS) = 0; cv.i(I_LEAPS_WARM) = 0;
  cv.i(I_SEARCH) = 0; cv.i(I_SEARCH_DIR) = 0;
  if (p.find_heuristic_step_size && p.adapt_step_size && p.num_warmup > 0) {
    cv.i(I_SEARCH) = 1;
    search_next_leaf<T>(cv, rng);
  } else {
    start_transition(cv, rng);
    set_next_leaf<T>(cv);
  }
  cv.i(I_CTR_LO) = (int)(unsigned int)rng.ctr;
  cv.i(I_CTR_HI) = (int)(unsigned int)(rng.ctr >> 32);
}

// One transition step of chain cv.c: consumes the (logp, grad) leaf of its pending position.
template <typename T>
__device__ void nuts_advance_chain(const NutsParams& p, const ChainView& cv) {
  const int c = cv.c;
  if (cv.i(I_DONE)) return;
  const int D = p.D;
  Philox rng{(unsigned int)p.seed, (unsigned int)(p.seed >> 32), (unsigned int)c, 0x6e757473u,
             ((unsigned long long)(unsigned int)cv.i(I_CTR_HI) << 32) | (unsigned int)cv.i(I_CTR_LO)};
  const bool warm = cv.i(I_IT) < p.num_warmup;
  cv.i(I_LEAPS) += 1;
  if (warm) cv.i(I_LEAPS_WARM) += 1;
  if (cv.i(I_SEARCH)) {
    // one search leapfrog finished: delta energy -> direction; continue while the direction repeats
    const int row_s = p.slot[c];
    const T* gs = reinterpret_cast<const T*>(p.grad) + (size_t)row_s * D;
    const double de_s = cv.s(S_DIREPS);
    double kin_s = 0.0;
    for (int d = 0; d < D; ++d) {
      const double rn = cv.v(V_RHALF, d) + 0.5 * de_s * (double)gs[d];
      kin_s += 0.5 * cv.v(V_IMM, d) * rn * rn;
    }
    double delta_s = (-p.logp[row_s] + kin_s) - cv.s(S_E0);
    if (isnan(delta_s)) delta_s = INFINITY;
    const int dir = (p.target_accept < exp(-delta_s)) ? 1 : -1;
    const int last = cv.i(I_SEARCH_DIR);
    double eps = cv.s(S_EPS);
    if ((last == 0 || dir == last) && eps > 1e-30 && eps < 1e30) {
      eps *= (dir > 0) ? 2.0 : 0.5;
      cv.s(S_EPS) = eps;
      cv.i(I_SEARCH_DIR) = dir;
      search_next_leaf<T>(cv, rng);
    } else {
      cv.i(I_SEARCH) = 0;
      cv.s(S_SS_PROX) = log(10.0 * eps);  // ss_init(log(10 * step_size))
      start_transition(cv, rng);
      set_next_leaf<T>(cv);
    }
    cv.i(I_CTR_LO) = (int)(unsigned int)rng.ctr;
    cv.i(I_CTR_HI) = (int)(unsigned int)(rng.ctr >> 32);
    return;
  }

  // ---- 1. finish the leapfrog at z_new, build the leaf (numpyro _build_basetree)
  const int row = p.slot[c];
  const T* gr = reinterpret_cast<const T*>(p.grad) + (size_t)row * D;
  const double de = cv.s(S_DIREPS);
  const double U_new = -p.logp[row];
  double kin = 0.0;
  // r_new is written straight into the subtree's outer edge slot after the combine decision;
  // stage it in V_RHALF (in place)
  for (int d = 0; d < D; ++d) {
    const double g = (double)gr[d];
    const double rn = cv.v(V_RHALF, d) + 0.5 * de * g;
    cv.v(V_RHALF, d) = rn;
    kin += 0.5 * cv.v(V_IMM, d) * rn * rn;
  }
  double delta = (U_new + kin) - cv.s(S_E0);
  if (isnan(delta)) delta = INFINITY;
  const double w_leaf = -delta;
  const bool leaf_div = delta > p.max_delta_energy;
  const double leaf_acc = fmin(1.0, exp(-delta));

  // ---- 2. fold the leaf into the subtree (uniform transition kernel)
  const bool right = cv.i(I_RIGHT) != 0;
  const int snum = cv.i(I_SNUM);
  bool take;
  if (snum == 0) {
    take = true;
    cv.s(S_SWEIGHT) = w_leaf;
    cv.s(S_SSUMACC) = leaf_acc;
    for (int d = 0; d < D; ++d) {
      const double zn = cv.v(V_ZNEW, d), rn = cv.v(V_RHALF, d), g = (double)gr[d];
      cv.v(V_SZL, d) = zn; cv.v(V_SRL, d) = rn; cv.v(V_SGL, d) = g;
      cv.v(V_SZR, d) = zn; cv.v(V_SRR, d) = rn; cv.v(V_SGR, d) = g;
      cv.v(V_SRSUM, d) = rn;
    }
  } else {
    const double sw = cv.s(S_SWEIGHT);
    const double tp = 1.0 / (1.0 + exp(-(w_leaf - sw)));  // expit(new.weight - current.weight)
    take = rng.uniform() < tp;
    cv.s(S_SWEIGHT) = logaddexp_d(sw, w_leaf);
    cv.s(S_SSUMACC) += leaf_acc;
    const int fz = right ? V_SZR : V_SZL, fr = right ? V_SRR : V_SRL, fg = right ? V_SGR : V_SGL;
    for (int d = 0; d < D; ++d) {
      const double rn = cv.v(V_RHALF, d);
      cv.v(fz, d) = cv.v(V_ZNEW, d); cv.v(fr, d) = rn; cv.v(fg, d) = (double)gr[d];
      cv.v(V_SRSUM, d) += rn;
    }
  }
  if (take) {
    for (int d = 0; d < D; ++d) { cv.v(V_SZP, d) = cv.v(V_ZNEW, d); cv.v(V_SGP, d) = (double)gr[d]; }
    cv.s(S_SUP) = U_new;
  }
  cv.i(I_SDIV) = leaf_div ? 1 : 0;
  cv.i(I_SNUM) = snum + 1;
  // checkpoints / iterative U-turn check (numpyro _leaf_idx_to_ckpt_idxs, _is_iterative_turning)
  {
    const int leaf_idx = snum;
    const int idx_max = __popc((unsigned int)(leaf_idx >> 1));
    const int num_sub = __ffs(~(unsigned int)leaf_idx) - 1;  // trailing ones
    const int idx_min = idx_max - num_sub + 1;
    if ((leaf_idx & 1) == 0) {
      for (int d = 0; d < D; ++d) {
        cv.ck(0, idx_max, d) = cv.v(V_RHALF, d);
        cv.ck(1, idx_max, d) = cv.v(V_SRSUM, d);
      }
    } else {
      bool turning = false;
      for (int i = idx_max; i >= idx_min && !turning; --i) {
        double dl = 0.0, dr = 0.0;
        for (int d = 0; d < D; ++d) {
          const double rl = cv.ck(0, i, d), rr = cv.v(V_RHALF, d);
          const double sub = cv.v(V_SRSUM, d) - cv.ck(1, i, d) + rl;
          const double rs = sub - 0.5 * (rl + rr);
          const double im = cv.v(V_IMM, d);
          dl += im * rl * rs;
          dr += im * rr * rs;
        }
        turning = (dl <= 0.0) || (dr <= 0.0);
      }
      cv.i(I_STURN) = turning ? 1 : 0;
    }
  }

  // ---- 3. subtree finished?  (numpyro _iterative_build_subtree cond_fn; then _combine_tree biased)
  const int depth = cv.i(I_DEPTH);
  if (cv.i(I_SNUM) == (1 << depth) || cv.i(I_STURN) || cv.i(I_SDIV)) {
    const bool sturn = cv.i(I_STURN) != 0, sdiv = cv.i(I_SDIV) != 0;
    if (right) { cv.copy(V_ZR, V_SZR); cv.copy(V_RR, V_SRR); cv.copy(V_GR, V_SGR); }
    else { cv.copy(V_ZL, V_SZL); cv.copy(V_RL, V_SRL); cv.copy(V_GL, V_SGL); }
    for (int d = 0; d < D; ++d) cv.v(V_RSUM, d) += cv.v(V_SRSUM, d);
    const double w = cv.s(S_WEIGHT), sw = cv.s(S_SWEIGHT);
    const double tp = (sturn || sdiv) ? 0.0 : fmin(1.0, exp(sw - w));
    const bool turning = sturn ? true : is_turning(cv, V_RL, V_RR, V_RSUM);
    if (rng.uniform() < tp) {
      cv.copy(V_ZP, V_SZP); cv.copy(V_GP, V_SGP);
      cv.s(S_UP) = cv.s(S_SUP);
    }
    cv.i(I_DEPTH) = depth + 1;
    cv.s(S_WEIGHT) = logaddexp_d(w, sw);
    cv.i(I_DIV) = sdiv ? 1 : 0;
    cv.i(I_TURN) = turning ? 1 : 0;
    cv.s(S_SUMACC) += cv.s(S_SSUMACC);
    cv.i(I_NUM) += cv.i(I_SNUM);
    if (depth + 1 >= p.max_depth || turning || sdiv) {
      // ---- transition finished (numpyro hmc.py sample_kernel)
      const double accept_prob = cv.s(S_SUMACC) / (double)cv.i(I_NUM);
      cv.copy(V_Z, V_ZP); cv.copy(V_G, V_GP);
      cv.s(S_U) = cv.s(S_UP);
      const int t = cv.i(I_IT);
      if (t >= p.num_warmup) {
        const int n = cv.i(I_SAVED);
        float* dst = p.samples + ((size_t)n * p.C + c) * D;
        for (int d = 0; d < D; ++d) dst[d] = (float)cv.v(V_Z, d);
        const size_t o = (size_t)n * p.C + c;
        p.stat_accept[o] = (float)accept_prob;
        p.stat_steps[o] = cv.i(I_NUM);
        p.stat_div[o] = sdiv ? 1 : 0;
        p.stat_pe[o] = cv.s(S_U);
        cv.i(I_SAVED) = n + 1;
      } else {
        adapt(cv, t, accept_prob);
      }
      cv.i(I_IT) = t + 1;
      if (t + 1 >= p.num_warmup + p.num_samples) {
        cv.i(I_DONE) = 1;
        atomicAdd(p.n_done, 1);
        cv.i(I_CTR_LO) = (int)(unsigned int)rng.ctr;
        cv.i(I_CTR_HI) = (int)(unsigned int)(rng.ctr >> 32);
        return;
      }
      start_transition(cv, rng);
    } else {
      choose_direction(cv, rng);
    }
  }
  // ---- 4. next leaf
  set_next_leaf<T>(cv);
  cv.i(I_CTR_LO) = (int)(unsigned int)rng.ctr;
  cv.i(I_CTR_HI) = (int)(unsigned int)(rng.ctr >> 32);
}

template <typename T>
__global__ void nuts_advance_kernel(NutsParams p) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= p.C) return;
  nuts_advance_chain<T>(p, ChainView(p, c));
}

// Few chains (what `fit` with the reference's default num_chains = 5 runs, biolith/utils/fit.py:24): thread-per-chain
// walks ~300 dependent global-memory accesses per step (measured: ~45 us per leapfrog at 5 chains, 40 % of the step
// beside a 63 us evaluation).  Here ONE WARP per chain copies the chain's state (~(27 D + 2 depth D) doubles) into
// shared memory with all its lanes (independent loads, one or two round trips), lane 0 runs the SAME transition code
// on the shared-memory view, and the warp writes the state back.  Same arithmetic in the same order as
// nuts_advance_kernel: the draws are bit-identical.
constexpr int kStagedWarps = 4;

__host__ __device__ inline size_t nuts_staged_doubles(int D, int max_depth) {
  return (size_t)V_COUNT * D + S_COUNT + 2 * (size_t)max_depth * D;
}
__host__ __device__ inline size_t nuts_staged_bytes_per_chain(int D, int max_depth) {
  return nuts_staged_doubles(D, max_depth) * sizeof(double) + ((I_COUNT * sizeof(int) + 7) & ~size_t(7));
}

template <typename T>
__global__ void __launch_bounds__(kStagedWarps * 32) nuts_advance_staged_kernel(NutsParams p) {
  extern __shared__ __align__(16) unsigned char nuts_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * kStagedWarps + warp;
  if (c >= p.C) return;
  if (p.isc[(size_t)I_DONE * p.C + c]) return;  // warp-uniform
  const int D = p.D;
  unsigned char* base = nuts_smem + (size_t)warp * nuts_staged_bytes_per_chain(D, p.max_depth);
  double* s_vec = reinterpret_cast<double*>(base);
  double* s_sc = s_vec + (size_t)V_COUNT * D;
  double* s_ck = s_sc + S_COUNT;
  int* s_i = reinterpret_cast<int*>(s_ck + 2 * (size_t)p.max_depth * D);
  const int nv = V_COUNT * D, nc = 2 * p.max_depth * D;
  const size_t C = (size_t)p.C;
  for (int e = lane; e < nv; e += 32) s_vec[e] = p.vec[(size_t)e * C + c];
  for (int e = lane; e < nc; e += 32) s_ck[e] = p.ckpt[(size_t)e * C + c];
  if (lane < S_COUNT) s_sc[lane] = p.sc[(size_t)lane * C + c];
  if (lane < I_COUNT) s_i[lane] = p.isc[(size_t)lane * C + c];
  static_assert(S_COUNT <= 32 && I_COUNT <= 32, "one lane per scalar field");
  __syncwarp();
  if (lane == 0) nuts_advance_chain<T>(p, ChainView(p, c, s_vec, s_sc, s_i, s_ck));
  __syncwarp();
  for (int e = lane; e < nv; e += 32) p.vec[(size_t)e * C + c] = s_vec[e];
  for (int e = lane; e < nc; e += 32) p.ckpt[(size_t)e * C + c] = s_ck[e];
  if (lane < S_COUNT) p.sc[(size_t)lane * C + c] = s_sc[lane];
  if (lane < I_COUNT) p.isc[(size_t)lane * C + c] = s_i[lane];
}

// Re-number the rows of the evaluation batch so that the still-running chains are contiguous (one
// block; order-preserving, hence identical on every rank of a site-sharded run) and re-emit their
// pending positions into the new rows.
template <typename T>
__global__ void nuts_compact_kernel(NutsParams p) {
  __shared__ int s_base;
  __shared__ int s_warp[32];
  if (threadIdx.x == 0) s_base = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  for (int c0 = 0; c0 < p.C; c0 += blockDim.x) {
    const int c = c0 + threadIdx.x;
    const int active = (c < p.C && !p.isc[(size_t)I_DONE * p.C + c]) ? 1 : 0;
    const unsigned int m = __ballot_sync(0xffffffffu, active);
    if (lane == 0) s_warp[warp] = __popc(m);
    __syncthreads();
    int before = s_base;
    for (int w = 0; w < warp; ++w) before += s_warp[w];
    const int row = before + __popc(m & ((1u << lane) - 1u));
    if (active) {
      p.slot[c] = row;
      T* th = reinterpret_cast<T*>(p.theta) + (size_t)row * p.D;
      for (int d = 0; d < p.D; ++d) th[d] = (T)p.vec[((size_t)V_ZNEW * p.D + d) * p.C + c];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int tot = 0;
      for (int w = 0; w < nwarp; ++w) tot += s_warp[w];
      s_base += tot;
    }
    __syncthreads();
  }
}

}  // namespace bl

using namespace bl;

struct bl_nuts {
  bl_dataset* ds = nullptr;
  NutsParams p{};
  void* d_grad = nullptr;
  void* d_logp_t = nullptr;  // logp in the dataset dtype (unused by the sampler)
  double* d_logp64 = nullptr;
  cudaStream_t stream = nullptr;
  int64_t steps = 0;
  bool started = false;
  // CUDA graph of `graph_steps` x (eval + advance): the inner loop is launch-bound on small datasets
  cudaGraphExec_t graph_exec = nullptr;
  int graph_steps = 0;
  int n_rows = 0;          // current evaluation batch (active chains rounded up to a chain chunk)
  int64_t rows_evaluated = 0;  // sum over global steps of the batch size actually evaluated (after compaction)
  bool compaction = true;
};

// numpyro build_adaptation_schedule (hmc_util.py): Stan's 75 / 25*2^k / 50 windows
static void build_schedule(int num_steps, std::vector<int>& ends) {
  ends.clear();
  if (num_steps <= 0) return;
  if (num_steps < 20) { ends.push_back(num_steps - 1); return; }
  int start_buffer = 75, end_buffer = 50, init_window = 25;
  if (start_buffer + end_buffer + init_window > num_steps) {
    start_buffer = (int)(0.15 * num_steps);
    end_buffer = (int)(0.1 * num_steps);
    init_window = num_steps - start_buffer - end_buffer;
  }
  ends.push_back(start_buffer - 1);
  const int end_window_start = num_steps - end_buffer;
  int next_size = init_window, next_start = start_buffer;
  while (next_start < end_window_start) {
    int cur_start = next_start, cur_size = next_size;
    if (3 * cur_size <= end_window_start - cur_start) next_size = 2 * cur_size;
    else cur_size = end_window_start - cur_start;
    next_start = cur_start + cur_size;
    ends.push_back(next_start - 1);
  }
  ends.push_back(num_steps - 1);
}

#define CU_TRY(expr)                                                                              \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess) return fail(BL_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(_e));     \
  } while (0)

extern "C" {

int bl_nuts_destroy(bl_nuts* s) {
  if (!s) return BL_OK;
  if (s->ds) cudaSetDevice(s->ds->desc.device);
  cudaFree(s->p.vec); cudaFree(s->p.sc); cudaFree(s->p.isc); cudaFree(s->p.ckpt); cudaFree(s->p.theta);
  cudaFree(s->d_grad); cudaFree(s->d_logp_t); cudaFree(s->d_logp64); cudaFree(s->p.samples);
  cudaFree(s->p.stat_accept); cudaFree(s->p.stat_steps); cudaFree(s->p.stat_div); cudaFree(s->p.stat_pe);
  cudaFree(s->p.n_done); cudaFree(s->p.slot);
  if (s->graph_exec) cudaGraphExecDestroy(s->graph_exec);
  if (s->stream) cudaStreamDestroy(s->stream);
  delete s;
  return BL_OK;
}

int bl_nuts_create(bl_dataset* ds, const bl_nuts_config* cfg, const void* theta0, bl_nuts** out) {
  if (!ds || !cfg || !theta0 || !out) return fail(BL_ERR_INVALID, "NULL argument");
  *out = nullptr;
  if (cfg->n_chains < 1 || cfg->num_samples < 1 || cfg->num_warmup < 0)
    return fail(BL_ERR_INVALID, "n_chains/num_samples/num_warmup out of range");
  if (cfg->max_tree_depth < 1 || cfg->max_tree_depth > kNutsMaxDepth)
    return fail(BL_ERR_INVALID, "max_tree_depth must be in [1, %d]", kNutsMaxDepth);
  if (!(ds->desc.flags & BL_FLAG_PRIOR))
    return fail(BL_ERR_INVALID, "NUTS needs the full potential: create the dataset with BL_FLAG_PRIOR");
  CU_TRY(cudaSetDevice(ds->desc.device));
  bl_nuts* s = new (std::nothrow) bl_nuts();
  if (!s) return fail(BL_ERR_NOMEM, "host allocation failed");
  s->ds = ds;
  NutsParams& p = s->p;
  p.C = cfg->n_chains; p.D = ds->D; p.max_depth = cfg->max_tree_depth;
  p.num_warmup = cfg->num_warmup; p.num_samples = cfg->num_samples;
  p.target_accept = cfg->target_accept_prob > 0 ? cfg->target_accept_prob : 0.8;
  p.max_delta_energy = cfg->max_delta_energy > 0 ? cfg->max_delta_energy : 1000.0;
  p.init_step_size = cfg->init_step_size > 0 ? cfg->init_step_size : 1.0;
  p.seed = cfg->seed;
  p.adapt_step_size = cfg->adapt_step_size; p.adapt_mass = cfg->adapt_mass_matrix;
  p.find_heuristic_step_size = cfg->find_heuristic_step_size;
  std::vector<int> ends;
  build_schedule(p.num_warmup, ends);
  if ((int)ends.size() > kNutsMaxWindows) { delete s; return fail(BL_ERR_INVALID, "too many adaptation windows"); }
  p.n_windows = (int)ends.size();
  for (int i = 0; i < kNutsMaxWindows; ++i) p.win_end[i] = i < p.n_windows ? ends[i] : -1;
  const size_t C = p.C, D = p.D, N = p.num_samples;
  const size_t es = ds->desc.dtype == BL_F32 ? 4 : 8;
  int rc = BL_OK;
  cudaError_t e;
#define CU_NB(expr) if (rc == BL_OK && (e = (expr)) != cudaSuccess) rc = fail(BL_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e))
  CU_NB(cudaMalloc(&p.vec, (size_t)V_COUNT * D * C * sizeof(double)));
  CU_NB(cudaMalloc(&p.sc, (size_t)S_COUNT * C * sizeof(double)));
  CU_NB(cudaMalloc(&p.isc, (size_t)I_COUNT * C * sizeof(int)));
  CU_NB(cudaMalloc(&p.ckpt, 2 * (size_t)p.max_depth * D * C * sizeof(double)));
  CU_NB(cudaMalloc(&p.theta, C * D * es));
  CU_NB(cudaMalloc(&s->d_grad, C * D * es));
  CU_NB(cudaMalloc(&s->d_logp_t, C * es));
  CU_NB(cudaMalloc(&s->d_logp64, C * sizeof(double)));
  CU_NB(cudaMalloc(&p.samples, N * C * D * sizeof(float)));
  CU_NB(cudaMalloc(&p.stat_accept, N * C * sizeof(float)));
  CU_NB(cudaMalloc(&p.stat_steps, N * C * sizeof(int)));
  CU_NB(cudaMalloc(&p.stat_div, N * C));
  CU_NB(cudaMalloc(&p.stat_pe, N * C * sizeof(double)));
  CU_NB(cudaMalloc(&p.n_done, sizeof(int)));
  CU_NB(cudaMalloc(&p.slot, C * sizeof(int)));
  CU_NB(cudaMemset(p.n_done, 0, sizeof(int)));
  CU_NB(cudaMemset(p.isc, 0, (size_t)I_COUNT * C * sizeof(int)));
  CU_NB(cudaMemset(p.vec, 0, (size_t)V_COUNT * D * C * sizeof(double)));
  CU_NB(cudaMemset(p.sc, 0, (size_t)S_COUNT * C * sizeof(double)));
  CU_NB(cudaMemset(p.ckpt, 0, 2 * (size_t)p.max_depth * D * C * sizeof(double)));
  CU_NB(cudaMemcpy(p.theta, theta0, C * D * es, cudaMemcpyHostToDevice));
  CU_NB(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
#undef CU_NB
  if (rc != BL_OK) { bl_nuts_destroy(s); return rc; }
  p.grad = s->d_grad;
  p.logp = s->d_logp64;
  s->n_rows = p.C;
  *out = s;
  return BL_OK;
}

int bl_nuts_run(bl_nuts* s, int64_t max_steps, int32_t poll_every, int64_t* steps_done, int32_t* chains_done) {
  if (!s) return fail(BL_ERR_INVALID, "NULL sampler");
  bl_dataset* ds = s->ds;
  CU_TRY(cudaSetDevice(ds->desc.device));
  const NutsParams& p = s->p;
  const int threads = 128, blocks = (p.C + threads - 1) / threads;
  const bool f32 = ds->desc.dtype == BL_F32;
  if (poll_every < 1) poll_every = 32;
  int done = 0;
  int64_t n = 0;
  if (!s->started) {
    int rc = eval_device(ds, p.theta, p.C, s->d_logp_t, s->d_grad, s->stream, 0, s->d_logp64);
    if (rc) return rc;
    if (f32) nuts_start_kernel<float><<<blocks, threads, 0, s->stream>>>(p);
    else nuts_start_kernel<double><<<blocks, threads, 0, s->stream>>>(p);
    CU_TRY(cudaGetLastError());
    g_launches.fetch_add(1);
    s->started = true;
  }
  // few chains and a state that fits in shared memory: one warp per chain (see nuts_advance_staged_kernel)
  const size_t staged_smem = (size_t)kStagedWarps * nuts_staged_bytes_per_chain(p.D, p.max_depth);
  bool staged = p.C <= 64 && staged_smem <= 96 * 1024;
  if (const char* ev = getenv("BL_NUTS_STAGED")) staged = staged && atoi(ev) != 0;
  if (staged && staged_smem > 48 * 1024) {
    cudaError_t ce = f32 ? cudaFuncSetAttribute(nuts_advance_staged_kernel<float>,
                                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)staged_smem)
                         : cudaFuncSetAttribute(nuts_advance_staged_kernel<double>,
                                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)staged_smem);
    if (ce != cudaSuccess) { cudaGetLastError(); staged = false; }
  }
  const int sblocks = (p.C + kStagedWarps - 1) / kStagedWarps;
  auto one_step = [&]() -> int {
    int rc = eval_device(ds, p.theta, s->n_rows, s->d_logp_t, s->d_grad, s->stream, 0, s->d_logp64);
    if (rc) return rc;
    if (staged) {
      if (f32) nuts_advance_staged_kernel<float><<<sblocks, kStagedWarps * 32, staged_smem, s->stream>>>(p);
      else nuts_advance_staged_kernel<double><<<sblocks, kStagedWarps * 32, staged_smem, s->stream>>>(p);
    } else if (f32) nuts_advance_kernel<float><<<blocks, threads, 0, s->stream>>>(p);
    else nuts_advance_kernel<double><<<blocks, threads, 0, s->stream>>>(p);
    g_launches.fetch_add(1);
    return BL_OK;
  };
  // Capture poll_every x (eval + advance) and replay it (not with a cross-rank exchange attached: the
  // P2P epoch is a kernel argument).  Re-captured whenever compaction shrinks the batch.
  auto capture = [&]() {
    if (s->graph_exec) { cudaGraphExecDestroy(s->graph_exec); s->graph_exec = nullptr; }
    s->graph_steps = 0;
    if (ds->comm) return;
    cudaGraph_t graph = nullptr;
    const int64_t launches_before = g_launches.load();
    if (cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
      int rc = BL_OK;
      for (int k = 0; k < poll_every && rc == BL_OK; ++k) rc = one_step();
      cudaError_t ce = cudaStreamEndCapture(s->stream, &graph);
      if (rc == BL_OK && ce == cudaSuccess && graph &&
          cudaGraphInstantiate(&s->graph_exec, graph, 0) == cudaSuccess)
        s->graph_steps = poll_every;
      if (graph) cudaGraphDestroy(graph);
    }
    cudaGetLastError();
    g_launches.store(launches_before);  // captured launches were not executed
  };
  bool need_capture = (s->graph_exec == nullptr) && max_steps >= poll_every;
  while (n < max_steps) {
    if (need_capture) {
      // one eager step first: creates the plan / workspace of this batch size outside the capture
      int rc = one_step();
      if (rc) return rc;
      ++n;
      s->rows_evaluated += s->n_rows;
      capture();
      need_capture = false;
    }
    if (s->graph_exec && max_steps - n >= s->graph_steps) {
      CU_TRY(cudaGraphLaunch(s->graph_exec, s->stream));
      g_launches.fetch_add(2 * (int64_t)s->graph_steps);
      n += s->graph_steps;
      s->rows_evaluated += (int64_t)s->graph_steps * s->n_rows;
    } else {
      for (int k = 0; k < poll_every && n < max_steps; ++k, ++n) {
        int rc = one_step();
        if (rc) return rc;
        s->rows_evaluated += s->n_rows;
      }
    }
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaMemcpyAsync(&done, p.n_done, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    CU_TRY(cudaStreamSynchronize(s->stream));
    if (done >= p.C) break;
    // compaction: evaluate only the chains that are still running, in whole warps (the lane = chain kernels
    // skip warps past the end of the batch, so a 32-chain granularity is what the tail of a run pays for)
    if (s->compaction) {
      const int active = p.C - done;
      const int rows = std::min(p.C, (active + 31) / 32 * 32);
      if (rows < s->n_rows) {
        if (f32) nuts_compact_kernel<float><<<1, 1024, 0, s->stream>>>(p);
        else nuts_compact_kernel<double><<<1, 1024, 0, s->stream>>>(p);
        CU_TRY(cudaGetLastError());
        g_launches.fetch_add(1);
        s->n_rows = rows;
        need_capture = max_steps - n >= poll_every;
        if (s->graph_exec) { cudaGraphExecDestroy(s->graph_exec); s->graph_exec = nullptr; }
      }
    }
  }
  s->steps += n;
  if (steps_done) *steps_done = s->steps;
  if (chains_done) *chains_done = done;
  return BL_OK;
}

int bl_nuts_rows_evaluated(bl_nuts* s, int64_t* rows) {
  if (!s || !rows) return fail(BL_ERR_INVALID, "NULL argument");
  *rows = s->rows_evaluated;
  return BL_OK;
}

int bl_nuts_get(bl_nuts* s, float* samples, float* accept_prob, int32_t* num_steps, uint8_t* diverging,
                double* potential_energy, double* step_size, double* inv_mass, int32_t* leapfrogs,
                int32_t* warmup_leapfrogs, int32_t* n_saved) {
  if (!s) return fail(BL_ERR_INVALID, "NULL sampler");
  CU_TRY(cudaSetDevice(s->ds->desc.device));
  CU_TRY(cudaStreamSynchronize(s->stream));
  const NutsParams& p = s->p;
  const size_t C = p.C, D = p.D, N = p.num_samples;
  if (samples) CU_TRY(cudaMemcpy(samples, p.samples, N * C * D * sizeof(float), cudaMemcpyDeviceToHost));
  if (accept_prob) CU_TRY(cudaMemcpy(accept_prob, p.stat_accept, N * C * sizeof(float), cudaMemcpyDeviceToHost));
  if (num_steps) CU_TRY(cudaMemcpy(num_steps, p.stat_steps, N * C * sizeof(int), cudaMemcpyDeviceToHost));
  if (diverging) CU_TRY(cudaMemcpy(diverging, p.stat_div, N * C, cudaMemcpyDeviceToHost));
  if (potential_energy) CU_TRY(cudaMemcpy(potential_energy, p.stat_pe, N * C * sizeof(double), cudaMemcpyDeviceToHost));
  if (step_size) CU_TRY(cudaMemcpy(step_size, p.sc + (size_t)S_EPS * C, C * sizeof(double), cudaMemcpyDeviceToHost));
  if (inv_mass) CU_TRY(cudaMemcpy(inv_mass, p.vec + (size_t)V_IMM * D * C, D * C * sizeof(double), cudaMemcpyDeviceToHost));
  if (leapfrogs) CU_TRY(cudaMemcpy(leapfrogs, p.isc + (size_t)I_LEAPS * C, C * sizeof(int), cudaMemcpyDeviceToHost));
  if (warmup_leapfrogs) CU_TRY(cudaMemcpy(warmup_leapfrogs, p.isc + (size_t)I_LEAPS_WARM * C, C * sizeof(int), cudaMemcpyDeviceToHost));
  if (n_saved) CU_TRY(cudaMemcpy(n_saved, p.isc + (size_t)I_SAVED * C, C * sizeof(int), cudaMemcpyDeviceToHost));
  return BL_OK;
}

}  // extern "C"
