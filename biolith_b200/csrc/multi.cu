// Composite dataset for n_species > 1 (reference: the species plate, biolith/models/occu.py:182-186 -- one beta / alpha
// row per species over the SAME covariates -- with the false-positive / score parameters sampled ONCE, outside the
// plate: occu.py:146-157, occu_rn.py:133-137, occu_cop.py:160-171, occu_cs.py:146-154).  Species are conditionally
// independent given the shared extras, so the joint log-density is the sum of the single-species likelihoods at
//     theta_sp = [ beta_sp | alpha_sp | extras ]   taken from   theta = [ beta (Sp x Kb) | alpha (Sp x Ka) | extras ],
// plus every prior once; the gradient of a shared extra is the sum over species.  One child handle per species (their
// own kernels, likelihood only), a gather kernel in front and a combine kernel behind, all on the caller's stream.
#include <vector>

#include "engine.cuh"
#include "handle.h"

namespace bl {

template <typename T>
__global__ void ms_gather_kernel(const T* __restrict__ theta, T* __restrict__ theta_sp, int C, int D, int Dsp, int KB,
                                 int KA, int E, int sp, int Sp) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= C * Dsp) return;
  const int c = idx / Dsp, i = idx % Dsp;
  int src;
  if (i < KB) src = sp * KB + i;
  else if (i < KB + KA) src = Sp * KB + sp * KA + (i - KB);
  else src = Sp * (KB + KA) + (i - KB - KA);
  theta_sp[idx] = theta[(size_t)c * D + src];
}

// one thread per (chain, entry of the full theta); entry D stands for the log-density
template <typename T>
__global__ void ms_combine_kernel(const EvalParams p, int Sp, int Dsp, const double* __restrict__ lp64_sp,
                                  const T* __restrict__ g_sp) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int D = p.D, C = p.C;
  if (idx >= C * (D + 1)) return;
  const int c = idx / (D + 1), i = idx % (D + 1);
  const int KB = p.L.ks + 1, KA = p.L.ko + 1, E = D - Sp * (KB + KA);
  const T* theta = reinterpret_cast<const T*>(p.theta) + (size_t)c * D;
  const bool prior = (p.flags & BL_FLAG_PRIOR) != 0;
  const double h2pi = 0.91893853320467274178;
  if (i == D) {
    double lp = 0.0;
    for (int sp = 0; sp < Sp; ++sp) lp += lp64_sp[(size_t)sp * C + c];
    if (prior) {
      for (int k = 0; k < Sp * KB; ++k) {
        const double z = ((double)theta[k] - p.prior_beta_loc) / p.prior_beta_scale;
        lp += -0.5 * z * z - log(p.prior_beta_scale) - h2pi;
      }
      for (int k = 0; k < Sp * KA; ++k) {
        const double z = ((double)theta[Sp * KB + k] - p.prior_alpha_loc) / p.prior_alpha_scale;
        lp += -0.5 * z * z - log(p.prior_alpha_scale) - h2pi;
      }
      const T* ex = theta + Sp * (KB + KA);
      if (p.model == BL_MODEL_OCCU_CS) {
        const double x[4] = {(double)ex[0], (double)ex[1], (double)ex[2], (double)ex[3]};
        lp += cs_prior(p, x, -1);
      } else {
        for (int e = 0; e < E; ++e) {
          const double x = (double)ex[e];
          if (p.model == BL_MODEL_OCCU_COP) lp += log(p.prior_fp_rate) - p.prior_fp_rate * exp(x) + x;
          else lp += p.prior_fp_a * log_sigmoid_d(x) + p.prior_fp_b * log_sigmoid_d(-x) + lgamma(p.prior_fp_a + p.prior_fp_b) -
                     lgamma(p.prior_fp_a) - lgamma(p.prior_fp_b);
        }
      }
    }
    reinterpret_cast<T*>(p.logp)[c] = (T)lp;
    if (p.logp64) p.logp64[c] = lp;
    return;
  }
  double g;
  const double x = (double)theta[i];
  if (i < Sp * KB) {
    const int sp = i / KB, k = i % KB;
    g = (double)g_sp[((size_t)sp * C + c) * Dsp + k];
    if (prior) g -= (x - p.prior_beta_loc) / (p.prior_beta_scale * p.prior_beta_scale);
  } else if (i < Sp * (KB + KA)) {
    const int j = i - Sp * KB, sp = j / KA, k = j % KA;
    g = (double)g_sp[((size_t)sp * C + c) * Dsp + KB + k];
    if (prior) g -= (x - p.prior_alpha_loc) / (p.prior_alpha_scale * p.prior_alpha_scale);
  } else {
    const int e = i - Sp * (KB + KA);
    g = 0.0;
    for (int sp = 0; sp < Sp; ++sp) g += (double)g_sp[((size_t)sp * C + c) * Dsp + KB + KA + e];
    if (prior) {
      if (p.model == BL_MODEL_OCCU_CS) {
        const T* ex = theta + Sp * (KB + KA);
        const double xe[4] = {(double)ex[0], (double)ex[1], (double)ex[2], (double)ex[3]};
        g += cs_prior(p, xe, e);
      } else if (p.model == BL_MODEL_OCCU_COP) {
        g += 1.0 - p.prior_fp_rate * exp(x);
      } else {
        const double c1 = 1.0 / (1.0 + exp(-x));
        g += p.prior_fp_a * (1.0 - c1) - p.prior_fp_b * c1;
      }
    }
  }
  reinterpret_cast<T*>(p.grad)[(size_t)c * D + i] = (T)g;
}

int eval_device_multi(bl_dataset* ds, const void* theta, int C, void* logp, void* grad, cudaStream_t st,
                      double* logp64, void (*fill)(const bl_dataset*, EvalParams&)) {
  const int Sp = (int)ds->species.size();
  bl_dataset* c0 = ds->species[0];
  const int Dsp = c0->D, KB = c0->L.ks + 1, KA = c0->L.ko + 1, E = Dsp - KB - KA;
  const size_t es = ds->desc.dtype == BL_F32 ? 4 : 8;
  if (C > ds->ms_cap) {
    cudaDeviceSynchronize();
    cudaFree(ds->ms_theta); cudaFree(ds->ms_lp); cudaFree(ds->ms_lp64); cudaFree(ds->ms_grad);
    ds->ms_theta = ds->ms_lp = ds->ms_grad = nullptr; ds->ms_lp64 = nullptr; ds->ms_cap = 0;
    if (cudaMalloc(&ds->ms_theta, (size_t)C * Dsp * es) != cudaSuccess ||
        cudaMalloc(&ds->ms_lp, (size_t)Sp * C * es) != cudaSuccess ||
        cudaMalloc(&ds->ms_lp64, (size_t)Sp * C * sizeof(double)) != cudaSuccess ||
        cudaMalloc(&ds->ms_grad, (size_t)Sp * C * Dsp * es) != cudaSuccess)
      return fail(BL_ERR_NOMEM, "multi-species workspace");
    ds->ms_cap = C;
  }
  const int th = 256;
  for (int sp = 0; sp < Sp; ++sp) {
    const int n = C * Dsp;
    if (es == 4)
      ms_gather_kernel<float><<<(n + th - 1) / th, th, 0, st>>>((const float*)theta, (float*)ds->ms_theta, C, ds->D, Dsp,
                                                                 KB, KA, E, sp, Sp);
    else
      ms_gather_kernel<double><<<(n + th - 1) / th, th, 0, st>>>((const double*)theta, (double*)ds->ms_theta, C, ds->D,
                                                                  Dsp, KB, KA, E, sp, Sp);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    int rc = eval_device(ds->species[sp], ds->ms_theta, C, (char*)ds->ms_lp + (size_t)sp * C * es,
                         (char*)ds->ms_grad + (size_t)sp * C * Dsp * es, st, 0, ds->ms_lp64 + (size_t)sp * C);
    if (rc) return rc;
  }
  EvalParams p;
  fill(ds, p);
  p.L = c0->L;
  p.theta = theta; p.logp = logp; p.logp64 = logp64; p.grad = grad; p.C = C; p.D = ds->D;
  const int n = C * (ds->D + 1);
  if (es == 4) ms_combine_kernel<float><<<(n + th - 1) / th, th, 0, st>>>(p, Sp, Dsp, ds->ms_lp64, (const float*)ds->ms_grad);
  else ms_combine_kernel<double><<<(n + th - 1) / th, th, 0, st>>>(p, Sp, Dsp, ds->ms_lp64, (const double*)ds->ms_grad);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(BL_ERR_CUDA, "multi-species combine: %s", cudaGetErrorString(e));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return BL_OK;
}

}  // namespace bl
