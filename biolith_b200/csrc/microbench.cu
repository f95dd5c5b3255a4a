// Pipe micro-benchmarks: the MEASURED denominators of the compute rooflines bench.py reports for the
// lane = chain kernels (they are bound by the SFU pipe and by instruction issue, not by HBM).
//   which = 0  MUFU.EX2 throughput        (lane-ops / s): 8 independent ex2 chains per thread
//   which = 1  warp-instruction issue rate (warp-instructions / s): independent FFMA chains, 4 schedulers / SM
// Timed with CUDA events around back-to-back launches on the given device; no data is read from HBM.
#include <cuda_runtime.h>

#include "handle.h"

namespace bl {

constexpr int kMbThreads = 256;
constexpr int kMbIters = 4096;
constexpr int kMbChains = 8;

__global__ void __launch_bounds__(kMbThreads) mufu_bench_kernel(float* out, float seed) {
  float v[kMbChains];
#pragma unroll
  for (int i = 0; i < kMbChains; ++i) v[i] = seed + 1e-3f * (float)(threadIdx.x + i);
  for (int it = 0; it < kMbIters; ++it) {
#pragma unroll
    for (int i = 0; i < kMbChains; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kMbChains; ++i) s += v[i];
  if (s == 12345.678f) out[0] = s;  // never true; keeps the chains alive
}

__global__ void __launch_bounds__(kMbThreads) issue_bench_kernel(float* out, float a, float b) {
  float v[kMbChains];
#pragma unroll
  for (int i = 0; i < kMbChains; ++i) v[i] = a * (float)(threadIdx.x + i);
  for (int it = 0; it < kMbIters; ++it) {
#pragma unroll
    for (int i = 0; i < kMbChains; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(v[i]) : "f"(a), "f"(b));
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kMbChains; ++i) s += v[i];
  if (s == 12345.678f) out[0] = s;
}

// Error statistics of the MUFU approximations the fp32 kernels use, against double-precision references:
// which = 0: ex2.approx(x), relative error; 1: lg2.approx(x), absolute error; 2: rcp.approx + one Newton step,
// relative error.  x uniform on [lo, hi]; out = {sum err, sum err^2, max |err|} (double atomics).
__global__ void mufu_error_kernel(int which, float lo, float hi, long long n, double* out) {
  double s1 = 0.0, s2 = 0.0, mx = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float x = lo + (hi - lo) * (float)((double)(i + 0.5) / (double)n);
    float y;
    double ref, err;
    if (which == 0) {
      asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
      ref = exp2((double)x);
      err = ((double)y - ref) / ref;
    } else if (which == 1) {
      asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
      ref = log2((double)x);
      err = (double)y - ref;
    } else {
      asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
      y = fmaf(y, fmaf(-x, y, 1.0f), y);
      ref = 1.0 / (double)x;
      err = ((double)y - ref) / ref;
    }
    s1 += err; s2 += err * err; mx = fmax(mx, fabs(err));
  }
  atomicAdd(&out[0], s1);
  atomicAdd(&out[1], s2);
  // max via CAS on the bit pattern (non-negative doubles order like integers)
  unsigned long long* m = reinterpret_cast<unsigned long long*>(&out[2]);
  unsigned long long v = (unsigned long long)__double_as_longlong(mx), old = *m;
  while (v > old) {
    const unsigned long long prev = atomicCAS(m, old, v);
    if (prev == old) break;
    old = prev;
  }
}

}  // namespace bl

#define CU_TRY(expr)                                                                              \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess) return bl::fail(BL_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(_e)); \
  } while (0)

extern "C" {

BL_API int bl_pipe_peak(int32_t device, int32_t which, double* per_second, double* sm_clock_mhz_hint) {
  using namespace bl;
  if (!per_second || which < 0 || which > 1) return fail(BL_ERR_INVALID, "bad argument");
  CU_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  CU_TRY(cudaGetDeviceProperties(&prop, device));
  float* d_out = nullptr;
  CU_TRY(cudaMalloc(&d_out, 16));
  cudaEvent_t e0, e1;
  CU_TRY(cudaEventCreate(&e0));
  CU_TRY(cudaEventCreate(&e1));
  const int blocks = prop.multiProcessorCount * 8;  // 8 x 256 threads resident per SM: all 4 schedulers saturated
  const int reps = 8;
  double best = 0.0;
  for (int trial = 0; trial < 4; ++trial) {  // first trial warms the clocks up
    CU_TRY(cudaEventRecord(e0));
    for (int r = 0; r < reps; ++r) {
      if (which == 0) mufu_bench_kernel<<<blocks, kMbThreads>>>(d_out, 0.25f);
      else issue_bench_kernel<<<blocks, kMbThreads>>>(d_out, 0.999f, 1e-3f);
      g_launches.fetch_add(1, std::memory_order_relaxed);
    }
    CU_TRY(cudaEventRecord(e1));
    CU_TRY(cudaEventSynchronize(e1));
    float ms = 0.f;
    CU_TRY(cudaEventElapsedTime(&ms, e0, e1));
    const double thread_ops = (double)blocks * kMbThreads * (double)kMbIters * kMbChains * reps;
    const double rate = (which == 0 ? thread_ops : thread_ops / 32.0) / (ms * 1e-3);
    if (trial > 0 && rate > best) best = rate;
  }
  *per_second = best;
  if (sm_clock_mhz_hint) *sm_clock_mhz_hint = prop.clockRate / 1e3;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d_out);
  return BL_OK;
}

BL_API int bl_mufu_error(int32_t device, int32_t which, float lo, float hi, int64_t n, double* mean, double* rms,
                         double* max_abs) {
  using namespace bl;
  if (which < 0 || which > 2 || n < 1 || !mean || !rms || !max_abs) return fail(BL_ERR_INVALID, "bad argument");
  CU_TRY(cudaSetDevice(device));
  double* d_out = nullptr;
  CU_TRY(cudaMalloc(&d_out, 3 * sizeof(double)));
  CU_TRY(cudaMemset(d_out, 0, 3 * sizeof(double)));
  mufu_error_kernel<<<592, 256>>>(which, lo, hi, (long long)n, d_out);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  double h[3];
  CU_TRY(cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost));
  cudaFree(d_out);
  *mean = h[0] / (double)n;
  *rms = sqrt(h[1] / (double)n);
  *max_abs = h[2];
  return BL_OK;
}

}  // extern "C"
