// K5: one-time dataset packer.  Resolves the NaN mask exactly as the reference does once per trace
// (biolith/models/occu.py:136-142; modeling.py:15-17), applies nan_to_num, and rewrites the
// reference-layout arrays  y (1,S,P,J), X (S,Ks), W (S,P,J,Ko), T (S,P,J)  into the "SoA in tile"
// layout documented in common.cuh.  Runs once per fit; everything afterwards reads only the pack.
#include "common.cuh"

namespace bl {

template <typename TO> __device__ __forceinline__ TO type_max();
template <> __device__ __forceinline__ float type_max<float>() { return FLT_MAX; }
template <> __device__ __forceinline__ double type_max<double>() { return DBL_MAX; }

template <typename TO>
__device__ __forceinline__ TO nan_to_num(TO v) {  // jnp.nan_to_num defaults
  if (isnan(v)) return TO(0);
  if (isinf(v)) return v > TO(0) ? type_max<TO>() : -type_max<TO>();
  return v;
}

template <typename TO> __device__ __forceinline__ TO word_as(uint32_t w);
template <> __device__ __forceinline__ float word_as<float>(uint32_t w) { return __uint_as_float(w); }
template <> __device__ __forceinline__ double word_as<double>(uint32_t w) {
  return __longlong_as_double((long long)(unsigned long long)w);
}

template <typename TI, typename TO>
__global__ void pack_kernel(const TI* __restrict__ y, const TI* __restrict__ X, const TI* __restrict__ W,
                            const TI* __restrict__ Tdur, TO* __restrict__ out, Layout L, int model, int K,
                            int* __restrict__ err_flag, unsigned long long* __restrict__ n_masked) {
  const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= L.n_units) return;
  const int64_t s = u / L.P;
  const int ks = L.ks, ko = L.ko, J = L.J;
  TO* base = out + (u / kWarp) * (int64_t)L.F * kWarp + (u % kWarp);
  bool site_nan = false;
  for (int k = 0; k < ks; ++k) {
    const TO v = (TO)X[s * ks + k];
    site_nan |= isnan(v);
    base[k * kWarp] = nan_to_num<TO>(v);
  }
  uint32_t ybits = 0, mw = 0;
  int n1 = 0;
  TO sy = TO(0), st = TO(0);
  double cst = 0.0;
  TO ymax = TO(0), y0 = TO(0), yw[kMaxCov];
  for (int k = 0; k < kMaxCov; ++k) yw[k] = TO(0);
  unsigned long long masked = 0;
  for (int j = 0; j < J; ++j) {
    bool cov_nan = site_nan;
    for (int k = 0; k < ko; ++k) {
      const TO v = (TO)W[(u * J + j) * ko + k];
      cov_nan |= isnan(v);
      base[(L.off_w + j * ko + k) * kWarp] = nan_to_num<TO>(v);
    }
    const TO yv = (TO)y[u * J + j];
    const bool m = isfinite(yv) && !cov_nan;
    masked += m ? 0 : 1;
    if (m) mw |= 1u << (j & 31);
    if (model == BL_MODEL_NMIXTURE) {
      if (m && (yv < TO(0) || yv != floor(yv) || yv > (TO)K)) atomicOr(err_flag, 4);
      base[(L.off_y + j) * kWarp] = m ? yv : TO(0);
      if (m) {
        ymax = yv > ymax ? yv : ymax;
        y0 += yv;
        for (int k = 0; k < ko; ++k) yw[k] += yv * base[(L.off_w + j * ko + k) * kWarp];
      }
    } else if (model == BL_MODEL_OCCU_CS) {
      base[(L.off_y + j) * kWarp] = m ? yv : TO(0);  // any finite score is valid
    } else if (model == BL_MODEL_OCCU_COP) {
      const TO tv = Tdur ? (TO)Tdur[u * J + j] : TO(1);
      if (m && (yv < TO(0) || !isfinite(tv))) atomicOr(err_flag, 2);
      base[(L.off_y + j) * kWarp] = m ? yv : TO(0);
      base[(L.off_t + j) * kWarp] = m ? tv : TO(0);
      if (m) {
        sy += yv; st += tv;
        cst += (yv > TO(0) ? (double)yv * log((double)tv) : 0.0) - lgamma((double)yv + 1.0);
      }
    } else {
      if (m) {
        if (yv == TO(1)) { ybits |= 1u << (j & 31); ++n1; }
        else if (yv != TO(0)) atomicOr(err_flag, 1);  // detections must be binary
      }
    }
    if ((j & 31) == 31 || j == J - 1) {
      base[(L.off_m + (j >> 5)) * kWarp] = word_as<TO>(mw);
      if (model != BL_MODEL_OCCU_COP && model != BL_MODEL_NMIXTURE && model != BL_MODEL_OCCU_CS)
        base[(L.off_y + (j >> 5)) * kWarp] = word_as<TO>(ybits);
      ybits = 0; mw = 0;
    }
  }
  if (model == BL_MODEL_NMIXTURE) {
    base[L.off_sy * kWarp] = ymax;
    base[(L.off_sy + 1) * kWarp] = y0;
    for (int k = 0; k < ko; ++k) base[(L.off_sy + 2 + k) * kWarp] = yw[k];
  } else if (model == BL_MODEL_OCCU_COP) {
    base[L.off_sy * kWarp] = sy;
    base[(L.off_sy + 1) * kWarp] = st;
    base[(L.off_sy + 2) * kWarp] = (TO)cst;
  } else if (model != BL_MODEL_OCCU_CS) {
    base[L.off_n1 * kWarp] = (TO)n1;
  }
  if (masked) atomicAdd(n_masked, masked);
}

template <typename TO>
__global__ void export_mask_kernel(const TO* __restrict__ packed, uint8_t* __restrict__ mask, Layout L) {
  const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= L.n_units) return;
  const TO* base = packed + (u / kWarp) * (int64_t)L.F * kWarp + (u % kWarp);
  for (int j = 0; j < L.J; ++j) {
    const uint32_t mw = Num<TO>::as_bits(base[(L.off_m + (j >> 5)) * kWarp]);
    mask[u * L.J + j] = (mw >> (j & 31)) & 1u;
  }
}

cudaError_t launch_pack(int data_dtype, int dtype, const void* y, const void* X, const void* W, const void* T,
                        void* out, const Layout& L, int model, int K, int* err_flag, unsigned long long* n_masked,
                        cudaStream_t st) {
  const int threads = 256;
  const unsigned blocks = (unsigned)((L.n_units + threads - 1) / threads);
  if (blocks == 0) return cudaSuccess;
#define BL_PACK(TI, TO)                                                                                        \
  pack_kernel<TI, TO><<<blocks, threads, 0, st>>>((const TI*)y, (const TI*)X, (const TI*)W, (const TI*)T,       \
                                                  (TO*)out, L, model, K, err_flag, n_masked)
  if (data_dtype == BL_F32 && dtype == BL_F32) BL_PACK(float, float);
  else if (data_dtype == BL_F64 && dtype == BL_F32) BL_PACK(double, float);
  else if (data_dtype == BL_F32 && dtype == BL_F64) BL_PACK(float, double);
  else BL_PACK(double, double);
#undef BL_PACK
  return cudaGetLastError();
}

cudaError_t launch_export_mask(int dtype, const void* packed, uint8_t* mask, const Layout& L, cudaStream_t st) {
  const int threads = 256;
  const unsigned blocks = (unsigned)((L.n_units + threads - 1) / threads);
  if (blocks == 0) return cudaSuccess;
  if (dtype == BL_F32) export_mask_kernel<float><<<blocks, threads, 0, st>>>((const float*)packed, mask, L);
  else export_mask_kernel<double><<<blocks, threads, 0, st>>>((const double*)packed, mask, L);
  return cudaGetLastError();
}

}  // namespace bl
