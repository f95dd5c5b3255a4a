// K1s host side + its instantiations (kernel and launch templates: occu_small.cuh).
#include "occu_small.cuh"

namespace bl {

bool occu_small_supported(int dtype, int ks, int ko, uint32_t flags) {
  if (const char* e = getenv("BL_SMALL_KERNEL"))  // A/B switch, read when a plan is made
    if (atoi(e) == 0) return false;
  if (dtype != BL_F32) return false;
  // strict math below 32 chains stays on the libm engine: the kernel carries a STRICT parameter like K1d's, but its
  // 40 libm instantiations (8 chain counts x 5 shapes, visits and chains unrolled) cost ~20 CPU-minutes of nvcc
  if (flags & (BL_FLAG_FP_CONSTANT | BL_FLAG_FP_UNOCCUPIED | BL_FLAG_STRICT_MATH)) return false;
  return ks >= 0 && ks <= kSmallMaxKs && ko >= 1 && ko <= 4;
}

int occu_small_max_chains() { return kSmallMaxNC; }
size_t occu_small_smem(const Layout& L, int nstage, int block_threads) {
  return kSmallHeader + (size_t)nstage * (block_threads / kWarp) * L.F * kWarp * sizeof(float);
}

// threads per block: as many warps as the register file carries for CB chains, fewer when the units are wide (every
// warp owns a ring of >= 2 warp-tiles of F x 128 B); 0 = does not fit (the engine serves the shape)
int occu_small_block_threads(const Layout& L, int CB, size_t smem_budget) {
  int bt = small_bt(CB);
  while (bt >= 128 && occu_small_smem(L, 2, bt) > smem_budget) bt -= 32;
  return bt >= 128 ? bt : 0;
}

cudaError_t launch_occu_small(const EvalParams& p, dim3 grid, size_t smem, cudaStream_t st, int* occ) {
  const int ks = p.L.ks, ko = p.L.ko;
  const bool spec = ks == 5 && ko == 3 && p.L.J == 8;
  if (p.flags & BL_FLAG_STRICT_MATH) return cudaErrorNotSupported;
  if (spec) return launch_small_nc<5, 3, true, false>(p, grid, smem, st, occ);
  if (ko == 1) return launch_small_nc<-1, 1, false, false>(p, grid, smem, st, occ);
  if (ko == 2) return launch_small_nc<-1, 2, false, false>(p, grid, smem, st, occ);
  if (ko == 3) return launch_small_nc<-1, 3, false, false>(p, grid, smem, st, occ);
  if (ko == 4) return launch_small_nc<-1, 4, false, false>(p, grid, smem, st, occ);
  return cudaErrorNotSupported;
}

}  // namespace bl

#ifdef BL_TRACE
extern "C" __attribute__((visibility("default"))) int bl_debug_small_trace(unsigned long long* out, int n_blocks) {
  return (int)cudaMemcpyFromSymbol(out, bl::g_small_trace, (size_t)n_blocks * 8 * sizeof(unsigned long long));
}
#endif
