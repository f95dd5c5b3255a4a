#include "engine.cuh"
namespace bl {
cudaError_t launch_occu_rn(const EvalParams&, int, dim3, size_t, cudaStream_t, int*) { return cudaErrorNotSupported; }
int occu_rn_derived_slots(uint32_t) { return 0; }
size_t occu_rn_extra_smem(const Layout&, int, int) { return 0; }
}  // namespace bl
