#include "engine.cuh"
namespace bl {
cudaError_t launch_occu_cop(const EvalParams&, int, dim3, size_t, cudaStream_t, int*) { return cudaErrorNotSupported; }
int occu_cop_derived_slots(uint32_t) { return 0; }
}  // namespace bl
