// Host-only check of the launch-geometry rules (csrc/engine.cuh:plan_geometry, csrc/common.cuh:make_layout):
// for a sweep of shapes / batch sizes / dtypes the plan must fit the shared-memory budget it was given,
// keep 8 warps per block, at least one ring stage, and cover every chain.  Runs on the CPU (no kernel
// launch); compiled and executed by tests/test_geometry_cpu.py.
#include <cstdio>

#include "../../biolith_b200/csrc/engine.cuh"

using namespace bl;

int main() {
  const size_t smem_limit = 227 * 1024 - 1024;  // sharedMemPerBlockOptin - 1 KB, as api.cu passes it
  const int models[] = {BL_MODEL_OCCU, BL_MODEL_OCCU_RN, BL_MODEL_OCCU_COP, BL_MODEL_NMIXTURE, BL_MODEL_OCCU_CS};
  const int Js[] = {1, 8, 12, 33, 52, 120};
  const int Ks[][2] = {{0, 0}, {1, 1}, {5, 3}, {16, 16}};
  const int Cs[] = {1, 2, 3, 5, 8, 16, 31, 64, 127, 256, 700, 1024, 5000};
  long checked = 0, failed = 0, unsupported = 0;
  for (int model : models)
    for (int J : Js)
      for (auto& k : Ks)
        for (int elem : {4, 8})
          for (int C : Cs)
            for (int bps : {1, 2}) {
              const Layout L = make_layout(model, 100000, 1, J, k[0], k[1]);
              const int extras = model == BL_MODEL_OCCU_CS ? 4 : model == BL_MODEL_OCCU_COP ? 2 : 1;
              const int D = k[0] + k[1] + 2 + extras, DS = D + 8;
              const Geometry g = plan_geometry(L, elem, C, D, DS, 148, bps, smem_limit);
              ++checked;
              bool ok = g.WC * g.WS == kWarpsPerBlock && g.nstage >= 1 && g.nstage <= kMaxStages &&
                        g.n_chunks >= 1 && (long)g.n_chunks * g.CB >= C && g.CB <= kMaxChainsPerBlock &&
                        g.nsplit >= 1 && g.nsplit <= g.n_block_tiles &&
                        g.n_block_tiles * g.WS >= L.n_tiles &&
                        g.smem_bytes == eval_smem_bytes(L, elem, g.WS, g.nstage, g.CB, D, DS);
              if (g.smem_bytes > smem_limit) {
                // a shape may be too wide even for one warp-tile and one stage: api.cu reports
                // BL_ERR_UNSUPPORTED; the rule must then already be at its smallest configuration
                ok = ok && g.WS == 1 && g.nstage == 1;
                ++unsupported;
              } else if (g.nstage >= 2) {
                ok = ok && g.smem_bytes <= smem_limit / bps + 0;  // the ring was sized for bps resident blocks
              }
              if (!ok) {
                ++failed;
                std::printf("FAIL model=%d J=%d ks=%d ko=%d elem=%d C=%d bps=%d -> WC=%d WS=%d nstage=%d CB=%d chunks=%d "
                            "nsplit=%d smem=%zu\n", model, J, k[0], k[1], elem, C, bps, g.WC, g.WS, g.nstage, g.CB,
                            g.n_chunks, g.nsplit, g.smem_bytes);
              }
            }
  std::printf("checked=%ld failed=%ld too_wide=%ld\n", checked, failed, unsupported);
  return failed ? 1 : 0;
}
