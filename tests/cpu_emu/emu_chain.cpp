// emu_chain.cpp -- TEST INFRASTRUCTURE ONLY.  Runs the product's fused chain kernel
// source (csrc/aes_chain_kernel.cuh) and plan compiler (csrc/aes_plan_build.h) on
// the CPU through the cuda_emu.h shim, so kernel logic can be checked against the
// oracle on a machine without a GPU.  Never linked into libaesim.so.
#include "cuda_emu.h"
#include "../../audio-effects-simulator_b200/csrc/aes_plan_build.h"
#include "../../audio-effects-simulator_b200/csrc/aes_chain_kernel.cuh"

static char g_err[512];

template <int K> static void entry(void *p) { aes_chain_body<K>(*reinterpret_cast<ChainArgs *>(p)); }

extern "C" __attribute__((visibility("default")))
const char *emu_last_error() { return g_err; }

extern "C" __attribute__((visibility("default")))
int emu_chain_run(const aes_stage_desc *stages, int n, int fs, const void *x, int in_fmt, void *y,
                  int out_fmt, long long B, long long N, int grid, double *state_out)
{
    static DevPlan plan;
    int rc = aes_build_devplan(stages, n, fs, &plan, g_err, sizeof g_err);
    if (rc) return rc;
    if (grid < 1) grid = 1;
    std::vector<float> scratch((size_t)grid * plan.scratch_floats, 1e30f);   // poison
    ChainArgs a{ &plan, x, y, B, N, scratch.data(), in_fmt, out_fmt, state_out };
    size_t smem = aes_plan_smem_bytes(plan);
    switch (plan.K) {
    case 8: emu::launch(entry<8>, &a, grid, AES_NT, smem); break;
    case 4: emu::launch(entry<4>, &a, grid, AES_NT, smem); break;
    case 2: emu::launch(entry<2>, &a, grid, AES_NT, smem); break;
    default: snprintf(g_err, sizeof g_err, "bad K"); return -1;
    }
    return 0;
}

extern "C" __attribute__((visibility("default")))
int emu_plan_info(const aes_stage_desc *stages, int n, int fs, int *T, int *smem_bytes, long long *scratch_floats)
{
    static DevPlan plan;
    int rc = aes_build_devplan(stages, n, fs, &plan, g_err, sizeof g_err);
    if (rc) return rc;
    *T = plan.T; *smem_bytes = (int)aes_plan_smem_bytes(plan); *scratch_floats = plan.scratch_floats;
    return 0;
}
