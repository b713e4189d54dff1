// aes_stream.cu -- host side of the block-streaming path (aes_stream_kernel.cuh).
#include <algorithm>
#include <mutex>
#include <string.h>

#include "aes_common.h"
#include "aes_chain_kernel.cuh"      // AES_DYN_SMEM
#include "aes_stream_kernel.cuh"

__global__ void __launch_bounds__(AESS_NT) aes_stream_kernel(const __grid_constant__ StreamArgs a) { aes_stream_body(a); }

namespace {
struct StreamCtx {
    std::mutex mu;
    int device = -1;
    float *d_x = nullptr, *d_y = nullptr;
    aes_stage_desc *d_stages = nullptr;
    double *d_state = nullptr;
    bool attr_set = false;
} g_ctx;
}

// One call of a run of blocks' process_into with carried state (any block size; internally cut
// into pieces of <= 4096 frames).  `stages` is updated in place so the caller can keep streaming:
// q[29] (frames since prepare) advances, and the carried scalars are replaced by their final values
// (BIQUAD p[8..15], GATE p[3], OCTAVER p[0] and q[1]).  Ring contents live in the device blobs the
// descriptors point to (q[28]).  x_host: (frames, channels_in) f32, y_host: (frames, 2) f32.
AES_EXPORT int aes_stream_process_host(aes_stage_desc *stages, int n_stages, const float *x_host, int channels_in,
                                       float *y_host, int64_t frames)
{
    AES_REQUIRE(n_stages >= 0 && n_stages <= AES_MAX_STAGES, "0..16 stages");
    AES_REQUIRE(channels_in == 1 || channels_in == 2, "mono or stereo input");
    if (frames <= 0) return 0;
    AES_REQUIRE(x_host != nullptr && y_host != nullptr && (n_stages == 0 || stages != nullptr), "NULL argument");
    for (int s = 0; s < n_stages; ++s) {
        const aes_stage_desc &d = stages[s];
        const bool needs_blob = d.kind == AES_STAGE_DELAY || d.kind == AES_STAGE_REVERB || d.kind == AES_STAGE_OCTAVER;
        AES_REQUIRE(!needs_blob || d.q[28] != 0, "stage %d: no state blob", s);
        AES_REQUIRE(d.kind != AES_STAGE_REVERB || (d.q[0] <= AES_MAX_COMB && d.q[1] <= AES_MAX_AP), "reverb: too many lines");
    }
    std::lock_guard<std::mutex> lock(g_ctx.mu);
    int dev = 0;
    AES_CUDA(cudaGetDevice(&dev));
    if (g_ctx.device != dev) {
        if (g_ctx.d_x) { cudaFree(g_ctx.d_x); cudaFree(g_ctx.d_y); cudaFree(g_ctx.d_stages); cudaFree(g_ctx.d_state); }
        g_ctx.d_x = g_ctx.d_y = nullptr; g_ctx.d_stages = nullptr; g_ctx.d_state = nullptr;
        AES_CUDA(cudaMalloc(&g_ctx.d_x, AESS_MAX_FRAMES * 2 * sizeof(float)));
        AES_CUDA(cudaMalloc(&g_ctx.d_y, AESS_MAX_FRAMES * 2 * sizeof(float)));
        AES_CUDA(cudaMalloc(&g_ctx.d_stages, AES_MAX_STAGES * sizeof(aes_stage_desc)));
        AES_CUDA(cudaMalloc(&g_ctx.d_state, AES_MAX_STAGES * 16 * sizeof(double)));
        AES_CUDA(cudaFuncSetAttribute(aes_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)(AESS_SMEM_FLOATS(AESS_MAX_FRAMES) * sizeof(float))));
        g_ctx.device = dev;
    }
    double fin[AES_MAX_STAGES * 16];
    for (int64_t f0 = 0; f0 < frames; f0 += AESS_MAX_FRAMES) {
        const int nf = (int)std::min<int64_t>(AESS_MAX_FRAMES, frames - f0);
        AES_CUDA(cudaMemcpy(g_ctx.d_x, x_host + f0 * channels_in, (size_t)nf * channels_in * sizeof(float), cudaMemcpyHostToDevice));
        if (n_stages)
            AES_CUDA(cudaMemcpy(g_ctx.d_stages, stages, (size_t)n_stages * sizeof(aes_stage_desc), cudaMemcpyHostToDevice));
        StreamArgs a;
        a.stages = g_ctx.d_stages; a.n_stages = n_stages; a.ci = channels_in;
        a.x = g_ctx.d_x; a.y = g_ctx.d_y; a.frames = nf; a.state_out = g_ctx.d_state;
        aes_stream_kernel<<<1, AESS_NT, AESS_SMEM_FLOATS(nf) * sizeof(float)>>>(a);
        aes_count_launch();
        AES_CUDA(cudaGetLastError());
        AES_CUDA(cudaMemcpy(y_host + f0 * 2, g_ctx.d_y, (size_t)nf * 2 * sizeof(float), cudaMemcpyDeviceToHost));
        if (n_stages)
            AES_CUDA(cudaMemcpy(fin, g_ctx.d_state, (size_t)n_stages * 16 * sizeof(double), cudaMemcpyDeviceToHost));
        for (int s = 0; s < n_stages; ++s) {
            aes_stage_desc &d = stages[s];
            d.q[29] += nf;
            if (d.kind == AES_STAGE_BIQUAD) for (int i = 0; i < 8; ++i) d.p[8 + i] = fin[16 * s + i];
            else if (d.kind == AES_STAGE_GATE) d.p[3] = fin[16 * s];
            else if (d.kind == AES_STAGE_OCTAVER) { d.p[0] = fin[16 * s]; d.q[1] = (d.q[1] + nf) % d.q[0]; }
        }
    }
    return 0;
}
