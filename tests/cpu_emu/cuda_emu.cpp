// cuda_emu.cpp -- TEST INFRASTRUCTURE ONLY: scheduler of the CUDA-on-CPU shim
// (see cuda_emu.h).  Coroutine per CUDA thread, round-robin, one CTA at a time.
#include "cuda_emu.h"

emu_idx threadIdx, blockIdx;
emu_dim3 blockDim, gridDim;

namespace emu {
Cta *g_cta = nullptr;
static LaunchArgs g_launch;

char *dyn_smem() { return g_cta->smem.data(); }

void run_thread_trampoline() {
    g_launch.fn(g_launch.args);
    g_cta->th[g_cta->cur].done = true;
    swapcontext(&g_cta->th[g_cta->cur].ctx, &g_cta->sched);
}

void launch(void (*fn)(void *), void *args, unsigned grid, unsigned threads, size_t smem_bytes) {
    g_launch = {fn, args};
    gridDim.x = grid; blockDim.x = threads;
    for (unsigned b = 0; b < grid; ++b) {
        Cta cta;
        cta.nthreads = (int)threads;
        cta.th.resize(threads);
        cta.w_count.assign((threads + 31) / 32, 0);
        cta.w_gen.assign((threads + 31) / 32, 0);
        cta.slot.assign(threads, 0);
        cta.smem.assign(smem_bytes + 64, 0x7f);        // poison: uninitialised reads show up
        g_cta = &cta;
        blockIdx.x = b;
        for (unsigned t = 0; t < threads; ++t) {
            Thread &th = cta.th[t];
            th.stack.resize(threads > 256 ? 64 * 1024 : 256 * 1024);
            getcontext(&th.ctx);
            th.ctx.uc_stack.ss_sp = th.stack.data();
            th.ctx.uc_stack.ss_size = th.stack.size();
            th.ctx.uc_link = &cta.sched;
            makecontext(&th.ctx, (void (*)())run_thread_trampoline, 0);
        }
        int remaining = (int)threads;
        long spins = 0;
        while (remaining > 0) {
            int progressed = 0;
            for (unsigned t = 0; t < threads; ++t) {
                Thread &th = cta.th[t];
                if (th.done) continue;
                cta.cur = (int)t;
                threadIdx.x = t;
                swapcontext(&cta.sched, &th.ctx);
                if (th.done) { --remaining; }
                ++progressed;
            }
            if (++spins > 50000000L) { std::fprintf(stderr, "emu: deadlock (barrier mismatch?)\n"); std::abort(); }
            (void)progressed;
        }
        g_cta = nullptr;
    }
}
}  // namespace emu
