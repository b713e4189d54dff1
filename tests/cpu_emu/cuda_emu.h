// cuda_emu.h -- TEST INFRASTRUCTURE ONLY (never part of the product library).
//
// A minimal CUDA-on-CPU execution shim: it lets the *same kernel source* that
// nvcc compiles for sm_100a (audio-effects-simulator_b200/csrc/*.cuh) be compiled
// by g++ and executed on a machine with no GPU, so the index arithmetic, scans,
// barriers and shuffles can be debugged against the oracle before a GPU box is
// spent on them.  One CTA at a time; every CUDA thread is a ucontext coroutine;
// __syncthreads / __syncwarp / __shfl_*_sync are real rendezvous points, so a
// missing barrier shows up as a wrong result here too (scheduling is round-robin,
// so it is not a race detector).  The product never includes this header: it is
// guarded by AES_CPU_EMU, which only tests/cpu_emu/Makefile defines.
#pragma once
#ifndef AES_CPU_EMU
#error "cuda_emu.h is test infrastructure; the product is built with nvcc"
#endif

#include <ucontext.h>
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __grid_constant__

struct emu_dim3 { unsigned x = 1, y = 1, z = 1; };
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct short2 { short x, y; };
struct double2 { double x, y; };
static inline float2 make_float2(float a, float b) { return {a, b}; }
static inline float4 make_float4(float a, float b, float c, float d) { return {a, b, c, d}; }
static inline short2 make_short2(short a, short b) { return {a, b}; }

namespace emu {
struct Thread {
    ucontext_t ctx;
    std::vector<char> stack;
    bool done = false;
};
struct Cta {
    std::vector<Thread> th;
    ucontext_t sched;
    int cur = 0;
    int nthreads = 0;
    // block barrier
    int bar_count = 0; unsigned bar_gen = 0;
    // per-warp rendezvous + shuffle slots
    std::vector<int> w_count; std::vector<unsigned> w_gen;
    std::vector<uint64_t> slot;       // nthreads x 4 values
    std::vector<char> smem;
    // named barriers (bar.sync / bar.arrive id, count): arrivals so far and completed generations
    int nb_count[16] = {}; unsigned nb_gen[16] = {};
};
extern Cta *g_cta;
extern thread_local int dummy;
inline void yield() { swapcontext(&g_cta->th[g_cta->cur].ctx, &g_cta->sched); }
inline void block_barrier() {
    Cta &c = *g_cta; unsigned gen = c.bar_gen;
    if (++c.bar_count == c.nthreads) { c.bar_count = 0; ++c.bar_gen; return; }
    while (c.bar_gen == gen) yield();
}
// bar.sync id, n: wait until n threads have arrived (by sync or arrive); bar.arrive id, n: count, do not wait
inline void named_sync(int id, int n) {
    Cta &c = *g_cta; unsigned gen = c.nb_gen[id];
    if (++c.nb_count[id] == n) { c.nb_count[id] = 0; ++c.nb_gen[id]; return; }
    if (c.nb_count[id] > n) { std::fprintf(stderr, "emu: named barrier %d over-subscribed\n", id); std::abort(); }
    while (c.nb_gen[id] == gen) yield();
}
inline void named_arrive(int id, int n) {
    Cta &c = *g_cta;
    if (++c.nb_count[id] == n) { c.nb_count[id] = 0; ++c.nb_gen[id]; }
    else if (c.nb_count[id] > n) { std::fprintf(stderr, "emu: named barrier %d over-subscribed\n", id); std::abort(); }
}
inline void warp_barrier() {
    Cta &c = *g_cta; int w = c.cur >> 5; unsigned gen = c.w_gen[w];
    int lanes = std::min(32, c.nthreads - (w << 5));
    if (++c.w_count[w] == lanes) { c.w_count[w] = 0; ++c.w_gen[w]; return; }
    while (c.w_gen[w] == gen) yield();
}
}  // namespace emu

// thread / block indices (per running coroutine)
struct emu_idx { unsigned x, y, z; };
extern emu_idx threadIdx, blockIdx;
extern emu_dim3 blockDim, gridDim;

static inline void __syncthreads() { emu::block_barrier(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu::warp_barrier(); }

template <typename T> static inline T emu_shfl(T v, int src_lane_rel, bool up, int delta) {
    static_assert(sizeof(T) <= 8, "shuffle payload");
    emu::Cta &c = *emu::g_cta;
    int me = c.cur, lane = me & 31, base = me & ~31;
    uint64_t raw = 0; std::memcpy(&raw, &v, sizeof(T));
    c.slot[me] = raw;
    emu::warp_barrier();
    int src = up ? lane - delta : src_lane_rel;
    T out = v;
    int lanes = std::min(32, c.nthreads - base);
    if (src >= 0 && src < lanes) { uint64_t r = c.slot[base + src]; std::memcpy(&out, &r, sizeof(T)); }
    emu::warp_barrier();
    return out;
}
template <typename T> static inline T __shfl_up_sync(unsigned, T v, int d) { return emu_shfl(v, 0, true, d); }
template <typename T> static inline T __shfl_sync(unsigned, T v, int lane) { return emu_shfl(v, lane & 31, false, 0); }
template <typename T> static inline T __shfl_down_sync(unsigned, T v, int d) {
    return emu_shfl(v, (int)(emu::g_cta->cur & 31) + d, false, 0);
}

template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int k) {
    return emu_shfl(v, (int)(emu::g_cta->cur & 31) ^ k, false, 0);
}
static inline unsigned __ballot_sync(unsigned, int pred) {
    emu::Cta &c = *emu::g_cta;
    int me = c.cur, base = me & ~31;
    c.slot[me] = pred ? 1 : 0;
    emu::warp_barrier();
    unsigned m = 0;
    int lanes = std::min(32, c.nthreads - base);
    for (int l = 0; l < lanes; ++l) if (c.slot[base + l]) m |= 1u << l;
    emu::warp_barrier();
    return m;
}
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __clz(int x) { return x == 0 ? 32 : __builtin_clz((unsigned)x); }
template <typename T> static inline T __ldg(const T *p) { return *p; }
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
static inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
static inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
static inline float sinpif(float x) { return (float)std::sin(M_PI * (double)x); }
static inline double sinpi(double x) { return std::sin(M_PI * x); }
static inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
static inline int __float_as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }
static inline float __saturatef(float x) { return x < 0.f ? 0.f : (x > 1.f ? 1.f : x); }
using std::fma; using std::floor; using std::fabs;
using std::min; using std::max;

namespace emu {
struct LaunchArgs { void (*fn)(void *); void *args; };
void run_thread_trampoline();
// Launch `fn(args)` as a grid of CTAs of `threads` threads with `smem_bytes` of
// dynamic shared memory; CTAs run one after another.
void launch(void (*fn)(void *), void *args, unsigned grid, unsigned threads, size_t smem_bytes);
char *dyn_smem();
}  // namespace emu

#define AES_DYN_SMEM(type, name) type *name = reinterpret_cast<type *>(emu::dyn_smem())
static inline int __float2int_rz(float f) { return (int)f; }
template <typename T> static inline T __ldcs(const T *p) { return *p; }
template <typename T> static inline void __stcs(T *p, T v) { *p = v; }
