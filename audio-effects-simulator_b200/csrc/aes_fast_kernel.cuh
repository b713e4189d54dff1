// aes_fast_kernel.cuh -- shape-specialised variant of the fused chain kernel.
//
// Same algorithm and thread mapping as aes_chain_kernel.cuh (thread = FR consecutive
// frames x 2 channels in registers, one CTA per clip, tiles of T = 256*FR frames), but
// the chain SHAPE (stage kinds, comb/all-pass counts, line modes) is a template
// parameter and every descriptor lives in the kernel-parameter constant bank:
//   * no plan loads from global memory (the generic interpreter spends ~100 LDG per
//     thread per tile on them), stage loop and comb loops fully unrolled;
//   * ring slots of the register-path lines are per-thread registers advanced by one
//     add + wrap per tile;  for a comb whose ring period is roundup4(L) the aligned read
//     base EQUALS the write slot (r = w - L = w + m  (mod len), m = len - L), so one
//     register per ring serves both;
//   * the comb one-pole is run on u = lp/(1-h):  u[n] = y[n] + h*u[n-1],
//     buf[n] = x[n] + (g*(1-h))*u[n]  -- one FFMA per sample per pass instead of FMUL+FFMA.
// The host (aes_chain.cu) picks an instantiation whose shape matches the plan, else runs
// the generic interpreter.  Also compiled for the CPU emulator (AES_CPU_EMU).
#pragma once
#include "aes_chain_kernel.cuh"

#define AESF_MAX_STAGES 4
#define AESF_MAX_WALK 12

// shape code of one stage (template parameter)
#define AESF_CODE(kind, nc, na, mode, premode, pf) \
    ((kind) | ((nc) << 4) | ((na) << 8) | ((mode) << 12) | ((premode) << 14) | ((pf) << 16))
#define AESF_KIND(c) ((c) & 15)
#define AESF_NC(c) (((c) >> 4) & 15)
#define AESF_NA(c) (((c) >> 8) & 15)
#define AESF_MODE(c) (((c) >> 12) & 3)
#define AESF_PREMODE(c) (((c) >> 14) & 3)      // 0 none, 1 REG (lag >= T), 3 REGB (lag < T: write, barrier, read)
#define AESF_PF(c) (((c) >> 16) & 1)

struct FRing {
    int off;        // float offset in the smem ring area or the CTA's global scratch
    int len;        // period (floats)
    int tinc;       // T mod len
    int lag;
};

struct FastStage {
    FRing ring[2][4];       // delay: ring[c][0]; reverb: comb rings
    FRing pre[2];           // reverb pre-delay line (REG); WALK lines use `walk` ids below
    int walk_pre[2];        // ids into FastArgs::walk (WALK pre-delay / WALK delay)
    int walk_ap[2][2];      // ids into FastArgs::walk
    int glob;               // delay / pre-delay REG rings live in the global scratch
    int nscan, nxw, pad0;
    float gs[2][4];         // g * (1 - h)
    float dry, wet, h, a, fb, mix, drive, pad1;
    float hp[6];            // h^(FR*2^s), hp[5] = h^(32*FR)
    int oct_size, oct_mask;
    double bq[5];
    double bq_pow[6][4];
    double bq_row[4][2];    // first row of (A^T)^j, j = 0..3: what an incoming TDF-II state adds to output j
    double thr, att, rel;
    double ph0, step, fsize;
};

struct FastArgs {
    FastStage st[AESF_MAX_STAGES];
    FRing walk[AESF_MAX_WALK];
    double init[AESF_MAX_STAGES][8];    // carried scalars at clip start (biquad 8, gate 1)
    const void *x;
    void *y;
    long long B, N;
    float *scratch;
    const float *lane_tab;              // global: per stage 32 x {hlane, pad, pad, pad, bq_lane[4] as 8 floats} -> see FAST_LANE_STRIDE
    double *state_out;
    long long scratch_floats;
    int in_fmt, out_fmt;
    int smem_floats, n_walk, n_stages, pad;
};
#define FAST_LANE_STRIDE 12             // floats per lane per stage: [0]=hlane, [4..11]=bq_lane as 4 doubles


// ---- cp.async.bulk (TMA, 1-D) staging of the next tile ------------------------------------
// One elected thread arms an mbarrier with the byte count and issues bulk copies
// global -> shared for the NEXT tile's input frames (8 KB) and, for a prefetched feedback
// delay, the T+4 line samples per channel it will read.  Nothing is held in registers and no
// scoreboard slot is occupied while the copies fly; consumers wait on the mbarrier parity
// at the top of the next tile and read their own 32 bytes with LDS.128.
#ifndef AES_CPU_EMU
__device__ __forceinline__ void aes_mbar_init(unsigned long long *bar, unsigned count)
{
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count) : "memory");
}
__device__ __forceinline__ void aes_mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void aes_mbar_expect(unsigned long long *bar, unsigned bytes)
{
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
}
// `policy`: an L2 cache policy from createpolicy (streamed input = evict_first, delay line = evict_last)
__device__ __forceinline__ void aes_bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar,
                                             unsigned long long policy)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst), b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(d), "l"(src), "r"(bytes), "r"(b), "l"(policy) : "memory");
}
__device__ __forceinline__ unsigned long long aes_policy_evict_first()
{
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ unsigned long long aes_policy_evict_last()
{
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void aes_mbar_wait(unsigned long long *bar, unsigned use_index)
{
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar), parity = use_index & 1u;
    unsigned done;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(a), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void aes_mbar_wait_parity(unsigned long long *bar, unsigned parity)
{
    aes_mbar_wait(bar, parity);                     // (only the low bit of the use index is looked at)
}
// generic-proxy global stores -> visible to later async-proxy (TMA) reads.  Executed by the
// issuing thread only, after the CTA barrier that orders every thread's ring stores before it
// (fences are cumulative); a per-thread fence after the stores cost 15 % of all stall samples.
__device__ __forceinline__ void aes_fence_proxy_async() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
// ... and the same for every state space (shared-memory buffers that the TMA overwrites after ordinary accesses)
__device__ __forceinline__ void aes_fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
// ... for this CTA's shared memory only (buffers that the TMA overwrites after ordinary loads / stores); the
// all-spaces form costs a MEMBAR.ALL.GPU
__device__ __forceinline__ void aes_fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
#else
// emulator: the copy completes at issue; the barrier word counts completed uses
static inline void aes_mbar_init(unsigned long long *bar, unsigned) { *bar = 0; }
static inline void aes_mbar_init_fence() {}
static inline void aes_mbar_expect(unsigned long long *, unsigned) {}
static inline void aes_bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *, unsigned long long) { std::memcpy(dst, src, bytes); }
static inline unsigned long long aes_policy_evict_first() { return 0; }
static inline unsigned long long aes_policy_evict_last() { return 0; }
static inline void aes_mbar_complete_emu(unsigned long long *bar) { *bar += 1; }
static inline void aes_mbar_wait(unsigned long long *bar, unsigned use_index) { while (*bar <= use_index) emu::yield(); }
// the phase with parity `parity` has completed once the count of completed phases has the other parity
static inline void aes_mbar_wait_parity(unsigned long long *bar, unsigned parity) { while ((*bar & 1u) == (parity & 1u)) emu::yield(); }
static inline void aes_fence_proxy_async() {}
static inline void aes_fence_proxy_async_all() {}
static inline void aes_fence_proxy_async_smem() {}
#endif

// ---- packed f32x2 arithmetic (sm_100: FFMA2 / FADD2 / FMUL2, one issue slot for two lanes' worth) ------
// Each component rounds exactly like the scalar __fmaf_rn / __fadd_rn / __fmul_rn (no contraction).
#ifndef AES_CPU_EMU
__device__ __forceinline__ float2 aes_fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 aes_add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 aes_mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
#else
static inline float2 aes_fma2(float2 a, float2 b, float2 c) { return make_float2(std::fmaf(a.x, b.x, c.x), std::fmaf(a.y, b.y, c.y)); }
static inline float2 aes_add2(float2 a, float2 b) { return make_float2(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y)); }
static inline float2 aes_mul2(float2 a, float2 b) { return make_float2(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)); }
#endif
// clip(dry*x + wet*w) for two samples: products and sum separately rounded like numpy (see aes_mix_clip)
__device__ __forceinline__ void aes_mix_clip2(float dry, float2 x, float wet, float2 w, float &o0, float &o1)
{
    const float2 t = aes_add2(aes_mul2(make_float2(dry, dry), x), aes_mul2(make_float2(wet, wet), w));
    o0 = aes_clip1(t.x);
    o1 = aes_clip1(t.y);
}

// ---- named barriers and per-role register budgets (warp-specialised kernels) ---------------------
#ifndef AES_CPU_EMU
__device__ __forceinline__ void aes_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ void aes_bar_arrive(int id, int nthreads) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
template <int R> __device__ __forceinline__ void aes_setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(R)); }
template <int R> __device__ __forceinline__ void aes_setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(R)); }
#else
static inline void aes_bar_sync(int id, int nthreads) { emu::named_sync(id, nthreads); }
static inline void aes_bar_arrive(int id, int nthreads) { emu::named_arrive(id, nthreads); }
template <int R> static inline void aes_setmaxnreg_inc() {}
template <int R> static inline void aes_setmaxnreg_dec() {}
#endif

struct FCtx {
    float *tile, *rings, *gscr;
    double *wtot;
    int *rpos;              // current parity
    int n0;                 // first frame of the tile inside its clip (clips hold < 2^31 frames)
    int len, tid, lane, warp;
    const float *ln_stage;  // this tile's staged delay line: [2 ch][T + 8] floats (PF shapes), or null
};

struct SRegs {              // per-thread, per-stage persistent ring slots (floats)
    int w[2][4];            // comb rings: write slot == aligned read base, as a BYTE offset in the ring
    int dw[2], da[2];       // delay / pre-delay REG line: write slot, aligned read base
};

__device__ __forceinline__ int aesf_adv(int s, int inc, int len)
{
    // (s + inc) mod len for 0 <= s, inc < len: as unsigned, the wrapped candidate is the smaller one
    // exactly when it did not underflow (add + add-min instead of add + compare + select)
    const unsigned a = (unsigned)s + (unsigned)inc, b = a - (unsigned)len;
    return (int)(a < b ? a : b);
}

// FR consecutive elements starting m (< FR) past the aligned base a0; the next aligned
// vector wraps at len.  m is uniform over the CTA.
template <int FR, int SP>
__device__ __forceinline__ void aesf_read(const float *rb, int a0, int m, int len, float (&o)[FR])
{
    if (FR < 4) {                                   // m is given relative to a 4-aligned base
        a0 += m & ~(FR - 1);
        if (a0 >= len) a0 -= len;
        m &= FR - 1;
    }
    float A[FR];
    aes_ldv_sp<FR, SP>(rb + a0, A);
    if (m == 0 || FR == 1) {
#pragma unroll
        for (int j = 0; j < FR; ++j) o[j] = A[j];
        return;
    }
    int b0 = a0 + FR;
    if (b0 >= len) b0 -= len;
    float Bv[FR];
    aes_ldv_sp<FR, SP>(rb + b0, Bv);
    if (m == 1) {
#pragma unroll
        for (int j = 0; j < FR; ++j) o[j] = (j + 1 < FR) ? A[(j + 1) % FR] : Bv[(j + 1) % FR];
    } else if (m == 2) {
#pragma unroll
        for (int j = 0; j < FR; ++j) o[j] = (j + 2 < FR) ? A[(j + 2) % FR] : Bv[(j + 2) % FR];
    } else {
#pragma unroll
        for (int j = 0; j < FR; ++j) o[j] = (j + 3 < FR) ? A[(j + 3) % FR] : Bv[(j + 3) % FR];
    }
}

// ---- compile-time reverb topologies ---------------------------------------------------------
// The reference's reverb has one default set of comb / all-pass times (reverb.py:74-81,158-177;
// SURVEY 3.1), so every reverb preset resolves to the same delay lengths at a given sample rate.
// TOPO > 0 bakes them into the kernel: ring periods, wrap compares, tile increments and the
// comb read misalignments become immediates, the misalignment switch folds away, and the
// all-pass walks unroll completely.  TOPO == 0 reads everything from the descriptors.
//   TOPO 1: 48 000 Hz    TOPO 2: 44 100 Hz
#define AESF_TOPO_NONE 0
#define AESF_TOPO_48K 1
#define AESF_TOPO_44K 2
__host__ __device__ constexpr int aesf_topo_comb(int topo, int ch, int cc)
{
    return topo == 1 ? (ch == 0 ? (cc == 0 ? 1440 : cc == 1 ? 1795 : cc == 2 ? 1987 : 2112)
                                : (cc == 0 ? 1411 : cc == 1 ? 1766 : cc == 2 ? 1958 : 2083))
         : topo == 2 ? (ch == 0 ? (cc == 0 ? 1323 : cc == 1 ? 1649 : cc == 2 ? 1825 : 1940)
                                : (cc == 0 ? 1296 : cc == 1 ? 1622 : cc == 2 ? 1799 : 1913))
                     : 0;
}
__host__ __device__ constexpr int aesf_topo_ap(int topo, int ch, int k)
{
    return topo == 1 ? (ch == 0 ? (k == 0 ? 242 : 84) : (k == 0 ? 237 : 78))
         : topo == 2 ? (ch == 0 ? (k == 0 ? 223 : 77) : (k == 0 ? 217 : 72))
                     : 0;
}
// float offsets in the shared-memory ring area: the plan builder lays the reverb's comb rings out
// first (side-major), then its all-pass rings, each padded to a multiple of 4 floats
__host__ __device__ constexpr int aesf_topo_comb_off(int topo, int ch, int cc)
{
    int off = 0;
    for (int i = 0; i < ch * 4 + cc; ++i) off += ((aesf_topo_comb(topo, i >> 2, i & 3) + 3) & ~3) + AES_COMB_RING_PAD;
    return off;
}
__host__ __device__ constexpr int aesf_topo_ap_off(int topo, int ch, int k)
{
    int off = aesf_topo_comb_off(topo, 2, 0);
    for (int i = 0; i < ch * 2 + k; ++i) off += (aesf_topo_ap(topo, i >> 1, i & 1) + 3) & ~3;
    return off;
}
// comb ring descriptor: read from the launch parameters, or constant under TOPO (ch, cc are
// unrolled loop counters, so everything below folds to immediates)
template <int TOPO, int FR>
__device__ __forceinline__ FRing aesf_comb_ring(const FastStage &st, int ch, int cc)
{
    if (TOPO) {
        const int L = aesf_topo_comb(TOPO, ch, cc), len = (L + 3) & ~3;
        FRing rg;
        rg.off = aesf_topo_comb_off(TOPO, ch, cc); rg.len = len; rg.lag = L; rg.tinc = (AES_NT * FR) % len;
        return rg;
    }
    return st.ring[ch][cc];
}

template <int FR> __device__ __forceinline__ void aesf_select4(const float4 A, const float4 B, int m, float (&o)[FR]);

// Comb slots are kept as BYTE offsets inside their ring, so a ring access is base + slot with the
// ring's constant offset folded into the instruction.
template <int FR>
__device__ __forceinline__ void aesf_comb_read(const float *ring, int wb, int m, int len, float (&o)[FR])
{
    static_assert(FR == 4, "");
    const char *rb = reinterpret_cast<const char *>(ring);
    const float4 A = aes_lds_v4(reinterpret_cast<const float *>(rb + wb));
    if (m == 0) { o[0] = A.x; o[1] = A.y; o[2] = A.z; o[3] = A.w; return; }
    int bb = wb + 16;
    bb = bb >= 4 * len ? bb - 4 * len : bb;
    const float4 B = aes_lds_v4(reinterpret_cast<const float *>(rb + bb));
    aesf_select4<FR>(A, B, m, o);
}

// ---- phase walk on the smem tile for a WALK ring described in the parameter bank --------
template <int FR, int OP>
__device__ __forceinline__ void aesf_walk(const FCtx &c, const FRing rg, int ring_id, bool glob, float p0, float p1, float p2)
{
    constexpr int T = AES_NT * FR;
    const int ch = c.tid >> 7, j0 = c.tid & 127;
    float *rb = (glob ? c.gscr : c.rings) + rg.off;
    const int L = rg.len, pos = c.rpos[ring_id];
    const int W = L < T ? L : T;
    float *s = c.tile + ch * T;
    const int len = c.len;
    for (int j = j0; j < W && j < len; j += 128) {
        int slot = pos + j;
        if (slot >= L) slot -= L;
        float line = (c.n0 + j >= L) ? rb[slot] : 0.0f;
        int i = j;
        for (; i + 3 * L < len; i += 4 * L) {           // 4 steps per trip: loads first, the walk is a serial chain
            float xs[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) xs[u] = s[i + u * L];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float x = xs[u];
                if (OP == 0) {
                    s[i + u * L] = aes_mix_clip(p1, x, p2, line);
                    line = fmaf(line, p0, x);
                } else if (OP == 1) {
                    s[i + u * L] = line;
                    line = x;
                } else {
                    const float yo = fmaf(-p0, x, line);
                    s[i + u * L] = yo;
                    line = fmaf(p0, yo, x);
                }
            }
        }
        for (; i < len; i += L) {
            const float x = s[i];
            if (OP == 0) {
                s[i] = aes_mix_clip(p1, x, p2, line);
                line = fmaf(line, p0, x);
            } else if (OP == 1) {
                s[i] = line;
                line = x;
            } else {
                const float yo = fmaf(-p0, x, line);
                s[i] = yo;
                line = fmaf(p0, yo, x);
            }
        }
        rb[slot] = line;
    }
}

// all-pass walk with a compile-time length L (< T): column j of the tile, seen as rows of L
// samples, is one serial chain  y = line - g*x,  line' = x + g*y  down the rows; the chain's
// state enters from / leaves to the ring (length L, phase `pos`).  Full tiles only.
template <int FR, int L>
__device__ __forceinline__ void aesf_ap_static(const FCtx &c, int off, int ring_id, float g)
{
    constexpr int T = AES_NT * FR;
    static_assert(L > 0 && L < T, "");
    const int ch = c.tid >> 7, j0 = c.tid & 127;
    float *rb = c.rings + off;
    const int pos = c.rpos[ring_id];
    float *s = c.tile + ch * T;
#pragma unroll
    for (int r = 0; r < (L + 127) / 128; ++r) {
        const int j = j0 + 128 * r;
        if (128 * r + 127 < L || j < L) {
            constexpr int KMAX = (T + L - 1) / L;           // rows that can hold column j
            int slot = pos + j;
            if (slot >= L) slot -= L;
            float line = rb[slot];
            if (c.n0 == 0) line = 0.0f;                     // L < T: only the first tile has no history
            float xs[KMAX];
#pragma unroll
            for (int k = 0; k < KMAX; ++k) {
                if (128 * r + 127 + k * L < T) xs[k] = s[j + k * L];            // row k is complete for this trip
                else if (128 * r + k * L < T) xs[k] = (j + k * L < T) ? s[j + k * L] : 0.0f;
            }
#pragma unroll
            for (int k = 0; k < KMAX; ++k) {
                if (128 * r + 127 + k * L < T) {
                    const float yo = fmaf(-g, xs[k], line);
                    s[j + k * L] = yo;
                    line = fmaf(g, yo, xs[k]);
                } else if (128 * r + k * L < T) {
                    if (j + k * L < T) {
                        const float yo = fmaf(-g, xs[k], line);
                        s[j + k * L] = yo;
                        line = fmaf(g, yo, xs[k]);
                    }
                }
            }
            rb[slot] = line;
        }
    }
}

// all-pass K of a reverb stage over the tile in shared memory (threads 0..127 left, 128..255 right)
template <int FR, int TOPO, int K>
__device__ __forceinline__ void aesf_allpass(const FastArgs &a, const FastStage &st, const FCtx &c)
{
    const int ch = c.tid >> 7;
    if (TOPO != 0 && c.len == AES_NT * FR) {            // full tile: unrolled walk, constant length
        constexpr int TP = TOPO ? TOPO : 1;
        if (ch == 0) aesf_ap_static<FR, aesf_topo_ap(TP, 0, K)>(c, aesf_topo_ap_off(TP, 0, K), st.walk_ap[0][K], st.a);
        else         aesf_ap_static<FR, aesf_topo_ap(TP, 1, K)>(c, aesf_topo_ap_off(TP, 1, K), st.walk_ap[1][K], st.a);
    } else {
        aesf_walk<FR, 2>(c, a.walk[st.walk_ap[ch][K]], st.walk_ap[ch][K], false, st.a, 0.f, 0.f);
    }
}

template <int FR> __device__ __forceinline__ void aesf_spill(const FCtx &c, const float (&v)[2][FR])
{
    constexpr int T = AES_NT * FR;
    aes_stv<FR>(c.tile + FR * c.tid, v[0]);
    aes_stv<FR>(c.tile + T + FR * c.tid, v[1]);
}
template <int FR> __device__ __forceinline__ void aesf_reload(const FCtx &c, float (&v)[2][FR])
{
    constexpr int T = AES_NT * FR;
    aes_ldv_sp<FR, 0>(c.tile + FR * c.tid, v[0]);
    aes_ldv_sp<FR, 0>(c.tile + T + FR * c.tid, v[1]);
}

// ---- register-path line (feedback delay / pre-delay with lag >= T) ------------------------
template <int FR>
__device__ __forceinline__ void aesf_line_init(const FRing rg, int i0, int &w, int &a0)
{
    w = i0;                                         // slot of sample n0 + i0 at n0 = 0
    const int lag4 = (rg.lag + 3) & ~3;
    a0 = i0 - lag4;
    if (a0 < 0) a0 += rg.len;                       // len = lag4 + T > lag4
}
template <int FR>
__device__ __forceinline__ void aesf_line_read(const FCtx &c, const FRing rg, bool glob, int a0, float (&o)[FR])
{
    const int m = ((rg.lag + 3) & ~3) - rg.lag;
    if (glob) aesf_read<FR, 1>(c.gscr + rg.off, a0, m, rg.len, o);
    else      aesf_read<FR, 0>(c.rings + rg.off, a0, m, rg.len, o);
}

template <int FR>
__device__ __forceinline__ void aesf_select4(const float4 A, const float4 B, int m, float (&o)[FR])
{
    static_assert(FR == 4 || FR == 2 || FR == 1, "");
    const float w[8] = { A.x, A.y, A.z, A.w, B.x, B.y, B.z, B.w };
    if (m == 0) {
#pragma unroll
        for (int j = 0; j < FR; ++j) o[j] = w[j];
    } else if (m == 1) {
#pragma unroll
        for (int j = 0; j < FR; ++j) o[j] = w[j + 1];
    } else if (m == 2) {
#pragma unroll
        for (int j = 0; j < FR; ++j) o[j] = w[j + 2];
    } else {
#pragma unroll
        for (int j = 0; j < FR; ++j) o[j] = w[j + 3];
    }
}

// ---- one stage, specialised on its shape code ------------------------------------------------
// Hermite read of the octaver ring: the four taps sit d0, d0-1, d0-2, d0-3 samples behind frame n,
// each wrapped into [0, size) as the reference's modulo does (octaver.py:41-58).  Away from the
// two ends of that range the taps are four consecutive slots and need no per-tap wrap.
__device__ __forceinline__ float aes_octaver_taps(const float *rb, int mask, int n, int size, int d0, float frac)
{
    float t[4];
    if (d0 >= 3 && d0 < size) {
        const int b = n - d0;
#pragma unroll
        for (int k = 0; k < 4; ++k) t[k] = rb[(b + k) & mask];
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            int d = d0 - k;
            if (d < 0) d += size;
            if (d >= size) d -= size;
            t[k] = rb[(n - d) & mask];
        }
    }
    return aes_hermite(frac, t[0], t[1], t[2], t[3]);
}

// stages that hand per-warp totals through shared memory (reverb, biquad, gate) alternate between
// two 32-double halves of the exchange area (WPAR = parity of the stage among them): a warp still
// reading stage k's totals is at most one user behind, and user k+2 only writes after user k+1's
// barrier -- so no stage needs a trailing barrier just to protect the area
#define AESF_USES_WTOT(code) (AESF_KIND(code) == AESK_REVERB || AESF_KIND(code) == AESK_BIQUAD || AESF_KIND(code) == AESK_GATE)

template <int FR, int CODE, int S, int TOPO, int WPAR>
__device__ __forceinline__ void aesf_stage(const FastArgs &a, const FCtx &c, SRegs &sr, float (&v)[2][FR],
                                           const float4 (&lnA)[2], const float4 (&lnB)[2], const double *sin, double *sout)
{
    constexpr int KIND = AESF_KIND(CODE);
    const FastStage &st = a.st[S];
    const int i0 = FR * c.tid, lane = c.lane, warp = c.warp;
    double *const wtot = c.wtot + WPAR * 32;

    if constexpr (KIND == AESK_DELAY) {
        if (AESF_MODE(CODE) == AES_MODE_REG) {
            const float fb = st.fb, dry = st.dry, wet = st.wet;
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
                const FRing rg = st.ring[ch][0];
                float line[FR];
                if (AESF_PF(CODE) && c.ln_stage != nullptr) {
                    // line samples staged by TMA a tile ago: [i0 + m, i0 + m + FR) of the staged span
                    constexpr int T = AES_NT * FR;
                    // (an unaligned line: the two aligned vectors around it and a CTA-uniform select -- four scalar
                    // loads at a 16-byte lane stride conflict 4-way)
                    const int m = ((rg.lag + 3) & ~3) - rg.lag;
                    const float *sp = c.ln_stage + ch * (T + 8) + i0;
                    const float4 A = aes_lds_v4(sp);
                    if (m == 0) {
                        line[0] = A.x; line[1 % FR] = A.y; line[2 % FR] = A.z; line[3 % FR] = A.w;
                    } else {
                        aesf_select4<FR>(A, aes_lds_v4(sp + 4), m, line);
                    }
                } else if (AESF_PF(CODE)) {
                    aesf_select4<FR>(lnA[ch], lnB[ch], ((rg.lag + 3) & ~3) - rg.lag, line);
                } else {
                    aesf_line_read<FR>(c, rg, st.glob != 0, sr.da[ch], line);
                }
                if (c.n0 < rg.lag) {                                // only the first tiles of a clip: zero history
#pragma unroll
                    for (int j = 0; j < FR; ++j)
                        if (c.n0 + i0 + j < rg.lag) line[j] = 0.0f;
                }
                float nb[FR];
#pragma unroll
                for (int j = 0; j < FR; ++j) {
                    const float x = v[ch][j];
                    nb[j] = fmaf(line[j], fb, x);
                    v[ch][j] = aes_mix_clip(dry, x, wet, line[j]);
                }
                if (st.glob) aes_stv<FR>(c.gscr + rg.off + sr.dw[ch], nb);
                else         aes_stv<FR>(c.rings + rg.off + sr.dw[ch], nb);
            }
        } else {
            aesf_spill<FR>(c, v);
            __syncthreads();
            const int ch = c.tid >> 7;
            aesf_walk<FR, 0>(c, a.walk[st.walk_pre[ch]], st.walk_pre[ch], st.glob != 0, st.fb, st.dry, st.wet);
            __syncthreads();
            aesf_reload<FR>(c, v);
        }
    } else if constexpr (KIND == AESK_REVERB) {
        constexpr int NC = AESF_NC(CODE), NA = AESF_NA(CODE), PM = AESF_PREMODE(CODE);
        float pre[2][FR];
        if constexpr (PM == 0) {
#pragma unroll
            for (int ch = 0; ch < 2; ++ch)
#pragma unroll
                for (int j = 0; j < FR; ++j) pre[ch][j] = v[ch][j];
        } else if constexpr (PM == 1) {
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
                const FRing rg = st.pre[ch];
                aesf_line_read<FR>(c, rg, st.glob != 0, sr.da[ch], pre[ch]);
                if (c.n0 < rg.lag) {
#pragma unroll
                    for (int j = 0; j < FR; ++j)
                        if (c.n0 + i0 + j < rg.lag) pre[ch][j] = 0.0f;
                }
                if (st.glob) aes_stv<FR>(c.gscr + rg.off + sr.dw[ch], v[ch]);
                else         aes_stv<FR>(c.rings + rg.off + sr.dw[ch], v[ch]);
            }
        } else {
            static_assert(PM == 3, "");
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
                const FRing rg = st.pre[ch];
                if (st.glob) aes_stv<FR>(c.gscr + rg.off + sr.dw[ch], v[ch]);
                else         aes_stv<FR>(c.rings + rg.off + sr.dw[ch], v[ch]);
            }
            __syncthreads();
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
                const FRing rg = st.pre[ch];
                aesf_line_read<FR>(c, rg, st.glob != 0, sr.da[ch], pre[ch]);
                if (c.n0 < rg.lag) {
#pragma unroll
                    for (int j = 0; j < FR; ++j)
                        if (c.n0 + i0 + j < rg.lag) pre[ch][j] = 0.0f;
                }
            }
        }

        // damped combs on u = lp/(1-h)
        const float h = st.h, hw = st.hp[5];
        const float hl = a.lane_tab[(S * 32 + lane) * FAST_LANE_STRIDE];
        float y[2][NC][FR], e[2][NC];
#pragma unroll
        for (int ch = 0; ch < 2; ++ch)
#pragma unroll
            for (int cc = 0; cc < NC; ++cc) {
                const FRing rg = aesf_comb_ring<TOPO, FR>(st, ch, cc);
                aesf_comb_read<FR>(c.rings + rg.off, sr.w[ch][cc], rg.len - rg.lag, rg.len, y[ch][cc]);
                float u = 0.0f;
#pragma unroll
                for (int j = 0; j < FR; ++j) u = fmaf(h, u, y[ch][cc][j]);
                e[ch][cc] = u;
            }
        const int nscan = st.nscan;
        for (int s = 0; s < nscan; ++s) {
            const float m = st.hp[s];
#pragma unroll
            for (int ch = 0; ch < 2; ++ch)
#pragma unroll
                for (int cc = 0; cc < NC; ++cc) {
                    const float t = __shfl_up_sync(0xffffffffu, e[ch][cc], 1 << s);
                    if (lane >= (1 << s)) e[ch][cc] = fmaf(m, t, e[ch][cc]);
                }
        }
        float *wt = reinterpret_cast<float *>(wtot);            // [8 warps][2][4]
        if (lane == 31) {
#pragma unroll
            for (int ch = 0; ch < 2; ++ch)
#pragma unroll
                for (int cc = 0; cc < NC; ++cc) wt[(warp * 2 + ch) * 4 + cc] = e[ch][cc];
        }
        __syncthreads();
        const int nxw = st.nxw;
        const int u0 = warp > nxw ? warp - nxw : 0;
        const float *fsin = reinterpret_cast<const float *>(sin);   // comb carries are kept as f32
        float *fsout = reinterpret_cast<float *>(sout);
        float sum[2][FR];
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
#pragma unroll
            for (int j = 0; j < FR; ++j) sum[ch][j] = 0.0f;
#pragma unroll
            for (int cc = 0; cc < NC; ++cc) {
                const float ex = __shfl_up_sync(0xffffffffu, e[ch][cc], 1);
                float C;
                if (nxw == 1) {                                  // usual case: h^(32*FR) < 2^-32
                    C = warp == 0 ? fsin[ch * 4 + cc] : wt[((warp - 1) * 2 + ch) * 4 + cc];
                } else {
                    C = u0 == 0 ? fsin[ch * 4 + cc] : 0.0f;
                    for (int t = u0; t < warp; ++t) C = fmaf(hw, C, wt[(t * 2 + ch) * 4 + cc]);
                }
                float u = fmaf(hl, C, lane == 0 ? 0.0f : ex);
                const FRing rg = aesf_comb_ring<TOPO, FR>(st, ch, cc);
                const float gs = st.gs[ch][cc];
                float nb[FR];
#pragma unroll
                for (int j = 0; j < FR; ++j) {
                    u = fmaf(h, u, y[ch][cc][j]);
                    nb[j] = fmaf(gs, u, pre[ch][j]);            // buf[n] = x + g*(1-h)*u
                    sum[ch][j] = __fadd_rn(sum[ch][j], y[ch][cc][j]);
                }
                aes_stv<FR>(reinterpret_cast<float *>(reinterpret_cast<char *>(c.rings + rg.off) + sr.w[ch][cc]), nb);
                if (c.tid == AES_NT - 1) fsout[ch * 4 + cc] = u;
            }
        }
        aesf_spill<FR>(c, sum);
        __syncthreads();
        static_assert(NA <= 2, "");
        if (NA > 0) { aesf_allpass<FR, TOPO, 0>(a, st, c); __syncthreads(); }
        if (NA > 1) { aesf_allpass<FR, TOPO, 1>(a, st, c); __syncthreads(); }
        aesf_reload<FR>(c, sum);
        const float dry = st.dry, wet = st.wet;
#pragma unroll
        for (int ch = 0; ch < 2; ++ch)
#pragma unroll
            for (int j = 0; j < FR; ++j) v[ch][j] = aes_mix_clip(dry, v[ch][j], wet, sum[ch][j]);
    } else if constexpr (KIND == AESK_BIQUAD) {
        // Transposed direct form II: y = b0*x + s1, s1' = b1*x - a1*y + s2, s2' = b2*x - a2*y.  The
        // 2-vector (s1, s2) is the whole state -- no input history -- so a thread needs nothing from
        // its neighbours before the scan.  Transition matrix A^T (A: the DF-I companion matrix whose
        // powers the plan tabulates).  Pass 1 runs the chunk from zero state and keeps the outputs;
        // the carried-in state S then adds row0((A^T)^j) . S to output j (independent FMAs).
        const double b0 = st.bq[0], b1 = st.bq[1], b2 = st.bq[2], a1 = st.bq[3], a2 = st.bq[4];
        double yz[2][FR], e1[2], e2[2];
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
            double s1 = 0.0, s2 = 0.0;
#pragma unroll
            for (int j = 0; j < FR; ++j) {
                const double xj = (double)v[ch][j];
                const double y = fma(b0, xj, s1);
                s1 = fma(b1, xj, fma(-a1, y, s2));
                s2 = fma(b2, xj, -a2 * y);
                yz[ch][j] = y;
            }
            e1[ch] = s1; e2[ch] = s2;
        }
        const int bq_nscan = st.nscan;                  // scan steps / previous warps that still matter (aes_fast_build.h)
        for (int s = 0; s < bq_nscan; ++s) {
            const double m0 = st.bq_pow[s][0], m1 = st.bq_pow[s][1], m2 = st.bq_pow[s][2], m3 = st.bq_pow[s][3];
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
                const double u1 = __shfl_up_sync(0xffffffffu, e1[ch], 1 << s);
                const double u2 = __shfl_up_sync(0xffffffffu, e2[ch], 1 << s);
                if (lane >= (1 << s)) {                          // (ptxas turns a predicated DFMA into DFMA + 2 FSEL too)
                    e1[ch] = fma(m0, u1, fma(m2, u2, e1[ch]));   // += (A^k)^T u, two dependent DFMAs
                    e2[ch] = fma(m1, u1, fma(m3, u2, e2[ch]));
                }
            }
        }
        if (lane == 31) {
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
                wtot[(warp * 2 + ch) * 2] = e1[ch];
                wtot[(warp * 2 + ch) * 2 + 1] = e2[ch];
            }
        }
        __syncthreads();
        const double w0 = st.bq_pow[5][0], w1 = st.bq_pow[5][1], w2 = st.bq_pow[5][2], w3 = st.bq_pow[5][3];
        const double *lt = reinterpret_cast<const double *>(a.lane_tab + (S * 32 + lane) * FAST_LANE_STRIDE + 4);
        const double l0 = lt[0], l1 = lt[1], l2 = lt[2], l3 = lt[3];
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
            // state carried between tiles is the reference's (x1, x2, y1, y2) (filter.py:35-40)
            const double cx1 = sin[4 * ch + 0], cx2 = sin[4 * ch + 1], cy1 = sin[4 * ch + 2], cy2 = sin[4 * ch + 3];
            double C1 = b1 * cx1 + b2 * cx2 - a1 * cy1 - a2 * cy2;
            double C2 = b2 * cx1 - a2 * cy1;
            const int u0 = warp > st.nxw ? warp - st.nxw : 0;
            if (u0 > 0) { C1 = 0.0; C2 = 0.0; }                 // the tile-start state has decayed by then
            for (int t = u0; t < warp; ++t) {
                const double t1 = fma(w0, C1, fma(w2, C2, wtot[(t * 2 + ch) * 2]));
                const double t2 = fma(w1, C1, fma(w3, C2, wtot[(t * 2 + ch) * 2 + 1]));
                C1 = t1; C2 = t2;
            }
            double x1 = __shfl_up_sync(0xffffffffu, e1[ch], 1), x2 = __shfl_up_sync(0xffffffffu, e2[ch], 1);
            if (lane == 0) { x1 = 0.0; x2 = 0.0; }
            const double S1 = fma(l0, C1, fma(l2, C2, x1));      // state entering this thread's chunk
            const double S2 = fma(l1, C1, fma(l3, C2, x2));
#pragma unroll
            for (int j = 0; j < FR; ++j) {
                const double y = j == 0 ? yz[ch][0] + S1 : fma(st.bq_row[j][0], S1, fma(st.bq_row[j][1], S2, yz[ch][j]));
                // the last two frames of the tile are the next tile's (x1, x2, y1, y2): in a full tile
                // (FR >= 2) they sit in the last thread -- one uniform branch instead of 4 predicated stores
                // per frame
                if (FR >= 2 && c.len == AES_NT * FR) {
                    if (j >= FR - 2 && c.tid == AES_NT - 1) {
                        sout[4 * ch + (FR - 1 - j)] = (double)v[ch][j]; sout[4 * ch + 2 + (FR - 1 - j)] = y;
                    }
                } else {
                    if (i0 + j == c.len - 1) { sout[4 * ch + 0] = (double)v[ch][j]; sout[4 * ch + 2] = y; }
                    if (i0 + j == c.len - 2) { sout[4 * ch + 1] = (double)v[ch][j]; sout[4 * ch + 3] = y; }
                }
                v[ch][j] = (float)y;
            }
            if (c.len == 1 && c.tid == 0) { sout[4 * ch + 1] = cx1; sout[4 * ch + 3] = cy1; }
        }
    } else if constexpr (KIND == AESK_GATE) {
        const double thr = st.thr, ka = 1.0 - st.att, kr = 1.0 - st.rel, att = st.att;
        bool open[FR];
        double A = 1.0, Bv = 0.0;
#pragma unroll
        for (int f = 0; f < FR; ++f) {
            const float lvl = fmaxf(fabsf(v[0][f]), fabsf(v[1][f]));
            open[f] = (double)lvl > thr;
            const double am = open[f] ? ka : kr, bm = open[f] ? att : 0.0;
            Bv = am * Bv + bm;
            A = am * A;
        }
#pragma unroll
        for (int s = 0; s < 5; ++s) {
            const double Au = __shfl_up_sync(0xffffffffu, A, 1 << s);
            const double Bu = __shfl_up_sync(0xffffffffu, Bv, 1 << s);
            if (lane >= (1 << s)) { Bv = A * Bu + Bv; A = A * Au; }
        }
        if (lane == 31) { wtot[2 * warp] = A; wtot[2 * warp + 1] = Bv; }
        __syncthreads();
        double g = sin[0];
        for (int t = 0; t < warp; ++t) g = wtot[2 * t] * g + wtot[2 * t + 1];
        double Ae = __shfl_up_sync(0xffffffffu, A, 1), Be = __shfl_up_sync(0xffffffffu, Bv, 1);
        if (lane == 0) { Ae = 1.0; Be = 0.0; }
        g = Ae * g + Be;
#pragma unroll
        for (int f = 0; f < FR; ++f) {
            g = (open[f] ? ka : kr) * g + (open[f] ? att : 0.0);
            const float gf = (float)g;
            v[0][f] *= gf;
            v[1][f] *= gf;
            if (c.len == AES_NT * FR) { if (f == FR - 1 && c.tid == AES_NT - 1) sout[0] = g; }
            else if (i0 + f == c.len - 1) sout[0] = g;
        }
    } else if constexpr (KIND == AESK_OCTAVER) {
        float *rb = c.rings + st.ring[0][0].off;
        const int mask = st.oct_mask, size = st.oct_size, half = size >> 1;
        const bool even = (size & 1) == 0;
        float mono[FR];
#pragma unroll
        for (int j = 0; j < FR; ++j) mono[j] = __fmul_rn(__fadd_rn(v[0][j], v[1][j]), 0.5f);
        aes_stv<FR>(rb + ((c.n0 + i0) & mask), mono);
        __syncthreads();
        const float wet_g = st.mix, dry_g = (float)(1.0 - (double)st.mix);
        const double step = st.step, fsize = st.fsize;
        // phasor of the thread's first frame in closed form, then one add (+ wrap) per frame
        double ph = st.ph0 + (double)(c.n0 + i0) * step;
        ph -= floor(ph);
#pragma unroll
        for (int j = 0; j < FR; ++j) {
            const int n = c.n0 + i0 + j;
            // grain 1 reads size*(1-ph) samples behind the write pointer (octaver.py:39); grain 2 runs
            // half a period apart, so for an even ring its read position is exactly size/2 away:
            // same fraction, integer part shifted -- one f64 -> (int, frac) split serves both
            const double raw = fsize - ph * fsize;
            const int m = (int)raw;
            const float frac = (float)(raw - (double)m);
            float s1, s2;
            s1 = aes_octaver_taps(rb, mask, n, size, size - m + 1, frac);
            if (even) {
                const int m2 = ph < 0.5 ? m - half : m + half;
                s2 = aes_octaver_taps(rb, mask, n, size, size - m2 + 1, frac);
            } else {
                double p2 = ph + 0.5;
                if (p2 >= 1.0) p2 -= 1.0;
                s2 = aes_octaver_tap(rb, mask, n, size, fsize, p2);
            }
            const float sn = sinpif((float)ph);
            const float g1 = sn * sn;
            const float g2 = 1.0f - g1;
            const float wet = __fmul_rn(s1 * g1 + s2 * g2, wet_g);
            v[0][j] = __fadd_rn(__fmul_rn(v[0][j], dry_g), wet);
            v[1][j] = __fadd_rn(__fmul_rn(v[1][j], dry_g), wet);
            ph += step;
            if (ph >= 1.0) ph -= 1.0;
            else if (ph < 0.0) ph += 1.0;
        }
    } else if constexpr (KIND == AESK_DISTORTION) {
        const float drive = st.drive, mix = st.mix, dry = 1.0f - st.mix;
#pragma unroll
        for (int ch = 0; ch < 2; ++ch)
#pragma unroll
            for (int j = 0; j < FR; ++j) {
                const float x = v[ch][j];
                const float t = tanhf(__fmul_rn(drive, x));
                v[ch][j] = aes_clip1(__fadd_rn(__fmul_rn(dry, x), __fmul_rn(mix, t)));
            }
    }
}

// per-clip initialisation / per-tile advance of a stage's register slots
template <int FR, int CODE, int S>
__device__ __forceinline__ void aesf_slots_init(const FastArgs &a, int i0, SRegs &sr)
{
    constexpr int KIND = AESF_KIND(CODE);
    const FastStage &st = a.st[S];
    if (KIND == AESK_DELAY && AESF_MODE(CODE) == AES_MODE_REG) {
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) aesf_line_init<FR>(st.ring[ch][0], i0, sr.dw[ch], sr.da[ch]);
    }
    if constexpr (KIND == AESK_REVERB) {
        if (AESF_PREMODE(CODE) != 0) {
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) aesf_line_init<FR>(st.pre[ch], i0, sr.dw[ch], sr.da[ch]);
        }
#pragma unroll
        for (int ch = 0; ch < 2; ++ch)
#pragma unroll
            for (int cc = 0; cc < AESF_NC(CODE); ++cc) sr.w[ch][cc] = 4 * i0;         // byte offset in the ring
    }
}

template <int FR, int CODE, int S, int TOPO>
__device__ __forceinline__ void aesf_slots_advance(const FastArgs &a, SRegs &sr)
{
    constexpr int KIND = AESF_KIND(CODE);
    const FastStage &st = a.st[S];
    if (KIND == AESK_DELAY && AESF_MODE(CODE) == AES_MODE_REG) {
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
            const FRing rg = st.ring[ch][0];
            sr.dw[ch] = aesf_adv(sr.dw[ch], rg.tinc, rg.len);
            sr.da[ch] = aesf_adv(sr.da[ch], rg.tinc, rg.len);
        }
    }
    if constexpr (KIND == AESK_REVERB) {
        if (AESF_PREMODE(CODE) != 0) {
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
                const FRing rg = st.pre[ch];
                sr.dw[ch] = aesf_adv(sr.dw[ch], rg.tinc, rg.len);
                sr.da[ch] = aesf_adv(sr.da[ch], rg.tinc, rg.len);
            }
        }
#pragma unroll
        for (int ch = 0; ch < 2; ++ch)
#pragma unroll
            for (int cc = 0; cc < AESF_NC(CODE); ++cc) {
                const FRing rg = aesf_comb_ring<TOPO, FR>(st, ch, cc);
                sr.w[ch][cc] = aesf_adv(sr.w[ch][cc], 4 * rg.tinc, 4 * rg.len);
            }
    }
}

// prefetch of the long feedback-delay line one tile ahead (lag >= 2T): issued at the top
// of tile i for tile i+1, consumed a full tile of work later.
template <int FR, int CODE, int S>
__device__ __forceinline__ void aesf_prefetch(const FastArgs &a, const FCtx &c, const SRegs &sr, int ahead,
                                              float4 (&pfA)[2], float4 (&pfB)[2])
{
    if (AESF_KIND(CODE) == AESK_DELAY && AESF_PF(CODE)) {
        static_assert(!(AESF_KIND(CODE) == AESK_DELAY && AESF_PF(CODE)) || FR == 4, "prefetch path is FR=4 only");
        const FastStage &st = a.st[S];                  // PF lines always live in the global scratch
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
            const FRing rg = st.ring[ch][0];
            const int a0 = ahead ? aesf_adv(sr.da[ch], rg.tinc, rg.len) : sr.da[ch];
            int b0 = a0 + 4;
            b0 = b0 >= rg.len ? b0 - rg.len : b0;
            const float *rb = c.gscr + rg.off;
            pfA[ch] = aes_ldg_v4(rb + a0);              // straight-line: no branch, no copy of the results
            pfB[ch] = aes_ldg_v4(rb + b0);
        }
    }
}

// issue the TMA copies for one tile (called by thread 0 only): input frames and, for a PF
// delay stage PS, the T+4 line samples per channel starting at thread 0's aligned read base
// `lnbase`: thread 0's aligned read base of each channel's line for the tile being staged, tracked
// by every thread in CTA-uniform arithmetic (so the copy operands need no per-lane broadcast)
template <int FR, int PCODE, int PS>
__device__ __forceinline__ void aesf_issue_tile(const FastArgs &a, const FCtx &c, const int (&lnbase)[2],
                                                long long frame0, float *stage_x, float *stage_ln,
                                                unsigned long long *bar)
{
    constexpr int T = AES_NT * FR;
    constexpr bool PF = AESF_KIND(PCODE) == AESK_DELAY && AESF_PF(PCODE);
    aes_mbar_expect(bar, (unsigned)(T * 8 + (PF ? 2 * (T + 4) * 4 : 0)));
    aes_bulk_g2s(stage_x, reinterpret_cast<const float *>(a.x) + 2 * frame0, T * 8, bar, aes_policy_evict_first());
    if (PF) {
        aes_fence_proxy_async();                    // the line was written with ordinary stores >= 1 tile ago
        const unsigned long long keep = aes_policy_evict_last();
        const FastStage &st = a.st[PS];
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
            const FRing rg = st.ring[ch][0];
            const int a0 = lnbase[ch];
            const float *rb = c.gscr + rg.off;
            float *dst = stage_ln + ch * (T + 8);
            const int n1 = (rg.len - a0) < (T + 4) ? (rg.len - a0) : (T + 4);
            aes_bulk_g2s(dst, rb + a0, (unsigned)n1 * 4, bar, keep);
            if (n1 < T + 4) aes_bulk_g2s(dst + n1, rb, (unsigned)(T + 4 - n1) * 4, bar, keep);   // the ring wraps
        }
    }
#ifdef AES_CPU_EMU
    aes_mbar_complete_emu(bar);
#endif
}

template <int FR, int C0, int C1, int C2, int C3, int TOPO = 0>
__device__ void aes_fast_body(const FastArgs &a)
{
    constexpr int T = AES_NT * FR;
    constexpr int NS = (C0 != 0) + (C1 != 0) + (C2 != 0) + (C3 != 0);
    // the (single) prefetched feedback-delay stage, if the shape has one
    constexpr int PS = (AESF_KIND(C0) == AESK_DELAY && AESF_PF(C0)) ? 0
                     : (AESF_KIND(C1) == AESK_DELAY && AESF_PF(C1)) ? 1
                     : (AESF_KIND(C2) == AESK_DELAY && AESF_PF(C2)) ? 2
                     : (AESF_KIND(C3) == AESK_DELAY && AESF_PF(C3)) ? 3 : -1;
    constexpr int PCODE = PS == 0 ? C0 : PS == 1 ? C1 : PS == 2 ? C2 : PS == 3 ? C3 : 0;
    constexpr bool STAGED = FR == 4;               // TMA staging needs 16-byte granules per thread
    constexpr int LASTC = C3 ? C3 : C2 ? C2 : C1 ? C1 : C0;
    constexpr bool ENDS_IN_REVERB = AESF_KIND(LASTC) == AESK_REVERB;
    AES_DYN_SMEM(float, smem);
    FCtx c;
    c.tid = threadIdx.x;
    c.lane = c.tid & 31;
    c.warp = c.tid >> 5;
    c.tile = smem;
    c.rings = smem + 2 * T;
    const int foff = (2 * T + a.smem_floats + 3) & ~3;
    c.wtot = reinterpret_cast<double *>(smem + foff);
    double *state = c.wtot + 64;                    // [2][NS*8]
    constexpr int NST = NS * 8;
    int *rpos2 = reinterpret_cast<int *>(state + 2 * NST);
    // staging area (16-byte aligned): 2 x input tile, 2 x [2 ch][T+8] line samples, 2 mbarriers
    float *stg = reinterpret_cast<float *>((reinterpret_cast<size_t>(rpos2 + 2 * AESF_MAX_WALK) + 15) & ~(size_t)15);
    float *stage_x = stg;                           // [2][2T]
    float *stage_ln = stg + 4 * T;                  // [2][2][T+8]
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(stage_ln + (PS >= 0 ? 4 * (T + 8) : 0));
    c.gscr = a.scratch + (long long)blockIdx.x * a.scratch_floats;
    c.ln_stage = nullptr;
    const int nw = a.n_walk, i0 = FR * c.tid;
    if (STAGED) {
        if (c.tid == 0) { aes_mbar_init(bars, 1); aes_mbar_init(bars + 1, 1); aes_mbar_init_fence(); }
        __syncthreads();
    }
    unsigned it_issue = 0, it_wait = 0;             // tiles issued / consumed through the staging ring (CTA-uniform)
    ChainArgs io;                                   // tile I/O helpers are shared with the generic kernel
    io.x = a.x; io.y = a.y; io.N = a.N; io.in_fmt = a.in_fmt; io.out_fmt = a.out_fmt;

    for (long long b = blockIdx.x; b < a.B; b += gridDim.x) {
        for (int i = c.tid; i < a.smem_floats; i += AES_NT) c.rings[i] = 0.0f;
        for (int i = c.tid; i < nw; i += AES_NT) rpos2[i] = 0;
        for (int i = c.tid; i < NST; i += AES_NT) state[i] = a.init[i >> 3][i & 7];
        SRegs sr0, sr1, sr2, sr3;
        if (C0) aesf_slots_init<FR, C0, 0>(a, i0, sr0);
        if (C1) aesf_slots_init<FR, C1, 1>(a, i0, sr1);
        if (C2) aesf_slots_init<FR, C2, 2>(a, i0, sr2);
        if (C3) aesf_slots_init<FR, C3, 3>(a, i0, sr3);
        int lnbase[2] = { 0, 0 };                       // staged line: thread 0's read base, next tile to stage
        if (PS >= 0) {
            const FastStage &pst = a.st[PS < 0 ? 0 : PS];
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
                int w0;
                aesf_line_init<FR>(pst.ring[ch][0], 0, w0, lnbase[ch]);
            }
        }
        __syncthreads();

        // a tile goes through TMA staging when it is full, 16-byte aligned and plain f32 stereo
        const int N = (int)a.N;                         // the host routes clips of 2^31 frames or more to the generic kernel
        const bool clip_staged = STAGED && a.in_fmt == AESK_F32_STEREO && ((b * a.N) & 1) == 0;
        const long long clip_frame0 = b * a.N;
        if (clip_staged && N >= T && c.tid == 0)
            aesf_issue_tile<FR, PCODE, PS < 0 ? 0 : PS>(a, c, lnbase, clip_frame0, stage_x + (it_issue & 1) * 2 * T,
                                                       stage_ln + (it_issue & 1) * 2 * (T + 8), bars + (it_issue & 1));
        if (clip_staged && N >= T) ++it_issue;
        float4 lnA[2], lnB[2];
        lnA[0] = lnA[1] = lnB[0] = lnB[1] = make_float4(0.f, 0.f, 0.f, 0.f);
        int par = 0;
        for (int n0 = 0; n0 < N; n0 += T, par ^= 1) {
            c.n0 = n0;
            const int rem = N - n0;
            c.len = rem < T ? rem : T;
            c.rpos = rpos2 + par * nw;
            float v[2][FR];
            const bool staged = clip_staged && c.len == T;
            if (staged) {
                const unsigned q = it_wait & 1;
                aes_mbar_wait(bars + q, it_wait >> 1);
                ++it_wait;
                const float *sx = stage_x + q * 2 * T + 2 * i0;          // interleaved L R L R ...
                if constexpr (FR == 4) {
                    // 32 bytes per thread: two 128-bit loads at a 32-byte lane stride conflict 2-way (ncu r2ad, Robot
                    // Voice: 8.0 wavefronts per load against 4.0), so lanes 4..7 of every eight take their halves in
                    // the other order (as aes_rv_kernel does)
                    const int sw = (c.tid >> 2) & 1;
                    const float4 ta = aes_lds_v4(sx + 4 * sw), tb = aes_lds_v4(sx + 4 * (sw ^ 1));
                    const float4 t0 = sw ? tb : ta, t1 = sw ? ta : tb;
                    v[0][0] = t0.x; v[1][0] = t0.y; v[0][1 % FR] = t0.z; v[1][1 % FR] = t0.w;
                    v[0][2 % FR] = t1.x; v[1][2 % FR] = t1.y; v[0][3 % FR] = t1.z; v[1][3 % FR] = t1.w;
                } else {
#pragma unroll
                    for (int j = 0; j < FR / 2; ++j) {
                        const float4 t = aes_lds_v4(sx + 4 * j);
                        v[0][2 * j] = t.x; v[1][2 * j] = t.y; v[0][(2 * j + 1) % FR] = t.z; v[1][(2 * j + 1) % FR] = t.w;
                    }
                }
                c.ln_stage = PS >= 0 ? stage_ln + q * 2 * (T + 8) : nullptr;
            } else {
                aes_load_frames<FR>(io, b, n0, c.len, c.tid, v);         // ragged / unaligned / non-f32 tiles
                c.ln_stage = nullptr;
                if (PS >= 0) {
                    if (C0) aesf_prefetch<FR, C0, 0>(a, c, sr0, 0, lnA, lnB);
                    if (C1) aesf_prefetch<FR, C1, 1>(a, c, sr1, 0, lnA, lnB);
                    if (C2) aesf_prefetch<FR, C2, 2>(a, c, sr2, 0, lnA, lnB);
                    if (C3) aesf_prefetch<FR, C3, 3>(a, c, sr3, 0, lnA, lnB);
                }
            }
            // next tile: hand it to the TMA now, a whole tile of work ahead of its use
            if (PS >= 0) {                              // the line base moves on with every tile, staged or not
                const FastStage &pst = a.st[PS < 0 ? 0 : PS];
#pragma unroll
                for (int ch = 0; ch < 2; ++ch) lnbase[ch] = aesf_adv(lnbase[ch], pst.ring[ch][0].tinc, pst.ring[ch][0].len);
            }
            if (clip_staged && rem >= 2 * T) {
                if (c.tid == 0)
                    aesf_issue_tile<FR, PCODE, PS < 0 ? 0 : PS>(a, c, lnbase, clip_frame0 + n0 + T,
                                                               stage_x + (it_issue & 1) * 2 * T,
                                                               stage_ln + (it_issue & 1) * 2 * (T + 8), bars + (it_issue & 1));
                ++it_issue;
            }
            const double *sin = state + par * NST;
            double *sout = state + (par ^ 1) * NST;
            if (C0) aesf_stage<FR, C0, 0, TOPO, (0) & 1>(a, c, sr0, v, lnA, lnB, sin, sout);
            if (C1) aesf_stage<FR, C1, 1, TOPO, (AESF_USES_WTOT(C0)) & 1>(a, c, sr1, v, lnA, lnB, sin + 8, sout + 8);
            if (C2) aesf_stage<FR, C2, 2, TOPO, (AESF_USES_WTOT(C0) + AESF_USES_WTOT(C1)) & 1>(a, c, sr2, v, lnA, lnB, sin + 16, sout + 16);
            if (C3) aesf_stage<FR, C3, 3, TOPO, (AESF_USES_WTOT(C0) + AESF_USES_WTOT(C1) + AESF_USES_WTOT(C2)) & 1>(a, c, sr3, v, lnA, lnB, sin + 24, sout + 24);
            aes_store_frames<FR>(io, b, n0, c.len, c.tid, v);
            if (C0) aesf_slots_advance<FR, C0, 0, TOPO>(a, sr0);
            if (C1) aesf_slots_advance<FR, C1, 1, TOPO>(a, sr1);
            if (C2) aesf_slots_advance<FR, C2, 2, TOPO>(a, sr2);
            if (C3) aesf_slots_advance<FR, C3, 3, TOPO>(a, sr3);
            if (c.tid < nw) {
                const FRing rg = a.walk[c.tid];
                int p = c.rpos[c.tid] + rg.tinc;
                if (p >= rg.len) p -= rg.len;
                rpos2[(par ^ 1) * nw + c.tid] = p;
            }
            // A chain that ENDS in a reverb needs no tile-end barrier: after the reverb's last barrier
            // (behind the second all-pass) a thread only touches its own tile entries, registers and
            // parity-switched words, and everything the next tile overwrites (staging slot, exchange
            // area, carried state of the other parity, ring slots) was last read before that barrier.
            if (!ENDS_IN_REVERB) __syncthreads();
        }
        if (a.state_out != nullptr)
            for (int i = c.tid; i < NST; i += AES_NT)
                a.state_out[b * (16 * NS) + (i >> 3) * 16 + (i & 7)] = state[par * NST + i];
    }
}

#ifndef AESF_MIN_CTAS
#define AESF_MIN_CTAS 2
#endif
#ifndef AES_CPU_EMU
template <int FR, int C0, int C1, int C2, int C3, int TOPO, int MIN_CTAS = AESF_MIN_CTAS>
__global__ void __launch_bounds__(AES_NT, MIN_CTAS) aes_fast_kernel(const __grid_constant__ FastArgs a)
{
    aes_fast_body<FR, C0, C1, C2, C3, TOPO>(a);
}
#endif
