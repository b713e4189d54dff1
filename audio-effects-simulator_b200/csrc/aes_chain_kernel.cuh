// aes_chain_kernel.cuh -- the fused effect-chain kernel (device code), revision 2.
//
// One CTA owns one clip at a time (persistent grid, clips strided over CTAs) and
// walks it in time tiles of T = 256*FR stereo frames.  Every thread keeps FR
// consecutive frames x 2 channels of the tile IN REGISTERS from the HBM load to the
// HBM store; a tile is read once and written once: 8 algorithmic bytes per output
// sample whatever the chain length.  The next tile (and the long delay
// line it needs) is prefetched into registers while the current one is computed.
//
// All recurrences of the reference are kept, only re-associated so a tile is parallel:
//
//   * lag-L lines with L >= T (feedback delay delay.py:7-22, pre-delay reverb.py:11-31):
//     every delayed sample of the tile predates the tile, so the line is elementwise;
//     "aligned" rings (period a multiple of 4) make the 4 frames of a thread one float4.
//   * damped comb (reverb.py:33-46): L >= T as well; the one-pole inside the feedback
//     path is a constant-coefficient first-order scan: FR samples serially per thread,
//     warp-shuffle Kogge-Stone across lanes (steps with h^(FR*2^s) < 2^-32 are skipped),
//     carry chain across the 8 warps through shared memory (truncated the same way).
//   * short lines (all-pass reverb.py:48-67, any lag < T): "phase walk" on the tile in
//     shared memory: thread j walks samples j, j+L, j+2L.. with the line in a register.
//   * biquad DF-I (filter.py:8-40): scan with 2x2 companion-matrix powers, all f64.
//   * gate (gate.py:6-42): the branch depends only on the input (gain in [0,1], target in
//     {0,1}), so the gain is an affine scan with per-sample coefficients; f64.
//   * octaver (octaver.py:17-82): no feedback; gather from the mono history, phasor in
//     closed form frac(ph0 + n*step).
//
// The file is also compiled by g++ against tests/cpu_emu/cuda_emu.h (AES_CPU_EMU) so the
// logic can be exercised without a GPU; the product build is nvcc only.
#pragma once
#include "aes_plan.h"

#ifndef AES_CPU_EMU
#define AES_DYN_SMEM(type, name)                                             \
    extern __shared__ __align__(16) unsigned char name##_raw_[];             \
    type *name = reinterpret_cast<type *>(name##_raw_)
#endif

// sample formats (mirror aes_format in include/aesim.h)
#define AESK_F32_STEREO 0
#define AESK_F32_MONO 1
#define AESK_I16_DOWNMIX 2
#define AESK_I16_STEREO 3

// stage kinds (mirror aes_stage_kind)
#define AESK_DELAY 1
#define AESK_REVERB 2
#define AESK_BIQUAD 3
#define AESK_GATE 4
#define AESK_OCTAVER 5
#define AESK_DISTORTION 6

struct ChainArgs {
    const DevPlan *plan;     // device memory
    const void *x;
    void *y;
    long long B, N;
    float *scratch;          // gridDim.x * plan->scratch_floats
    int in_fmt, out_fmt;
    double *state_out;       // optional: [B][plan->n_state] carried scalars at the end of each clip
};

struct KCtx {
    const DevPlan *P;
    float *tile, *rings, *gscr;
    double *wtot;
    int *rpos;
    long long n0;
    int len, tid, lane, warp;
};

__device__ __forceinline__ float aes_clip1(float v) { return fminf(fmaxf(v, -1.0f), 1.0f); }

// numpy evaluates dry*x + wet*w as three separately rounded f32 ufuncs
// (delay.py:94-96, reverb.py:275-277): keep the products and the sum unfused.
__device__ __forceinline__ float aes_mix_clip(float dry, float x, float wet, float w)
{
    return aes_clip1(__fadd_rn(__fmul_rn(dry, x), __fmul_rn(wet, w)));
}

__device__ __forceinline__ float *aes_ring_base(const KCtx &c, const DevRing &r)
{
    return (r.space == AES_SPACE_GLOBAL ? c.gscr : c.rings) + r.off;
}

// 128-bit loads that the optimiser may not split: NVVM otherwise scalarises a float4 load
// whose elements are consumed in different branches of the misalignment switch (ncu showed
// 141 LDS.32 + 33 LDS.64 with 2-way bank conflicts instead of ~30 LDS.128).
#ifndef AES_CPU_EMU
__device__ __forceinline__ float4 aes_lds_v4(const float *p)
{
    float4 r;
    const unsigned a = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(a));
    return r;
}
__device__ __forceinline__ float4 aes_ldg_v4(const float *p)
{
    float4 r;
    asm volatile("ld.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
// 128-bit store to GLOBAL memory through a pointer whose address space the compiler cannot see
__device__ __forceinline__ void aes_stg_v4(float *p, const float (&v)[4])
{
    asm volatile("st.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
}
#else
static inline void aes_stg_v4(float *p, const float (&v)[4]) { *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]); }
static inline float4 aes_lds_v4(const float *p) { return *reinterpret_cast<const float4 *>(p); }
static inline float4 aes_ldg_v4(const float *p) { return *reinterpret_cast<const float4 *>(p); }
#endif

// ---- FR-wide vector access (p aligned to 4*FR bytes) -----------------------------------
// SP: 0 = shared memory, 1 = global memory, 2 = unknown (generic load)
template <int FR, int SP> __device__ __forceinline__ void aes_ldv_sp(const float *p, float (&o)[FR])
{
    if (FR == 4 && SP != 2) {
        const float4 t = SP == 0 ? aes_lds_v4(p) : aes_ldg_v4(p);
        o[0] = t.x; o[1 % FR] = t.y; o[2 % FR] = t.z; o[3 % FR] = t.w;
    } else if (FR == 4) { const float4 t = *reinterpret_cast<const float4 *>(p); o[0] = t.x; o[1 % FR] = t.y; o[2 % FR] = t.z; o[3 % FR] = t.w; }
    else if (FR == 2) { const float2 t = *reinterpret_cast<const float2 *>(p); o[0] = t.x; o[1 % FR] = t.y; }
    else o[0] = p[0];
}
template <int FR> __device__ __forceinline__ void aes_ldv(const float *p, float (&o)[FR])
{
    if (FR == 4) { const float4 t = *reinterpret_cast<const float4 *>(p); o[0] = t.x; o[1] = t.y; o[2 % FR] = t.z; o[3 % FR] = t.w; }
    else if (FR == 2) { const float2 t = *reinterpret_cast<const float2 *>(p); o[0] = t.x; o[1 % FR] = t.y; }
    else o[0] = p[0];
}
template <int FR> __device__ __forceinline__ void aes_stv(float *p, const float (&o)[FR])
{
    if (FR == 4) *reinterpret_cast<float4 *>(p) = make_float4(o[0], o[1], o[2 % FR], o[3 % FR]);
    else if (FR == 2) *reinterpret_cast<float2 *>(p) = make_float2(o[0], o[1 % FR]);
    else p[0] = o[0];
}

// FR consecutive ring elements starting at r (0 <= r < len, len a multiple of 4).  The
// misalignment m = r mod FR is the same for every thread (it is -lag mod FR), so the
// branch is uniform: one aligned vector load when m == 0, two plus a static select else.
template <int FR> __device__ __forceinline__ void aes_ring_read(const float *rb, int r, int len, float (&o)[FR])
{
    const int m = r & (FR - 1);
    const int a0 = r - m;
    float A[FR];
    aes_ldv<FR>(rb + a0, A);
    if (m == 0) {
#pragma unroll
        for (int j = 0; j < FR; ++j) o[j] = A[j];
        return;
    }
    int b0 = a0 + FR;
    if (b0 >= len) b0 -= len;
    float Bv[FR];
    aes_ldv<FR>(rb + b0, Bv);
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

// first read / write slot of this thread in a REG ring whose tile-start slot is `wpos`
__device__ __forceinline__ int aes_rslot(int wpos, int i0, const DevRing &rg)
{
    int r = wpos + i0 - rg.lag;            // > -len because i0 >= 0, wpos >= 0, lag <= len
    if (r < 0) r += rg.len;
    if (r >= rg.len) r -= rg.len;          // only a REGB line (lag < T) can land here: wpos + i0 < len + T
    return r;
}
__device__ __forceinline__ int aes_wslot(int wpos, int i0, const DevRing &rg)
{
    int w = wpos + i0;
    if (w >= rg.len) w -= rg.len;
    return w;
}

// ---- tile I/O: FR frames x 2 channels per thread, registers <-> HBM ----------------------
template <int FR>
__device__ __forceinline__ void aes_load_frames(const ChainArgs &a, long long b, long long n0, int len,
                                                int tid, float (&x)[2][FR])
{
    const int i0 = FR * tid;
    const long long base = b * a.N + n0 + i0;
    if (a.in_fmt == AESK_F32_STEREO && FR >= 2 && i0 + FR <= len && ((b * a.N) & 1) == 0) {
        const float4 *p = reinterpret_cast<const float4 *>(reinterpret_cast<const float *>(a.x) + 2 * base);
#pragma unroll
        for (int j = 0; j < FR / 2; ++j) {
            const float4 t = __ldcs(p + j);
            x[0][2 * j] = t.x; x[1][2 * j] = t.y; x[0][(2 * j + 1) % FR] = t.z; x[1][(2 * j + 1) % FR] = t.w;
        }
        return;
    }
#pragma unroll
    for (int j = 0; j < FR; ++j) {
        float l = 0.f, r = 0.f;
        if (i0 + j < len) {
            if (a.in_fmt == AESK_F32_STEREO) {
                const float2 t = __ldcs(reinterpret_cast<const float2 *>(a.x) + base + j);
                l = t.x; r = t.y;
            } else if (a.in_fmt == AESK_F32_MONO) {
                l = r = __ldcs(reinterpret_cast<const float *>(a.x) + base + j);     // core.py:147-149
            } else {
                // engine.py:78-84: int16 -> /32768 -> mean over channels (exact in f32)
                const short2 t = reinterpret_cast<const short2 *>(a.x)[base + j];
                l = r = (float)((int)t.x + (int)t.y) * (1.0f / 65536.0f);
            }
        }
        x[0][j] = l; x[1][j] = r;
    }
}

template <int FR>
__device__ __forceinline__ void aes_store_frames(const ChainArgs &a, long long b, long long n0, int len,
                                                 int tid, const float (&v)[2][FR])
{
    const int i0 = FR * tid;
    const long long base = b * a.N + n0 + i0;
    if (a.out_fmt == AESK_F32_STEREO && FR >= 2 && i0 + FR <= len && ((b * a.N) & 1) == 0) {
        float4 *p = reinterpret_cast<float4 *>(reinterpret_cast<float *>(a.y) + 2 * base);
#pragma unroll
        for (int j = 0; j < FR / 2; ++j)
            __stcs(p + j, make_float4(v[0][2 * j], v[1][2 * j], v[0][(2 * j + 1) % FR], v[1][(2 * j + 1) % FR]));
        return;
    }
#pragma unroll
    for (int j = 0; j < FR; ++j) {
        if (i0 + j < len) {
            if (a.out_fmt == AESK_F32_STEREO) {
                __stcs(reinterpret_cast<float2 *>(a.y) + base + j, make_float2(v[0][j], v[1][j]));
            } else {
                // engine.py:104-105: clip, *32767, astype(int16) truncates toward zero
                const short ql = (short)__float2int_rz(__fmul_rn(aes_clip1(v[0][j]), 32767.0f));
                const short qr = (short)__float2int_rz(__fmul_rn(aes_clip1(v[1][j]), 32767.0f));
                reinterpret_cast<short2 *>(a.y)[base + j] = make_short2(ql, qr);
            }
        }
    }
}

// registers <-> the planar tile buffer in shared memory (own entries only)
template <int FR> __device__ __forceinline__ void aes_spill(const KCtx &c, const float (&v)[2][FR])
{
    constexpr int T = AES_NT * FR;
    aes_stv<FR>(c.tile + FR * c.tid, v[0]);
    aes_stv<FR>(c.tile + T + FR * c.tid, v[1]);
}
template <int FR> __device__ __forceinline__ void aes_reload(const KCtx &c, float (&v)[2][FR])
{
    constexpr int T = AES_NT * FR;
    aes_ldv<FR>(c.tile + FR * c.tid, v[0]);
    aes_ldv<FR>(c.tile + T + FR * c.tid, v[1]);
}

// ---- phase walk over the tile in shared memory (any lag >= 1) ---------------------------
// Threads 0..127 take channel 0, 128..255 channel 1; thread j walks samples j, j+L, j+2L..
// of the tile with the line value in a register.  OP 0: feedback delay + mix/clip
// (delay.py:7-22,94-96); OP 1: pure delay (reverb.py:11-31); OP 2: all-pass (reverb.py:48-67).
template <int FR, int OP>
__device__ void aes_walk(const KCtx &c, int ring_id, float p0, float p1, float p2)
{
    constexpr int T = AES_NT * FR;
    const int ch = c.tid >> 7, j0 = c.tid & 127;
    const DevRing rg = c.P->ring[ring_id];
    float *rb = aes_ring_base(c, rg);
    const int L = rg.len, pos = c.rpos[ring_id];
    const int W = L < T ? L : T;
    float *s = c.tile + ch * T;
    for (int j = j0; j < W && j < c.len; j += 128) {
        int slot = pos + j;
        if (slot >= L) slot -= L;
        float line = (c.n0 + j >= L) ? rb[slot] : 0.0f;       // zero history on a fresh clip
        for (int i = j; i < c.len; i += 4 * L) {
            float xs[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {                       // loads first: the walk is a serial chain
                const long long ii = (long long)i + (long long)u * L;
                xs[u] = ii < c.len ? s[ii] : 0.0f;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const long long ii = (long long)i + (long long)u * L;
                if (ii < c.len) {
                    const float x = xs[u];
                    if (OP == 0) {                              // buf[n] = x + fb*buf[n-L]; out = clip(dry*x + wet*buf[n-L])
                        s[ii] = aes_mix_clip(p1, x, p2, line);
                        line = fmaf(line, p0, x);
                    } else if (OP == 1) {                       // y[n] = x[n-L]
                        s[ii] = line;
                        line = x;
                    } else {                                    // y = d - a*x ; buf = x + a*y
                        const float yo = fmaf(-p0, x, line);
                        s[ii] = yo;
                        line = fmaf(p0, yo, x);
                    }
                }
            }
        }
        rb[slot] = line;
    }
}

// ---- feedback delay, lag >= T: elementwise on registers ------------------------------------
template <int FR>
__device__ __forceinline__ void aes_delay_fetch(const DevStage &st, const KCtx &c, int tile_ahead, float (&ln)[2][FR])
{
    const int i0 = FR * c.tid;
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {
        const int rid = st.ring[ch][0];
        const DevRing rg = c.P->ring[rid];
        int wpos = c.rpos[rid];
        if (tile_ahead) { wpos += rg.tinc; if (wpos >= rg.len) wpos -= rg.len; }
        aes_ring_read<FR>(aes_ring_base(c, rg), aes_rslot(wpos, i0, rg), rg.len, ln[ch]);
    }
}

template <int FR>
__device__ __forceinline__ void aes_stage_delay_reg(const DevStage &st, const KCtx &c, float (&v)[2][FR],
                                                    const float (&ln)[2][FR])
{
    const int i0 = FR * c.tid;
    const float fb = st.fb, dry = st.dry, wet = st.wet;
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {
        const int rid = st.ring[ch][0];
        const DevRing rg = c.P->ring[rid];
        float nb[FR];
#pragma unroll
        for (int j = 0; j < FR; ++j) {
            const float line = (c.n0 + i0 + j >= rg.lag) ? ln[ch][j] : 0.0f;   // zero history on a fresh clip
            const float x = v[ch][j];
            nb[j] = fmaf(line, fb, x);                           // buf[n] = x + buf[n-L]*fb
            v[ch][j] = aes_mix_clip(dry, x, wet, line);
        }
        aes_stv<FR>(aes_ring_base(c, rg) + aes_wslot(c.rpos[rid], i0, rg), nb);
    }
}

// ---- Schroeder/Moorer reverb (reverb.py:208-277) ---------------------------------------------
template <int FR>
__device__ void aes_stage_reverb(const DevStage &st, const KCtx &c, float (&v)[2][FR], const double *sin,
                                 double *sout)
{
    const int i0 = FR * c.tid, lane = c.lane, warp = c.warp;
    float pre[2][FR];

    // pre-delay (reverb.py:11-31)
    if (st.pre_ring[0] < 0) {
#pragma unroll
        for (int ch = 0; ch < 2; ++ch)
#pragma unroll
            for (int j = 0; j < FR; ++j) pre[ch][j] = v[ch][j];
    } else if (st.mode == AES_MODE_REG) {
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
            const int rid = st.pre_ring[ch];
            const DevRing rg = c.P->ring[rid];
            float *rb = aes_ring_base(c, rg);
            aes_ring_read<FR>(rb, aes_rslot(c.rpos[rid], i0, rg), rg.len, pre[ch]);
#pragma unroll
            for (int j = 0; j < FR; ++j)
                if (c.n0 + i0 + j < rg.lag) pre[ch][j] = 0.0f;
            aes_stv<FR>(rb + aes_wslot(c.rpos[rid], i0, rg), v[ch]);
        }
    } else {                                    // AES_MODE_REGB: lag < T, so part of what is read is written in this tile
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
            const int rid = st.pre_ring[ch];
            const DevRing rg = c.P->ring[rid];
            aes_stv<FR>(aes_ring_base(c, rg) + aes_wslot(c.rpos[rid], i0, rg), v[ch]);
        }
        __syncthreads();
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
            const int rid = st.pre_ring[ch];
            const DevRing rg = c.P->ring[rid];
            aes_ring_read<FR>(aes_ring_base(c, rg), aes_rslot(c.rpos[rid], i0, rg), rg.len, pre[ch]);
#pragma unroll
            for (int j = 0; j < FR; ++j)
                if (c.n0 + i0 + j < rg.lag) pre[ch][j] = 0.0f;
        }
    }

    // damped combs (reverb.py:33-46), 4 per side at a time; sum in comb order in f32
    float sum[2][FR];
#pragma unroll
    for (int ch = 0; ch < 2; ++ch)
#pragma unroll
        for (int j = 0; j < FR; ++j) sum[ch][j] = 0.0f;
    const float h = st.h, omh = st.omh, hl = st.hlane[lane], hw = st.hp[5];
    const int nc = st.nc, nscan = st.nscan;
    float *wt = reinterpret_cast<float *>(c.wtot);              // [8 warps][2 ch][4 combs]
    for (int gi = 0; gi < nc; gi += 4) {
        float y[2][4][FR], e[2][4];
#pragma unroll
        for (int ch = 0; ch < 2; ++ch)
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                e[ch][cc] = 0.0f;
                if (gi + cc < nc) {
                    const int rid = st.ring[ch][gi + cc];
                    const DevRing rg = c.P->ring[rid];
                    aes_ring_read<FR>(c.rings + rg.off, aes_rslot(c.rpos[rid], i0, rg), rg.len, y[ch][cc]);
                    float lp = 0.0f;
#pragma unroll
                    for (int j = 0; j < FR; ++j) lp = fmaf(h, lp, omh * y[ch][cc][j]);   // zero-state one-pole
                    e[ch][cc] = lp;
                }
            }
        // inclusive warp scan of the chunk end values, multiplier h^(FR*2^s)
        for (int s = 0; s < nscan; ++s) {
            const float m = st.hp[s];
#pragma unroll
            for (int ch = 0; ch < 2; ++ch)
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) {
                    const float u = __shfl_up_sync(0xffffffffu, e[ch][cc], 1 << s);
                    if (lane >= (1 << s)) e[ch][cc] = fmaf(m, u, e[ch][cc]);
                }
        }
        if (lane == 31) {
#pragma unroll
            for (int ch = 0; ch < 2; ++ch)
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) wt[(warp * 2 + ch) * 4 + cc] = e[ch][cc];
        }
        __syncthreads();
        const int u0 = warp > st.nxw ? warp - st.nxw : 0;       // older warps' carries are below 2^-32
#pragma unroll
        for (int ch = 0; ch < 2; ++ch)
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                const float ex = __shfl_up_sync(0xffffffffu, e[ch][cc], 1);
                if (gi + cc < nc) {
                    float C = u0 == 0 ? (float)sin[ch * 8 + gi + cc] : 0.0f;     // lp at the end of the previous tile
                    for (int u = u0; u < warp; ++u) C = fmaf(hw, C, wt[(u * 2 + ch) * 4 + cc]);
                    float lp = fmaf(hl, C, lane == 0 ? 0.0f : ex);               // lp just before this thread's chunk
                    const int rid = st.ring[ch][gi + cc];
                    const DevRing rg = c.P->ring[rid];
                    const float g = st.g[ch][gi + cc];
                    float nb[FR];
#pragma unroll
                    for (int j = 0; j < FR; ++j) {
                        lp = fmaf(h, lp, omh * y[ch][cc][j]);                    // damped = (1-h)*y + h*lp
                        nb[j] = fmaf(g, lp, pre[ch][j]);                         // buf[n] = x + g*damped
                        sum[ch][j] = __fadd_rn(sum[ch][j], y[ch][cc][j]);        // comb output = delayed sample
                    }
                    aes_stv<FR>(c.rings + rg.off + aes_wslot(c.rpos[rid], i0, rg), nb);
                    if (c.tid == AES_NT - 1) sout[ch * 8 + gi + cc] = (double)lp;
                }
            }
        if (gi + 4 < nc) __syncthreads();                       // wt is reused by the next group
    }

    // series all-passes (reverb.py:48-67): phase walk on the tile
    aes_spill<FR>(c, sum);
    __syncthreads();
    for (int k = 0; k < st.na; ++k) {
        aes_walk<FR, 2>(c, st.apring[c.tid >> 7][k], st.a, 0.f, 0.f);
        __syncthreads();
    }
    aes_reload<FR>(c, sum);

    // mix + clip (reverb.py:275-277)
    const float dry = st.dry, wet = st.wet;
#pragma unroll
    for (int ch = 0; ch < 2; ++ch)
#pragma unroll
        for (int j = 0; j < FR; ++j) v[ch][j] = aes_mix_clip(dry, v[ch][j], wet, sum[ch][j]);
}

// ---- biquad, Direct Form I in f64 (filter.py:8-40) ---------------------------------------------
template <int FR>
__device__ void aes_stage_biquad(const DevStage &st, const KCtx &c, float (&v)[2][FR], const double *sin,
                                 double *sout)
{
    constexpr int T = AES_NT * FR;
    const int i0 = FR * c.tid, lane = c.lane, warp = c.warp;
    const double b0 = st.bq[0], b1 = st.bq[1], b2 = st.bq[2], a1 = st.bq[3], a2 = st.bq[4];
    aes_spill<FR>(c, v);                                        // neighbours' x[n-1], x[n-2]
    __syncthreads();
    double xm1[2], xm2[2], e1[2], e2[2];
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {
        const float *xc = c.tile + ch * T;
        xm1[ch] = i0 >= 1 ? (double)xc[i0 - 1] : sin[4 * ch + 0];
        xm2[ch] = i0 >= 2 ? (double)xc[i0 - 2] : (i0 == 1 ? sin[4 * ch + 0] : sin[4 * ch + 1]);
        // zero-state response of the chunk: end state (y[FR-1], y[FR-2])
        double y1 = 0.0, y2 = 0.0, p1 = xm1[ch], p2 = xm2[ch];
#pragma unroll
        for (int j = 0; j < FR; ++j) {
            const double xj = (double)v[ch][j];
            const double y = b0 * xj + b1 * p1 + b2 * p2 - a1 * y1 - a2 * y2;
            y2 = y1; y1 = y; p2 = p1; p1 = xj;
        }
        e1[ch] = y1; e2[ch] = y2;
    }
    for (int s = 0; s < 5; ++s) {
        const double m0 = st.bq_pow[s][0], m1 = st.bq_pow[s][1], m2 = st.bq_pow[s][2], m3 = st.bq_pow[s][3];
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
            const double u1 = __shfl_up_sync(0xffffffffu, e1[ch], 1 << s);
            const double u2 = __shfl_up_sync(0xffffffffu, e2[ch], 1 << s);
            if (lane >= (1 << s)) {
                e1[ch] += m0 * u1 + m1 * u2;
                e2[ch] += m2 * u1 + m3 * u2;
            }
        }
    }
    if (lane == 31) {
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
            c.wtot[(warp * 2 + ch) * 2] = e1[ch];
            c.wtot[(warp * 2 + ch) * 2 + 1] = e2[ch];
        }
    }
    __syncthreads();
    const double w0 = st.bq_pow[5][0], w1 = st.bq_pow[5][1], w2 = st.bq_pow[5][2], w3 = st.bq_pow[5][3];
    const double l0 = st.bq_lane[lane][0], l1 = st.bq_lane[lane][1], l2 = st.bq_lane[lane][2], l3 = st.bq_lane[lane][3];
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {
        double C1 = sin[4 * ch + 2], C2 = sin[4 * ch + 3];       // (y[n0-1], y[n0-2])
        for (int u = 0; u < warp; ++u) {
            const double t1 = w0 * C1 + w1 * C2 + c.wtot[(u * 2 + ch) * 2];
            const double t2 = w2 * C1 + w3 * C2 + c.wtot[(u * 2 + ch) * 2 + 1];
            C1 = t1; C2 = t2;
        }
        double x1 = __shfl_up_sync(0xffffffffu, e1[ch], 1), x2 = __shfl_up_sync(0xffffffffu, e2[ch], 1);
        if (lane == 0) { x1 = 0.0; x2 = 0.0; }
        double y1 = x1 + l0 * C1 + l1 * C2;
        double y2 = x2 + l2 * C1 + l3 * C2;
        double p1 = xm1[ch], p2 = xm2[ch];
#pragma unroll
        for (int j = 0; j < FR; ++j) {
            const double xj = (double)v[ch][j];
            const double y = b0 * xj + b1 * p1 + b2 * p2 - a1 * y1 - a2 * y2;
            v[ch][j] = (float)y;
            y2 = y1; y1 = y; p2 = p1; p1 = xj;
            if (i0 + j == c.len - 1) {                           // DF-I state after the tile's last frame
                sout[4 * ch + 0] = p1; sout[4 * ch + 1] = p2;
                sout[4 * ch + 2] = y1; sout[4 * ch + 3] = y2;
            }
        }
    }
    __syncthreads();                                             // wtot / tile are free again
}

// ---- noise gate (gate.py:6-42) -------------------------------------------------------------------
template <int FR>
__device__ void aes_stage_gate(const DevStage &st, const KCtx &c, float (&v)[2][FR], const double *sin, double *sout)
{
    const int i0 = FR * c.tid, lane = c.lane, w = c.warp;
    const double thr = st.thr, ka = 1.0 - st.att, kr = 1.0 - st.rel, att = st.att;
    bool open[FR];
    double A = 1.0, Bv = 0.0;
#pragma unroll
    for (int f = 0; f < FR; ++f) {
        const float lvl = fmaxf(fabsf(v[0][f]), fabsf(v[1][f]));     // stereo-linked level
        open[f] = (double)lvl > thr;
        const double am = open[f] ? ka : kr, bm = open[f] ? att : 0.0;
        Bv = am * Bv + bm;                                       // compose g -> am*g + bm
        A = am * A;
    }
    for (int s = 0; s < 5; ++s) {
        const double Au = __shfl_up_sync(0xffffffffu, A, 1 << s);
        const double Bu = __shfl_up_sync(0xffffffffu, Bv, 1 << s);
        if (lane >= (1 << s)) { Bv = A * Bu + Bv; A = A * Au; }
    }
    if (lane == 31) { c.wtot[2 * w] = A; c.wtot[2 * w + 1] = Bv; }
    __syncthreads();
    double g = sin[0];
    for (int u = 0; u < w; ++u) g = c.wtot[2 * u] * g + c.wtot[2 * u + 1];
    double Ae = __shfl_up_sync(0xffffffffu, A, 1), Be = __shfl_up_sync(0xffffffffu, Bv, 1);
    if (lane == 0) { Ae = 1.0; Be = 0.0; }
    g = Ae * g + Be;
#pragma unroll
    for (int f = 0; f < FR; ++f) {
        g = (open[f] ? ka : kr) * g + (open[f] ? att : 0.0);
        const float gf = (float)g;
        v[0][f] *= gf;
        v[1][f] *= gf;
        if (i0 + f == c.len - 1) sout[0] = g;                    // gain after the tile's last frame
    }
    __syncthreads();                                             // wtot is free again
}

// ---- octaver (octaver.py:17-82 + wrapper 116-150) ---------------------------------------------------
__device__ __forceinline__ float aes_hermite(float t, float y0, float y1, float y2, float y3)
{
    const float c1 = 0.5f * (y2 - y0);
    const float c2 = y0 - 2.5f * y1 + 2.0f * y2 - 0.5f * y3;
    const float c3 = 0.5f * (y3 - y0) + 1.5f * (y1 - y2);
    return ((c3 * t + c2) * t + c1) * t + y1;
}

__device__ __forceinline__ float aes_octaver_tap(const float *rb, int mask, long long n, int size,
                                                 double fsize, double p)
{
    // read position raw = w - p*size + size (octaver.py:39); relative to the write
    // pointer it is size*(1-p) in (0, size]; ring slot (w - d) holds mono[n - d].
    const double raw = fsize - p * fsize;
    const int m = (int)raw;
    const float frac = (float)(raw - (double)m);
    const int d0 = size - m + 1;                                 // taps k=-1..2 -> d0, d0-1, d0-2, d0-3 (mod size)
    float t[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        int d = d0 - k;
        if (d < 0) d += size;
        if (d >= size) d -= size;
        t[k] = rb[(int)((n - d) & mask)];
    }
    return aes_hermite(frac, t[0], t[1], t[2], t[3]);
}

template <int FR>
__device__ void aes_stage_octaver(const DevStage &st, const KCtx &c, float (&v)[2][FR])
{
    const int i0 = FR * c.tid;
    const DevRing rg = c.P->ring[st.ring[0][0]];
    float *rb = c.rings + rg.off;
    const int mask = st.oct_mask, size = st.oct_size;
    float mono[FR];
#pragma unroll
    for (int j = 0; j < FR; ++j) mono[j] = __fmul_rn(__fadd_rn(v[0][j], v[1][j]), 0.5f);   // np.mean, octaver.py:124-126
    aes_stv<FR>(rb + (int)((c.n0 + i0) & mask), mono);
    __syncthreads();
    const float wet_g = st.mix, dry_g = (float)(1.0 - (double)st.mix);
    const double ph0 = st.ph0, step = st.step, fsize = st.fsize;
#pragma unroll
    for (int j = 0; j < FR; ++j) {
        const long long n = c.n0 + i0 + j;
        double ph = ph0 + (double)n * step;                      // phasor in closed form
        ph -= floor(ph);
        double p2 = ph + 0.5;
        if (p2 >= 1.0) p2 -= 1.0;
        const float s1 = aes_octaver_tap(rb, mask, n, size, fsize, ph);
        const float s2 = aes_octaver_tap(rb, mask, n, size, fsize, p2);
        const float sn = sinpif((float)ph);
        const float g1 = sn * sn;                                // 0.5*(1-cos(2*pi*p))
        const float g2 = 1.0f - g1;                              // p2 = p + 1/2
        const float wet = __fmul_rn(s1 * g1 + s2 * g2, wet_g);
        v[0][j] = __fadd_rn(__fmul_rn(v[0][j], dry_g), wet);
        v[1][j] = __fadd_rn(__fmul_rn(v[1][j], dry_g), wet);
    }
}

// ---- distortion (our definition, no reference block) ---------------------------------------------
template <int FR>
__device__ __forceinline__ void aes_stage_distortion(const DevStage &st, float (&v)[2][FR])
{
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

// ---- the kernel body ------------------------------------------------------------------------------
template <int FR>
__device__ void aes_chain_body(const ChainArgs &a)
{
    constexpr int T = AES_NT * FR;
    AES_DYN_SMEM(float, smem);
    const DevPlan *P = a.plan;
    KCtx c;
    c.P = P;
    c.tid = threadIdx.x;
    c.lane = c.tid & 31;
    c.warp = c.tid >> 5;
    c.tile = smem;
    c.rings = smem + 2 * T;
    const int foff = (2 * T + P->smem_floats + 3) & ~3;
    c.wtot = reinterpret_cast<double *>(smem + foff);
    double *state = c.wtot + 64;
    const int nstate = P->n_state;
    int *rpos2 = reinterpret_cast<int *>(state + 2 * nstate);   // [2][n_rings], ping-pong by tile parity
    c.gscr = a.scratch + (long long)blockIdx.x * P->scratch_floats;
    const int nst = P->n_stages, nr = P->n_rings, pf_stage = P->pf_stage;

    for (long long b = blockIdx.x; b < a.B; b += gridDim.x) {
        // fresh block state for every clip (the chain's re-prepare, core.py:123-129)
        for (int i = c.tid; i < P->smem_floats; i += AES_NT) c.rings[i] = 0.0f;
        for (int i = c.tid; i < nr; i += AES_NT) rpos2[i] = 0;
        c.rpos = rpos2;
        for (int i = c.tid; i < nstate; i += AES_NT) state[i] = P->stage[i >> 4].init[i & 15];
        __syncthreads();

        float xn[2][FR], pf[2][FR];
#pragma unroll
        for (int ch = 0; ch < 2; ++ch)
#pragma unroll
            for (int j = 0; j < FR; ++j) pf[ch][j] = 0.0f;
        c.n0 = 0;
        c.len = a.N < (long long)T ? (int)a.N : T;
        aes_load_frames<FR>(a, b, 0, c.len, c.tid, xn);
        if (pf_stage >= 0) aes_delay_fetch<FR>(P->stage[pf_stage], c, 0, pf);
        int par = 0;
        for (long long n0 = 0; n0 < a.N; n0 += T, par ^= 1) {
            c.n0 = n0;
            c.len = (a.N - n0 < (long long)T) ? (int)(a.N - n0) : T;
            c.rpos = rpos2 + par * nr;
            float v[2][FR], ln[2][FR];
#pragma unroll
            for (int ch = 0; ch < 2; ++ch)
#pragma unroll
                for (int j = 0; j < FR; ++j) { v[ch][j] = xn[ch][j]; ln[ch][j] = pf[ch][j]; }
            // software pipeline: next tile's frames and the long delay line it needs
            const bool more = n0 + T < a.N;
            if (more) {
                const long long rem = a.N - n0 - T;
                aes_load_frames<FR>(a, b, n0 + T, rem < (long long)T ? (int)rem : T, c.tid, xn);
                if (pf_stage >= 0) aes_delay_fetch<FR>(P->stage[pf_stage], c, 1, pf);
            }
            for (int s = 0; s < nst; ++s) {
                const DevStage &st = P->stage[s];
                const double *sin = state + par * nstate + 16 * s;
                double *sout = state + (par ^ 1) * nstate + 16 * s;
                switch (st.kind) {
                case AESK_DELAY:
                    if (st.mode == AES_MODE_REG) {
                        if (s != pf_stage) {                     // fetched on the spot, into its own registers:
                            float ln_now[2][FR];                 // `ln` holds the prefetched stage's line
                            aes_delay_fetch<FR>(st, c, 0, ln_now);
                            aes_stage_delay_reg<FR>(st, c, v, ln_now);
                        } else {
                            aes_stage_delay_reg<FR>(st, c, v, ln);
                        }
                    } else {
                        aes_spill<FR>(c, v);
                        __syncthreads();
                        aes_walk<FR, 0>(c, st.ring[c.tid >> 7][0], st.fb, st.dry, st.wet);
                        __syncthreads();
                        aes_reload<FR>(c, v);
                    }
                    break;
                case AESK_REVERB:     aes_stage_reverb<FR>(st, c, v, sin, sout); break;
                case AESK_BIQUAD:     aes_stage_biquad<FR>(st, c, v, sin, sout); break;
                case AESK_GATE:       aes_stage_gate<FR>(st, c, v, sin, sout); break;
                case AESK_OCTAVER:    aes_stage_octaver<FR>(st, c, v); break;
                case AESK_DISTORTION: aes_stage_distortion<FR>(st, v); break;
                default: break;
                }
            }
            aes_store_frames<FR>(a, b, n0, c.len, c.tid, v);
            if (c.tid < nr) {                                    // next tile's slots go to the other parity:
                const DevRing rg = P->ring[c.tid];               // slower threads may still read this tile's
                int p = c.rpos[c.tid] + rg.tinc;
                if (p >= rg.len) p -= rg.len;
                rpos2[(par ^ 1) * nr + c.tid] = p;
            }
            __syncthreads();
        }
        if (a.state_out != nullptr)
            for (int i = c.tid; i < nstate; i += AES_NT) a.state_out[b * nstate + i] = state[par * nstate + i];
    }
}

#ifndef AES_CPU_EMU
template <int FR>
__global__ void __launch_bounds__(AES_NT, 2) aes_chain_kernel(const ChainArgs a)
{
    aes_chain_body<FR>(a);
}
#endif
