// aes_chain_kernel.cuh -- the fused effect-chain kernel (device code).
//
// One CTA owns one clip at a time (persistent grid, clips strided over CTAs) and
// walks it in time tiles of T = 128*K stereo frames.  A tile is read from HBM
// once, pushed through every block of the chain while it sits in shared memory,
// and written once: 8 algorithmic bytes per output sample whatever the chain
// length.  All recurrences of the reference are kept exactly, only
// re-associated so that a tile is parallel:
//
//   * lag-L feedback lines (delay.py:7-22, reverb.py:48-67 all-pass, pre-delay
//     reverb.py:11-31): thread j walks the samples j, j+L, j+2L.. of the tile with
//     the line value in a register ("phase walk"); the ring (length exactly L,
//     slot = n mod L) is touched once per thread per tile.  L >= T degenerates to
//     one read-modify-write per sample, fully parallel.
//   * damped comb (reverb.py:33-46): lag L >= T, so every delayed sample of the
//     tile is already in the ring; the one-pole that runs along time inside the
//     feedback path is a constant-coefficient first-order scan: K samples per
//     thread serially, warp-shuffle Kogge-Stone across lanes (truncated when
//     h^(K*2^s) < 2^-32), 4-entry carry chain across warps through shared memory.
//   * biquad DF-I (filter.py:8-40): same shape with 2x2 companion-matrix powers,
//     everything in f64 like the reference's promoted arithmetic.
//   * gate (gate.py:6-42): the branch depends only on the input (gain stays in
//     [0,1], target in {0,1}), so the gain is an affine scan with per-sample
//     coefficients; f64.
//   * octaver (octaver.py:17-82): no feedback; a gather from the mono history with
//     the phasor in closed form, frac(ph0 + n*step).
//
// The file is also compiled by g++ against tests/cpu_emu/cuda_emu.h (AES_CPU_EMU)
// so the logic can be exercised without a GPU; the product build is nvcc only.
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

struct TileCtx {
    const DevPlan *P;
    float *cur, *aux, *rings, *gscr;
    double *wtot;
    int *rpos;
    long long n0;
    int len, tid;
};

__device__ __forceinline__ float aes_clip1(float v) { return fminf(fmaxf(v, -1.0f), 1.0f); }

// numpy evaluates dry*x + wet*w as three separately rounded f32 ufuncs
// (delay.py:94-96, reverb.py:275-277): keep the products and the sum unfused.
__device__ __forceinline__ float aes_mix_clip(float dry, float x, float wet, float w)
{
    return aes_clip1(__fadd_rn(__fmul_rn(dry, x), __fmul_rn(wet, w)));
}

__device__ __forceinline__ float *aes_ring_base(const TileCtx &c, const DevRing &r)
{
    return (r.space == AES_SPACE_GLOBAL ? c.gscr : c.rings) + r.off;
}

// ---- tile I/O ------------------------------------------------------------------
template <int K>
__device__ __forceinline__ void aes_load_tile(const ChainArgs &a, long long b, const TileCtx &c)
{
    constexpr int T = 128 * K;
    const long long base = b * a.N + c.n0;
#pragma unroll
    for (int m = 0; m < T / AES_NT; ++m) {
        const int i = c.tid + AES_NT * m;
        float l = 0.f, r = 0.f;
        if (i < c.len) {
            if (a.in_fmt == AESK_F32_STEREO) {
                const float2 v = reinterpret_cast<const float2 *>(a.x)[base + i];
                l = v.x; r = v.y;
            } else if (a.in_fmt == AESK_F32_MONO) {
                l = r = reinterpret_cast<const float *>(a.x)[base + i];   // core.py:147-149
            } else {
                // engine.py:78-84: int16 -> /32768 -> mean over channels (exact in f32)
                const short2 v = reinterpret_cast<const short2 *>(a.x)[base + i];
                l = r = (float)((int)v.x + (int)v.y) * (1.0f / 65536.0f);
            }
        }
        c.cur[i] = l;
        c.cur[T + i] = r;
    }
}

template <int K>
__device__ __forceinline__ void aes_store_tile(const ChainArgs &a, long long b, const TileCtx &c)
{
    constexpr int T = 128 * K;
    const long long base = b * a.N + c.n0;
#pragma unroll
    for (int m = 0; m < T / AES_NT; ++m) {
        const int i = c.tid + AES_NT * m;
        if (i < c.len) {
            const float l = c.cur[i], r = c.cur[T + i];
            if (a.out_fmt == AESK_F32_STEREO) {
                reinterpret_cast<float2 *>(a.y)[base + i] = make_float2(l, r);
            } else {
                // engine.py:104-105: clip, *32767, astype(int16) truncates toward zero
                const short ql = (short)__float2int_rz(__fmul_rn(aes_clip1(l), 32767.0f));
                const short qr = (short)__float2int_rz(__fmul_rn(aes_clip1(r), 32767.0f));
                reinterpret_cast<short2 *>(a.y)[base + i] = make_short2(ql, qr);
            }
        }
    }
}

// ---- feedback delay (delay.py:7-22 + mix/clip delay.py:94-96), in place on cur ------
template <int K>
__device__ void aes_stage_delay(const DevStage &st, const TileCtx &c)
{
    constexpr int T = 128 * K;
    const int ch = c.tid >> 7, j0 = c.tid & 127;
    const DevRing rg = c.P->ring[st.ring[ch][0]];
    float *rb = aes_ring_base(c, rg);
    const int L = rg.len, pos = c.rpos[st.ring[ch][0]];
    const int W = L < T ? L : T;
    float *xc = c.cur + ch * T;
    const float fb = st.fb, dry = st.dry, wet = st.wet;
    for (int j = j0; j < W && j < c.len; j += 128) {
        int slot = pos + j;
        if (slot >= L) slot -= L;
        float line = (c.n0 + j >= L) ? rb[slot] : 0.0f;        // buf[n-L]; zero history on a fresh clip
        for (int i = j; i < c.len; i += L) {
            const float x = xc[i];
            const float nb = fmaf(line, fb, x);                // buf[n] = x + buf[n-L]*fb
            xc[i] = aes_mix_clip(dry, x, wet, line);
            line = nb;
        }
        rb[slot] = line;
    }
    __syncthreads();
}

// ---- Schroeder/Moorer reverb (reverb.py:208-277): cur -> cur, aux as scratch ----------
template <int K>
__device__ void aes_stage_reverb(const DevStage &st, const TileCtx &c, const double *sin, double *sout)
{
    constexpr int T = 128 * K;
    const int tid = c.tid;
    const int ch = tid >> 7, q = tid & 127, lane = tid & 31, wq = q >> 5;
    const float *in = c.cur;

    // pre-delay (reverb.py:11-31): pure shift into aux
    if (st.pre_ring[0] >= 0) {
        const DevRing rg = c.P->ring[st.pre_ring[ch]];
        float *rb = aes_ring_base(c, rg);
        const int L = rg.len, pos = c.rpos[st.pre_ring[ch]];
        const int W = L < T ? L : T;
        for (int j = q; j < W && j < c.len; j += 128) {
            int slot = pos + j;
            if (slot >= L) slot -= L;
            float line = (c.n0 + j >= L) ? rb[slot] : 0.0f;
            for (int i = j; i < c.len; i += L) {
                const float x = c.cur[ch * T + i];
                c.aux[ch * T + i] = line;
                line = x;
            }
            rb[slot] = line;
        }
        __syncthreads();
        in = c.aux;
    }

    // damped combs (reverb.py:33-46), 4 at a time; sum accumulated in comb order in f32
    const int i0 = q * K;
    float x[K], sum[K];
#pragma unroll
    for (int j = 0; j < K; ++j) { x[j] = in[ch * T + i0 + j]; sum[j] = 0.0f; }
    const float h = st.h, omh = st.omh, hl = st.hlane[lane];
    float *wt = reinterpret_cast<float *>(c.wtot);            // [2 ch][4 warps][4 combs]
    for (int gi = 0; gi < st.nc; gi += 4) {
        float y[4][K], e[4];
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
            e[cc] = 0.0f;
            if (gi + cc < st.nc) {
                const int rid = st.ring[ch][gi + cc];
                const DevRing rg = c.P->ring[rid];
                const float *rb = c.rings + rg.off;
                int idx = c.rpos[rid] + i0;                     // i0 < T <= len of every comb ring
                if (idx >= rg.len) idx -= rg.len;
                float lp = 0.0f;
#pragma unroll
                for (int j = 0; j < K; ++j) {
                    y[cc][j] = rb[idx];
                    if (++idx == rg.len) idx = 0;
                    lp = fmaf(h, lp, omh * y[cc][j]);           // zero-state one-pole over the chunk
                }
                e[cc] = lp;
            }
        }
        // inclusive warp scan of the chunk end values, multiplier h^(K*2^s)
        for (int s = 0; s < st.nscan; ++s) {
            const float m = st.hp[s];
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                const float v = __shfl_up_sync(0xffffffffu, e[cc], 1 << s);
                if (lane >= (1 << s)) e[cc] = fmaf(m, v, e[cc]);
            }
        }
        if (lane == 31) {
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) wt[(ch * 4 + wq) * 4 + cc] = e[cc];
        }
        __syncthreads();
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
            const float ex = __shfl_up_sync(0xffffffffu, e[cc], 1);
            if (gi + cc < st.nc) {
                float C = (float)sin[ch * 8 + gi + cc];         // lp at the end of the previous tile
                for (int w = 0; w < wq; ++w) C = fmaf(st.hp[5], C, wt[(ch * 4 + w) * 4 + cc]);
                float lp = fmaf(hl, C, lane == 0 ? 0.0f : ex);  // lp just before this thread's chunk
                const int rid = st.ring[ch][gi + cc];
                const DevRing rg = c.P->ring[rid];
                float *rb = c.rings + rg.off;
                const float g = st.g[ch][gi + cc];
                int idx = c.rpos[rid] + i0;
                if (idx >= rg.len) idx -= rg.len;
#pragma unroll
                for (int j = 0; j < K; ++j) {
                    lp = fmaf(h, lp, omh * y[cc][j]);           // damped = (1-h)*y + h*lp
                    rb[idx] = fmaf(g, lp, x[j]);                // buf[n] = x + g*damped
                    if (++idx == rg.len) idx = 0;
                    sum[j] = __fadd_rn(sum[j], y[cc][j]);       // comb output is the delayed sample
                }
                if (q == 127) sout[ch * 8 + gi + cc] = (double)lp;
            }
        }
        __syncthreads();                                        // wt is reused by the next group
    }
#pragma unroll
    for (int j = 0; j < K; ++j) c.aux[ch * T + i0 + j] = sum[j];
    __syncthreads();

    // series all-passes (reverb.py:48-67), in place on aux
    const float a = st.a;
    for (int k = 0; k < st.na; ++k) {
        const DevRing rg = c.P->ring[st.apring[ch][k]];
        float *rb = aes_ring_base(c, rg);
        const int L = rg.len, pos = c.rpos[st.apring[ch][k]];
        const int W = L < T ? L : T;
        float *s = c.aux + ch * T;
        for (int j = q; j < W && j < c.len; j += 128) {
            int slot = pos + j;
            if (slot >= L) slot -= L;
            float line = (c.n0 + j >= L) ? rb[slot] : 0.0f;
            for (int i = j; i < c.len; i += L) {
                const float xi = s[i];
                const float yo = fmaf(-a, xi, line);            // y = delayed - a*x
                s[i] = yo;
                line = fmaf(a, yo, xi);                         // buf = x + a*y
            }
            rb[slot] = line;
        }
        __syncthreads();
    }

    // mix + clip (reverb.py:275-277)
    const float dry = st.dry, wet = st.wet;
#pragma unroll
    for (int m = 0; m < 2 * T / AES_NT; ++m) {
        const int e = tid + AES_NT * m;
        c.cur[e] = aes_mix_clip(dry, c.cur[e], wet, c.aux[e]);
    }
    __syncthreads();
}

// ---- biquad, Direct Form I in f64 (filter.py:8-40), in place on cur ---------------------
template <int K>
__device__ void aes_stage_biquad(const DevStage &st, const TileCtx &c, const double *sin, double *sout)
{
    constexpr int T = 128 * K;
    const int tid = c.tid;
    const int ch = tid >> 7, q = tid & 127, lane = tid & 31, wq = q >> 5;
    const int i0 = q * K;
    float *xc = c.cur + ch * T;
    const double b0 = st.bq[0], b1 = st.bq[1], b2 = st.bq[2], a1 = st.bq[3], a2 = st.bq[4];
    double xs[K];
#pragma unroll
    for (int j = 0; j < K; ++j) xs[j] = (double)xc[i0 + j];
    double xm1, xm2;
    if (q == 0) { xm1 = sin[4 * ch + 0]; xm2 = sin[4 * ch + 1]; }
    else        { xm1 = (double)xc[i0 - 1]; xm2 = (double)xc[i0 - 2]; }

    // zero-state response of the chunk: end state (y[K-1], y[K-2])
    double y1 = 0.0, y2 = 0.0, p1 = xm1, p2 = xm2;
#pragma unroll
    for (int j = 0; j < K; ++j) {
        const double y = b0 * xs[j] + b1 * p1 + b2 * p2 - a1 * y1 - a2 * y2;
        y2 = y1; y1 = y; p2 = p1; p1 = xs[j];
    }
    double e1 = y1, e2 = y2;
    for (int s = 0; s < 5; ++s) {
        const double u1 = __shfl_up_sync(0xffffffffu, e1, 1 << s);
        const double u2 = __shfl_up_sync(0xffffffffu, e2, 1 << s);
        if (lane >= (1 << s)) {
            e1 += st.bq_pow[s][0] * u1 + st.bq_pow[s][1] * u2;
            e2 += st.bq_pow[s][2] * u1 + st.bq_pow[s][3] * u2;
        }
    }
    if (lane == 31) { c.wtot[(ch * 4 + wq) * 2] = e1; c.wtot[(ch * 4 + wq) * 2 + 1] = e2; }
    __syncthreads();
    double C1 = sin[4 * ch + 2], C2 = sin[4 * ch + 3];           // (y[n0-1], y[n0-2])
    for (int w = 0; w < wq; ++w) {
        const double t1 = st.bq_pow[5][0] * C1 + st.bq_pow[5][1] * C2 + c.wtot[(ch * 4 + w) * 2];
        const double t2 = st.bq_pow[5][2] * C1 + st.bq_pow[5][3] * C2 + c.wtot[(ch * 4 + w) * 2 + 1];
        C1 = t1; C2 = t2;
    }
    double x1 = __shfl_up_sync(0xffffffffu, e1, 1), x2 = __shfl_up_sync(0xffffffffu, e2, 1);
    if (lane == 0) { x1 = 0.0; x2 = 0.0; }
    y1 = x1 + st.bq_lane[lane][0] * C1 + st.bq_lane[lane][1] * C2;
    y2 = x2 + st.bq_lane[lane][2] * C1 + st.bq_lane[lane][3] * C2;
    p1 = xm1; p2 = xm2;
#pragma unroll
    for (int j = 0; j < K; ++j) {
        const double y = b0 * xs[j] + b1 * p1 + b2 * p2 - a1 * y1 - a2 * y2;
        xc[i0 + j] = (float)y;
        y2 = y1; y1 = y; p2 = p1; p1 = xs[j];
        if (i0 + j == c.len - 1) {                               // DF-I state after the tile's last frame
            sout[4 * ch + 0] = p1; sout[4 * ch + 1] = p2;
            sout[4 * ch + 2] = y1; sout[4 * ch + 3] = y2;
        }
    }
    __syncthreads();
}

// ---- noise gate (gate.py:6-42), in place on cur ---------------------------------------
template <int K>
__device__ void aes_stage_gate(const DevStage &st, const TileCtx &c, const double *sin, double *sout)
{
    constexpr int T = 128 * K;
    constexpr int F = T / AES_NT;                                // consecutive frames per thread
    const int tid = c.tid, lane = tid & 31, w = tid >> 5;
    const int i0 = tid * F;
    const double thr = st.thr, ka = 1.0 - st.att, kr = 1.0 - st.rel, att = st.att;
    bool open[F];
    double A = 1.0, Bv = 0.0;
#pragma unroll
    for (int f = 0; f < F; ++f) {
        const float lvl = fmaxf(fabsf(c.cur[i0 + f]), fabsf(c.cur[T + i0 + f]));   // stereo-linked level
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
    for (int v = 0; v < w; ++v) g = c.wtot[2 * v] * g + c.wtot[2 * v + 1];
    double Ae = __shfl_up_sync(0xffffffffu, A, 1), Be = __shfl_up_sync(0xffffffffu, Bv, 1);
    if (lane == 0) { Ae = 1.0; Be = 0.0; }
    g = Ae * g + Be;
#pragma unroll
    for (int f = 0; f < F; ++f) {
        g = (open[f] ? ka : kr) * g + (open[f] ? att : 0.0);
        const float gf = (float)g;
        c.cur[i0 + f] *= gf;
        c.cur[T + i0 + f] *= gf;
        if (i0 + f == c.len - 1) sout[0] = g;                    // gain after the tile's last frame
    }
    __syncthreads();
}

// ---- octaver (octaver.py:17-82 + wrapper 116-150), in place on cur ------------------------
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
    int d0 = size - m + 1;                                       // tap k=-1 .. k=2 -> d0, d0-1, d0-2, d0-3 (mod size)
    float v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        int d = d0 - k;
        if (d < 0) d += size;
        if (d >= size) d -= size;
        v[k] = rb[(int)((n - d) & mask)];
    }
    return aes_hermite(frac, v[0], v[1], v[2], v[3]);
}

template <int K>
__device__ void aes_stage_octaver(const DevStage &st, const TileCtx &c)
{
    constexpr int T = 128 * K;
    const DevRing rg = c.P->ring[st.ring[0][0]];
    float *rb = c.rings + rg.off;
    const int mask = st.oct_mask, size = st.oct_size;
#pragma unroll
    for (int m = 0; m < T / AES_NT; ++m) {
        const int i = c.tid + AES_NT * m;
        // np.mean over the 2 channels in f32 (octaver.py:124-126)
        rb[(int)((c.n0 + i) & mask)] = __fmul_rn(__fadd_rn(c.cur[i], c.cur[T + i]), 0.5f);
    }
    __syncthreads();
    const float wet_g = st.mix, dry_g = (float)(1.0 - (double)st.mix);
#pragma unroll
    for (int m = 0; m < T / AES_NT; ++m) {
        const int i = c.tid + AES_NT * m;
        const long long n = c.n0 + i;
        double ph = st.ph0 + (double)n * st.step;                // phasor in closed form
        ph -= floor(ph);
        double p2 = ph + 0.5;
        if (p2 >= 1.0) p2 -= 1.0;
        const float s1 = aes_octaver_tap(rb, mask, n, size, st.fsize, ph);
        const float s2 = aes_octaver_tap(rb, mask, n, size, st.fsize, p2);
        const float sn = sinpif((float)ph);
        const float g1 = sn * sn;                                // 0.5*(1-cos(2*pi*p))
        const float g2 = 1.0f - g1;                              // p2 = p + 1/2
        const float wet = s1 * g1 + s2 * g2;
        c.cur[i] = __fadd_rn(__fmul_rn(c.cur[i], dry_g), __fmul_rn(wet, wet_g));
        c.cur[T + i] = __fadd_rn(__fmul_rn(c.cur[T + i], dry_g), __fmul_rn(wet, wet_g));
    }
    __syncthreads();
}

// ---- distortion (our definition, no reference block) -----------------------------------------
template <int K>
__device__ void aes_stage_distortion(const DevStage &st, const TileCtx &c)
{
    constexpr int T = 128 * K;
    const float drive = st.drive, mix = st.mix, dry = 1.0f - st.mix;
#pragma unroll
    for (int m = 0; m < 2 * T / AES_NT; ++m) {
        const int e = c.tid + AES_NT * m;
        const float v = c.cur[e];
        const float t = tanhf(__fmul_rn(drive, v));
        c.cur[e] = aes_clip1(__fadd_rn(__fmul_rn(dry, v), __fmul_rn(mix, t)));
    }
    __syncthreads();
}

// ---- the kernel body ------------------------------------------------------------------------
template <int K>
__device__ void aes_chain_body(const ChainArgs &a)
{
    constexpr int T = 128 * K;
    AES_DYN_SMEM(float, smem);
    const DevPlan *P = a.plan;
    TileCtx c;
    c.P = P;
    c.tid = threadIdx.x;
    c.cur = smem;
    c.aux = smem + 2 * T;
    c.rings = smem + 4 * T;
    const int foff = (4 * T + P->smem_floats + 1) & ~1;
    c.wtot = reinterpret_cast<double *>(smem + foff);
    double *state = c.wtot + 64;
    const int nstate = P->n_state;
    c.rpos = reinterpret_cast<int *>(state + 2 * nstate);
    c.gscr = a.scratch + (long long)blockIdx.x * P->scratch_floats;
    const int nst = P->n_stages, nr = P->n_rings;

    for (long long b = blockIdx.x; b < a.B; b += gridDim.x) {
        // fresh block state for every clip (the chain's re-prepare, core.py:123-129)
        for (int i = c.tid; i < P->smem_floats; i += AES_NT) c.rings[i] = 0.0f;
        for (int i = c.tid; i < nr; i += AES_NT) c.rpos[i] = 0;
        for (int i = c.tid; i < nstate; i += AES_NT) state[i] = P->stage[i >> 4].init[i & 15];
        __syncthreads();
        int par = 0;
        for (long long n0 = 0; n0 < a.N; n0 += T, par ^= 1) {
            c.n0 = n0;
            c.len = (a.N - n0 < (long long)T) ? (int)(a.N - n0) : T;
            aes_load_tile<K>(a, b, c);
            __syncthreads();
            for (int s = 0; s < nst; ++s) {
                const DevStage &st = P->stage[s];
                const double *sin = state + par * nstate + 16 * s;
                double *sout = state + (par ^ 1) * nstate + 16 * s;
                switch (st.kind) {
                case AESK_DELAY:      aes_stage_delay<K>(st, c); break;
                case AESK_REVERB:     aes_stage_reverb<K>(st, c, sin, sout); break;
                case AESK_BIQUAD:     aes_stage_biquad<K>(st, c, sin, sout); break;
                case AESK_GATE:       aes_stage_gate<K>(st, c, sin, sout); break;
                case AESK_OCTAVER:    aes_stage_octaver<K>(st, c); break;
                case AESK_DISTORTION: aes_stage_distortion<K>(st, c); break;
                default: break;
                }
            }
            aes_store_tile<K>(a, b, c);
            if (c.tid < nr) {
                const DevRing rg = P->ring[c.tid];
                int p = c.rpos[c.tid] + rg.tinc;
                if (p >= rg.len) p -= rg.len;
                c.rpos[c.tid] = p;
            }
            __syncthreads();
        }
        if (a.state_out != nullptr)
            for (int i = c.tid; i < nstate; i += AES_NT) a.state_out[b * nstate + i] = state[par * nstate + i];
    }
}

#ifndef AES_CPU_EMU
template <int K>
__global__ void __launch_bounds__(AES_NT, 2) aes_chain_kernel(const ChainArgs a)
{
    aes_chain_body<K>(a);
}
#endif
