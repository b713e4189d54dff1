// aes_biquad_scan.cuh -- time-parallel biquad cascade for ONE (or a few) long clips:
// BASELINE configs[1] (60 s stereo clip, 4 biquads).  A batch kernel with one CTA per
// clip would walk 2813 tiles serially; here every tile of 1024 frames is its own CTA and
// the recurrence is carried across CTAs by a single-pass chained scan with decoupled
// look-back (Merrill & Garland), once per cascade stage, all in one launch.
//
// Per stage the filter runs in transposed direct form II,
//     y = b0*x + s1 ;  s1' = b1*x - a1*y + s2 ;  s2' = b2*x - a2*y,
// the same transfer function and poles as the reference's DF-I loop (filter.py:8-40) but
// with a 2-vector state that does not involve past inputs, so a tile's zero-state
// response needs nothing from its predecessor.  All arithmetic is f64 (the reference
// promotes to f64; f32 state misses the 1e-5 bar below ~100 Hz); the
// difference to DF-I is rounding at the 1e-16 level.
//
// Tile t (one CTA of AESB_NT = 1024 / AESB_FR threads), stage s:
//   1. zero-state pass over the thread's AESB_FR frames -> outputs yz (kept) and chunk end state e
//   2. warp Kogge-Stone scan with A^(FR*2^k), Horner over the warp totals in smem -> tile aggregate E
//   3. look-back -> true state C at the tile start.  Filters that forget within AESB_LBW tiles (all
//      but extreme low-cutoff / high-Q ones): C = sum of the lb_k nearest AGGREGATES, which are
//      published as self-validating 16-byte (value, epoch) records before any look-back -- no chain
//      across tiles.  Otherwise the original chained decoupled look-back: flags, aggregates and
//      inclusive states E + M*C, AESB_NT predecessors per window.
//   4. outputs = yz + row0(A^j) . (state at the thread's first frame); they are stage s+1's inputs
// Tiles take their index from an atomic ticket so every predecessor is resident or done.
#pragma once
#include "aes_plan.h"

#define AESB_MAX_STAGES 8
#ifndef AESB_FR
#define AESB_FR 16                      // frames per thread: the scan / Horner / look-back cost is per warp, not per frame
#endif
#define AESB_T 1024                     // frames per tile (CTA)
#define AESB_NT (AESB_T / AESB_FR)      // threads per CTA
#define AESB_NW (AESB_NT / 32)          // warps per CTA
#ifndef AESB_KEEP_YZ
#define AESB_KEEP_YZ (AESB_FR <= 8)     // keep the zero-state outputs (2 DFMA per frame in step 4) or rerun the recurrence (5)
#endif
#define AESB_LBW 256                    // tile-power table entries per stage = deepest truncated look-back

struct BqStage {
    double b0, b1, b2, a1, a2;
    double pw[5][4];        // A^(4*2^k), k=0..4
    double wp[8][4];        // P^w, P = A^(32*AESB_FR) (one warp of chunks); w < AESB_NW used
    double row[AESB_FR - 1][2];   // first row of A^j, j = 1..AESB_FR-1: what a start state adds to output j
    double tile[4];         // M = A^1024
    double tile256[4];      // M^256 (one look-back window)
    double init[2][2];      // TDF-II state at the start of the clip, per channel
    int lb_k;               // > 0: tiles further back than lb_k contribute < 2^-44 (aes_biquad_build.h): the
    int pad_;               //      look-back sums lb_k aggregates and never chains; 0: chained look-back
};

// One published double of a tile aggregate with the launch's epoch beside it: written and read as ONE
// 16-byte access, so a reader that sees the epoch has the value -- no flag, no fence, no clearing
// between launches (the packing CUB's decoupled look-back uses for its tile descriptors).
struct alignas(16) BqRec { double v; unsigned long long tag; };
#define AESB_SMEM_DOUBLES 112

struct BqArgs {
    BqStage st[AESB_MAX_STAGES];
    const float *x;
    float *y;
    long long N;            // frames per clip
    long long n_tiles;      // per clip
    long long B;            // clips
    int n_stages;
    int dbg_skip;           // tests only: treat inclusive records as aggregates unless tile % dbg_skip == 0
    // global scan state: [clip][stage][tile] records, and per-lane / per-warp power tables
    double *agg;            // 4 doubles (2 ch x 2) per record
    double *inc;
    int *flag;              // 0 none, 1 aggregate, 2 inclusive
    unsigned int *ticket;   // one counter, never reset: a launch's tickets start at ticket_base
    unsigned int ticket_base;
    BqRec *rec16;           // truncated look-back: 4 self-validating (value, epoch) pairs per record
    unsigned long long epoch;   // this launch's tag (never 0; the array starts zeroed and is never cleared again)
    const double *lane_pw;  // [stage][32][4]  A^(4*lane)
    const double *tile_pw;  // [stage][AESB_LBW][4] M^i, i = look-back distance - 1
    double *final_state;    // optional [clip][stage][16]: [4*ch + {0,1,2,3}] = x1,x2,y1,y2 (DF-I view) at the clip end
};

__device__ __forceinline__ void bq_matvec(const double m[4], double v1, double v2, double &o1, double &o2)
{
    o1 = m[0] * v1 + m[1] * v2;
    o2 = m[2] * v1 + m[3] * v2;
}

#ifdef AES_CPU_EMU
static inline int bq_ld_flag(const int *p) { return *p; }
static inline int bq_ld_flag_acquire(const int *p) { return *p; }
static inline void bq_st_flag(int *p, int v) { *p = v; }
static inline void __threadfence() {}
static inline void bq_st_rec(BqRec *p, double v, unsigned long long tag) { p->v = v; p->tag = tag; }
static inline bool bq_ld_rec(const BqRec *p, unsigned long long tag, double &v) { v = p->v; return p->tag == tag; }
static inline unsigned atomicAdd(unsigned *p, unsigned v) { unsigned o = *p; *p += v; return o; }
#else
// spin with relaxed loads (no L1 invalidation per iteration); one acquire fence once the
// flag is seen orders the following reads of the record
__device__ __forceinline__ int bq_ld_flag(const int *p)
{
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// (re-reading the flag with ld.acquire orders the record reads after it without a full fence)
__device__ __forceinline__ int bq_ld_flag_acquire(const int *p)
{
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void bq_st_rec(BqRec *p, double v, unsigned long long tag)
{
    asm volatile("st.relaxed.gpu.global.v2.b64 [%0], {%1, %2};" ::"l"(p), "l"(__double_as_longlong(v)), "l"(tag) : "memory");
}
__device__ __forceinline__ bool bq_ld_rec(const BqRec *p, unsigned long long tag, double &v)
{
    long long b; unsigned long long t;
    asm volatile("ld.relaxed.gpu.global.v2.b64 {%0, %1}, [%2];" : "=l"(b), "=l"(t) : "l"(p) : "memory");
    v = __longlong_as_double(b);
    return t == tag;
}
__device__ __forceinline__ void bq_st_flag(int *p, int v)
{
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
#endif

__device__ void aes_biquad_scan_body(const BqArgs &a)
{
    AES_DYN_SMEM(double, sm);                     // 2 x [8 warps][2 ch][2] totals (by stage parity) | ticket | look-back
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned *tk = reinterpret_cast<unsigned *>(sm + 64);
    double *lb = sm + 66;                         // chained look-back: [8 warps][4 partial sums + found]

    if (tid == 0) *tk = atomicAdd(a.ticket, 1u) - a.ticket_base;
    __syncthreads();
    const long long work = (long long)*tk;        // ordered: clip-major, tile-minor
    if (work >= a.B * a.n_tiles) return;
    const long long clip = work / a.n_tiles, tile = work % a.n_tiles;
    const long long n0 = tile * AESB_T;
    const int len = (a.N - n0 < (long long)AESB_T) ? (int)(a.N - n0) : AESB_T;
    const int i0 = AESB_FR * tid;

    // load AESB_FR frames x 2 channels; between stages the samples stay f32 (what the reference stores)
    float v[2][AESB_FR];
    {
        const float2 *xp = reinterpret_cast<const float2 *>(a.x) + clip * a.N + n0 + i0;
#pragma unroll
        for (int j = 0; j < AESB_FR; ++j) {
            float2 t = make_float2(0.f, 0.f);
            if (i0 + j < len) t = xp[j];
            v[0][j] = t.x; v[1][j] = t.y;
        }
    }

    for (int s = 0; s < a.n_stages; ++s) {
        const BqStage &st = a.st[s];
        const double b0 = st.b0, b1 = st.b1, b2 = st.b2, a1 = st.a1, a2 = st.a2;
        const long long rec = (clip * a.n_stages + s) * a.n_tiles + tile;
        double *const wtot = sm + 32 * (s & 1);     // a warp may be one stage ahead of another, never two
        // 1. zero-state chunk response: outputs kept, the true start state only adds row0(A^j).S later
        double yz[2][AESB_KEEP_YZ ? AESB_FR : 1], e1[2], e2[2];
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
            double s1 = 0.0, s2 = 0.0;
#pragma unroll
            for (int j = 0; j < AESB_FR; ++j) {
                const double xj = (double)v[ch][j];
                const double y = fma(b0, xj, s1);
                s1 = fma(b1, xj, fma(-a1, y, s2));
                s2 = fma(b2, xj, -a2 * y);
                if (AESB_KEEP_YZ) yz[ch][AESB_KEEP_YZ ? j : 0] = y;
            }
            e1[ch] = s1; e2[ch] = s2;
        }
        // 2. warp Kogge-Stone scan of the chunk end states
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            const double m0 = st.pw[k][0], m1 = st.pw[k][1], m2 = st.pw[k][2], m3 = st.pw[k][3];
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
                const double u1 = __shfl_up_sync(0xffffffffu, e1[ch], 1 << k);
                const double u2 = __shfl_up_sync(0xffffffffu, e2[ch], 1 << k);
                if (lane >= (1 << k)) {
                    e1[ch] = fma(m0, u1, fma(m1, u2, e1[ch]));
                    e2[ch] = fma(m2, u1, fma(m3, u2, e2[ch]));
                }
            }
        }
        if (lane == 31) {
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) { wtot[(warp * 2 + ch) * 2] = e1[ch]; wtot[(warp * 2 + ch) * 2 + 1] = e2[ch]; }
        }
        __syncthreads();
        // zero-start state at the warp's first frame (Horner over the previous warps' totals, P = A^128)
        // and, per thread, the exclusive scan value inside the warp
        double c1[2], c2[2], x1[2], x2[2];
        {
            const double P0 = st.wp[1][0], P1 = st.wp[1][1], P2 = st.wp[1][2], P3 = st.wp[1][3];
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
                double q1 = 0.0, q2 = 0.0;
                for (int w = 0; w < warp; ++w) {
                    const double n1 = fma(P0, q1, fma(P1, q2, wtot[(w * 2 + ch) * 2]));
                    const double n2 = fma(P2, q1, fma(P3, q2, wtot[(w * 2 + ch) * 2 + 1]));
                    q1 = n1; q2 = n2;
                }
                c1[ch] = q1; c2[ch] = q2;
                x1[ch] = __shfl_up_sync(0xffffffffu, e1[ch], 1); x2[ch] = __shfl_up_sync(0xffffffffu, e2[ch], 1);
                if (lane == 0) { x1[ch] = 0.0; x2[ch] = 0.0; }
            }
        }
        // 3. tile aggregate, look-back -> state C at the tile start
        double acc[4] = { 0.0, 0.0, 0.0, 0.0 };         // carry-in, the same in every thread
        const int K = a.dbg_skip > 0 ? 0 : st.lb_k;
        if (K > 0) {
            // Truncated look-back: the filter forgets, M^K is below 2^-44, so C is the sum of the K nearest
            // AGGREGATES (zero-state tile responses, published before any look-back): no tile waits for
            // another tile's look-back, and the sum has a fixed order (bit-reproducible output).
            // The last warp publishes (its Horner carry covers warps 0..6); the first warp, which has no
            // Horner steps and gets here first, looks back for the CTA (a look-back in every warp cost
            // 8x the polling instructions for the same wait: 138 vs 106 us per 60 s clip).
            if (warp == AESB_NW - 1) {
                const double P0 = st.wp[1][0], P1 = st.wp[1][1], P2 = st.wp[1][2], P3 = st.wp[1][3];
                double ev = 0.0;
#pragma unroll
                for (int ch = 0; ch < 2; ++ch) {
                    const double E1 = fma(P0, c1[ch], fma(P1, c2[ch], wtot[((AESB_NW - 1) * 2 + ch) * 2]));
                    const double E2 = fma(P2, c1[ch], fma(P3, c2[ch], wtot[((AESB_NW - 1) * 2 + ch) * 2 + 1]));
                    if (lane == 2 * ch) ev = E1;
                    if (lane == 2 * ch + 1) ev = E2;
                }
                if (lane < 4) bq_st_rec(a.rec16 + rec * 4 + lane, ev, a.epoch);
            }
            if (warp == 0) {
            int KP = 1;
            while (KP < K && KP < 32) KP <<= 1;
            // lane i (mod KP) takes predecessors i, i + KP, ...; groups of KP lanes work redundantly so the
            // butterfly below leaves the sum in every lane
            for (int i = lane & (KP - 1); i < K; i += KP) {
                const long long pt = tile - 1 - i;
                if (pt < -1) break;
                double val[4];
                if (pt >= 0) {
                    const BqRec *r = a.rec16 + (rec - 1 - i) * 4;
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        while (!bq_ld_rec(r + q, a.epoch, val[q])) { }
                } else {                                    // the clip's initial state sits where tile -1 would
                    val[0] = st.init[0][0]; val[1] = st.init[0][1]; val[2] = st.init[1][0]; val[3] = st.init[1][1];
                }
                const double *tp = a.tile_pw + ((long long)s * AESB_LBW + i) * 4;            // M^i
                acc[0] = fma(tp[0], val[0], fma(tp[1], val[1], acc[0]));
                acc[1] = fma(tp[2], val[0], fma(tp[3], val[1], acc[1]));
                acc[2] = fma(tp[0], val[2], fma(tp[1], val[3], acc[2]));
                acc[3] = fma(tp[2], val[2], fma(tp[3], val[3], acc[3]));
            }
            for (int k = KP >> 1; k >= 1; k >>= 1)
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], k);
            if (lane < 4) lb[lane] = lane == 0 ? acc[0] : lane == 1 ? acc[1] : lane == 2 ? acc[2] : acc[3];
            }
            __syncthreads();
            acc[0] = lb[0]; acc[1] = lb[1]; acc[2] = lb[2]; acc[3] = lb[3];
        } else {
            double E1[2] = { 0.0, 0.0 }, E2[2] = { 0.0, 0.0 };
            if (warp == 0) {                                // only thread 0 publishes
                const double P0 = st.wp[1][0], P1 = st.wp[1][1], P2 = st.wp[1][2], P3 = st.wp[1][3];
#pragma unroll
                for (int ch = 0; ch < 2; ++ch) {
                    double q1 = 0.0, q2 = 0.0;
#pragma unroll
                    for (int w = 0; w < AESB_NW; ++w) {
                        const double n1 = fma(P0, q1, fma(P1, q2, wtot[(w * 2 + ch) * 2]));
                        const double n2 = fma(P2, q1, fma(P3, q2, wtot[(w * 2 + ch) * 2 + 1]));
                        q1 = n1; q2 = n2;
                    }
                    E1[ch] = q1; E2[ch] = q2;
                }
            }
            if (tile > 0 && tid == 0) {                     // publish the aggregate first: successors may run ahead
                a.agg[rec * 4 + 0] = E1[0]; a.agg[rec * 4 + 1] = E2[0];
                a.agg[rec * 4 + 2] = E1[1]; a.agg[rec * 4 + 3] = E2[1];
                bq_st_flag(a.flag + rec, 1);                // st.release orders the record before the flag
            }
            double W[4] = { 1.0, 0.0, 0.0, 1.0 };           // M^base
            long long base = 0;
            for (;;) {
                // thread i looks at predecessor tile-1-base-i; index -1 is the clip's initial state (inclusive)
                const long long pt = tile - 1 - base - tid;
                int f = 2;
                double val[4] = { 0.0, 0.0, 0.0, 0.0 };
                if (pt >= 0) {
                    const long long prec = rec - 1 - base - tid;
                    do { f = bq_ld_flag(a.flag + prec); } while (f == 0);
                    f = bq_ld_flag_acquire(a.flag + prec);      // >= the value just seen: flags only go 0 -> 1 -> 2
                    const bool incl_ok = !(a.dbg_skip > 0 && pt % a.dbg_skip != 0);   // tests: force the aggregate path
                    if (f == 2 && !incl_ok) f = 1;
                    const double *src = (f == 2 ? a.inc : a.agg) + prec * 4;
                    val[0] = src[0]; val[1] = src[1]; val[2] = src[2]; val[3] = src[3];
                } else if (pt == -1) {
                    val[0] = st.init[0][0]; val[1] = st.init[0][1]; val[2] = st.init[1][0]; val[3] = st.init[1][1];
                }
                // nearest inclusive predecessor inside each warp, and the warp's partial sum up to it
                const unsigned incl = __ballot_sync(0xffffffffu, f == 2);
                const int first = incl ? __ffs((int)incl) - 1 : 32;
                double t[4] = { 0.0, 0.0, 0.0, 0.0 };
                if (lane <= first && pt >= -1) {
                    const double *tp = a.tile_pw + ((long long)s * AESB_LBW + tid) * 4;      // M^tid
                    bq_matvec(tp, val[0], val[1], t[0], t[1]);
                    bq_matvec(tp, val[2], val[3], t[2], t[3]);
                }
#pragma unroll
                for (int k = 16; k >= 1; k >>= 1)
#pragma unroll
                    for (int q = 0; q < 4; ++q) t[q] += __shfl_xor_sync(0xffffffffu, t[q], k);
                if (lane == 0) {
                    lb[warp * 5 + 0] = t[0]; lb[warp * 5 + 1] = t[1]; lb[warp * 5 + 2] = t[2]; lb[warp * 5 + 3] = t[3];
                    lb[warp * 5 + 4] = first < 32 ? 1.0 : 0.0;
                }
                __syncthreads();
                double sm4[4] = { 0.0, 0.0, 0.0, 0.0 };
                bool found = false;
                for (int w = 0; w < AESB_NW && !found; ++w) {
                    sm4[0] += lb[w * 5 + 0]; sm4[1] += lb[w * 5 + 1]; sm4[2] += lb[w * 5 + 2]; sm4[3] += lb[w * 5 + 3];
                    found = lb[w * 5 + 4] != 0.0;
                }
                __syncthreads();                              // lb is rewritten by the next window
                double o0, o1, o2, o3;
                bq_matvec(W, sm4[0], sm4[1], o0, o1);
                bq_matvec(W, sm4[2], sm4[3], o2, o3);
                acc[0] += o0; acc[1] += o1; acc[2] += o2; acc[3] += o3;
                if (found) break;
                const double n0w = W[0] * st.tile256[0] + W[1] * st.tile256[2], n1w = W[0] * st.tile256[1] + W[1] * st.tile256[3];
                const double n2w = W[2] * st.tile256[0] + W[3] * st.tile256[2], n3w = W[2] * st.tile256[1] + W[3] * st.tile256[3];
                W[0] = n0w; W[1] = n1w; W[2] = n2w; W[3] = n3w;
                base += AESB_NT;
            }
            if (tid == 0) {
                double i0v, i1v, i2v, i3v;
                bq_matvec(st.tile, acc[0], acc[1], i0v, i1v);
                bq_matvec(st.tile, acc[2], acc[3], i2v, i3v);
                a.inc[rec * 4 + 0] = E1[0] + i0v; a.inc[rec * 4 + 1] = E2[0] + i1v;
                a.inc[rec * 4 + 2] = E1[1] + i2v; a.inc[rec * 4 + 3] = E2[1] + i3v;
                bq_st_flag(a.flag + rec, 2);
            }
        }
        // 4. true state at this thread's first frame: exclusive scan value + A^(4*lane) (warp carry + P^warp C);
        //    the outputs are the zero-state outputs plus row0(A^j) . state
        {
            const double *lp = a.lane_pw + (s * 32 + lane) * 4;
            const double l0 = lp[0], l1 = lp[1], l2 = lp[2], l3 = lp[3];
            const double W0 = st.wp[warp][0], W1 = st.wp[warp][1], W2 = st.wp[warp][2], W3 = st.wp[warp][3];
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
                const double C1 = acc[2 * ch], C2 = acc[2 * ch + 1];
                const double cw1 = fma(W0, C1, fma(W1, C2, c1[ch])), cw2 = fma(W2, C1, fma(W3, C2, c2[ch]));
                double s1 = fma(l0, cw1, fma(l1, cw2, x1[ch])), s2 = fma(l2, cw1, fma(l3, cw2, x2[ch]));
#pragma unroll
                for (int j = 0; j < AESB_FR; ++j) {
                    const double xj = (double)v[ch][j];
                    double y;
                    if (AESB_KEEP_YZ) {
                        y = j == 0 ? yz[ch][0] + s1
                                   : fma(st.row[j - (j > 0)][0], s1, fma(st.row[j - (j > 0)][1], s2, yz[ch][AESB_KEEP_YZ ? j : 0]));
                    } else {                              // many frames per thread: rerun the recurrence, keep no outputs
                        y = fma(b0, xj, s1);
                        s1 = fma(b1, xj, fma(-a1, y, s2));
                        s2 = fma(b2, xj, -a2 * y);
                    }
                    v[ch][j] = (float)y;                  // the reference stores every stage's output as f32
                    if (a.final_state != nullptr && tile == a.n_tiles - 1) {
                        double *fs = a.final_state + (clip * a.n_stages + s) * 16 + 4 * ch;
                        if (i0 + j == len - 1) { fs[0] = xj; fs[2] = y; }
                        if (i0 + j == len - 2) { fs[1] = xj; fs[3] = y; }
                    }
                }
            }
        }
        // (no barrier here: every read of wtot / lb above sits before a barrier the next stage's writes come after)
    }
    {
        float2 *yp = reinterpret_cast<float2 *>(a.y) + clip * a.N + n0 + i0;
#pragma unroll
        for (int j = 0; j < AESB_FR; ++j)
            if (i0 + j < len) yp[j] = make_float2(v[0][j], v[1][j]);
    }
}

#ifndef AES_CPU_EMU
#ifndef AESB_MIN_CTAS
#define AESB_MIN_CTAS (AESB_FR == 32 ? 16 : AESB_FR == 16 ? 10 : AESB_FR == 8 ? 5 : 4)
#endif
__global__ void __launch_bounds__(AESB_NT, AESB_MIN_CTAS) aes_biquad_scan_kernel(const __grid_constant__ BqArgs a)
{
    aes_biquad_scan_body(a);
}
#endif
