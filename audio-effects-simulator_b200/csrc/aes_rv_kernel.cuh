// aes_rv_kernel.cuh -- software-pipelined kernel for chains that END in the reference's default
// reverb topology (reverb.py:74-81,158-177 at 48 kHz / 44.1 kHz): Rain Delay, Cathedral, Guitar
// Filter, the bare reverb.
//
// Why a second kernel: ncu on aes_fast_kernel (profiles/r1o) showed the reverb chains issue-bound
// with 47 % of the executed instructions integer / control / uniform bookkeeping, and five CTA
// barriers per tile that put the latency-bound all-pass walks strictly behind the comb bank.
// Here
//   * the tile loop is software-pipelined:  the all-pass walks, the dry/wet mix and the store of
//     tile i-1 run in the same barrier phases as the delay / comb work of tile i (two tile buffers
//     alternate), so a tile costs TWO CTA barriers instead of five and warps that have no
//     all-pass column to walk go straight on to comb work;
//
//         phase 1:  all-pass 1 of tile i-1   |  input, feedback delay / biquad pass 1, comb ring
//                                            |  reads, local one-pole ends, warp scan        -- barrier
//         phase 2:  all-pass 2 of tile i-1   |  carries, comb update + ring writes, comb sum -- barrier
//         phase 3:  mix + store of tile i-1, slots advance, loads for tile i+1 issued
//
//   * the next tile's input frames and feedback-delay line samples are fetched with plain 128-bit
//     loads issued in phase 3 (registers are slack there) instead of a TMA staging round trip
//     through shared memory: the shared-memory data pipe was 56 % busy and is the second limiter;
//   * everything about the reverb is a compile-time constant (TOPO), per-lane constants are loaded
//     once per kernel, carried values move as float4, all-pass ring phases live in registers.
// Arithmetic per sample is exactly that of aes_fast_kernel (same parity tests).  Any other chain
// shape / sample rate / comb set keeps using aes_fast_kernel or the interpreter.
#pragma once
#include "aes_fast_kernel.cuh"

#define AESRV_PRE_NONE 0
#define AESRV_PRE_DELAY 1           // prefetched feedback delay ahead of the reverb (lag >= 2T, global lines)
#define AESRV_PRE_BIQUAD 2          // one biquad ahead of the reverb

// shared memory (floats): S[2][2][T] | rings[smem_floats] | wt[8][8] | cst[2][8] | f64: wtot[32] | bst[2][8]
__host__ __device__ inline size_t aes_rv_smem_bytes(int smem_floats)
{
    const size_t T = AES_NT * 4;
    size_t f = (4 * T + (size_t)smem_floats + 3) & ~(size_t)3;
    return (f + 64 + 16) * 4 + (32 + 16) * 8 + 16;
}

// all-pass walk over one channel's tile `s` (T samples): column j of the tile, seen as rows of L
// samples, is one serial chain  y = line - g*x,  line' = x + g*y  (reverb.py:48-67) down the rows;
// the chain's state enters from / leaves to the ring (length L, phase `pos` = tile start mod L).
template <int L>
__device__ __forceinline__ void aesrv_walk(float *s, float *rb, int pos, int j0, float g)
{
    constexpr int T = AES_NT * 4;
    static_assert(L > 0 && L < T, "");
#pragma unroll
    for (int r = 0; r < (L + 127) / 128; ++r) {
        const int j = j0 + 128 * r;
        if (128 * r + 127 < L || j < L) {
            constexpr int KMAX = (T + L - 1) / L;           // rows that can hold column j
            int slot = pos + j;
            if (slot >= L) slot -= L;
            float line = rb[slot];
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

template <int TOPO, int K>
__device__ __forceinline__ void aesrv_allpass(float *S, float *rings, int ch, int j0, int pos, float g)
{
    constexpr int T = AES_NT * 4;
    if (ch == 0) aesrv_walk<aesf_topo_ap(TOPO, 0, K)>(S, rings + aesf_topo_ap_off(TOPO, 0, K), pos, j0, g);
    else         aesrv_walk<aesf_topo_ap(TOPO, 1, K)>(S + T, rings + aesf_topo_ap_off(TOPO, 1, K), pos, j0, g);
}

template <int TOPO, int PRE, int PM>
__device__ void aes_rv_body(const FastArgs &a)
{
    constexpr int FR = 4, T = AES_NT * FR, NC = 4;
    constexpr int SR = PRE == AESRV_PRE_NONE ? 0 : 1;      // index of the reverb stage
    static_assert(TOPO == AESF_TOPO_48K || TOPO == AESF_TOPO_44K, "compile-time reverb topology only");
    static_assert(!(PRE == AESRV_PRE_BIQUAD && PM != 0), "biquad + pre-delay is not instantiated (the biquad output only exists in phase 2)");
    AES_DYN_SMEM(float, smem);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, i0 = FR * tid;
    const int wch = tid >> 7, wj0 = tid & 127;              // all-pass walks: threads 0..127 left, 128..255 right
    float *const rings = smem + 4 * T;
    const int foff = (4 * T + a.smem_floats + 3) & ~3;
    float *const wt = smem + foff;                          // [8 warps][2 ch][4 combs] warp totals of the comb one-poles
    float *const cst = wt + 64;                             // [2 parity][8] one-pole values carried across tiles
    double *const wtot = reinterpret_cast<double *>(cst + 16);  // biquad: [8 warps][2 ch][2]
    double *const bst = wtot + 32;                          // biquad: [2 parity][8] DF-I state carried across tiles
    float *const gscr = a.scratch + (long long)blockIdx.x * a.scratch_floats;

    const FastStage &rs = a.st[SR];
    const float h = rs.h, hw = rs.hp[5], apg = rs.a, rdry = rs.dry, rwet = rs.wet;
    const int nscan = rs.nscan, nxw = rs.nxw;
    const float hl = a.lane_tab[(SR * 32 + lane) * FAST_LANE_STRIDE];       // h^(FR*lane)
    // where this warp finds the one-pole values entering its first frame: the previous warp's totals,
    // or (warp 0) the values carried from the previous tile, by tile parity
    const float *const cp0 = warp == 0 ? cst : wt + (warp - 1) * 8;
    const float *const cp1 = warp == 0 ? cst + 8 : wt + (warp - 1) * 8;
    constexpr int apT1_0 = T % aesf_topo_ap(TOPO, 0, 0), apT1_1 = T % aesf_topo_ap(TOPO, 1, 0);
    constexpr int apT2_0 = T % aesf_topo_ap(TOPO, 0, 1), apT2_1 = T % aesf_topo_ap(TOPO, 1, 1);
    const int apL1 = wch ? aesf_topo_ap(TOPO, 1, 0) : aesf_topo_ap(TOPO, 0, 0), apI1 = wch ? apT1_1 : apT1_0;
    const int apL2 = wch ? aesf_topo_ap(TOPO, 1, 1) : aesf_topo_ap(TOPO, 0, 1), apI2 = wch ? apT2_1 : apT2_0;

    ChainArgs io;                                           // tile I/O helpers are shared with the generic kernel
    io.x = a.x; io.y = a.y; io.N = a.N; io.in_fmt = a.in_fmt; io.out_fmt = a.out_fmt;
    const int N = (int)a.N;                                 // the host routes clips of 2^31 frames or more elsewhere
    const int ntiles = (N + T - 1) / T;

    for (long long b = blockIdx.x; b < a.B; b += gridDim.x) {
        {   // fresh lines, carried values and ring phases at every clip start (core.py:123-129 re-prepares)
            float4 *r4 = reinterpret_cast<float4 *>(rings);
            const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int i = tid; i < (a.smem_floats + 3) / 4; i += AES_NT) r4[i] = z4;
            if (tid < 16) cst[tid] = 0.0f;
            if (PRE == AESRV_PRE_BIQUAD && tid < 8) { bst[tid] = a.init[0][tid]; bst[8 + tid] = a.init[0][tid]; }
        }
        int wb[2][NC];                                      // comb slots: BYTE offset of this thread's float4 in the ring
#pragma unroll
        for (int ch = 0; ch < 2; ++ch)
#pragma unroll
            for (int cc = 0; cc < NC; ++cc) wb[ch][cc] = 4 * i0;
        int dw[2] = { 0, 0 }, da[2] = { 0, 0 };             // feedback-delay line: write slot, aligned read base
        if constexpr (PRE == AESRV_PRE_DELAY) {
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) aesf_line_init<FR>(a.st[0].ring[ch][0], i0, dw[ch], da[ch]);
        }
        int pw[2] = { 0, 0 }, pa[2] = { 0, 0 };             // reverb pre-delay line
        if constexpr (PM != 0) {
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) aesf_line_init<FR>(rs.pre[ch], i0, pw[ch], pa[ch]);
        }
        int pos1 = 0, pos2 = 0;                             // all-pass ring phases of the tile the walks work on
        __syncthreads();

        const bool fast_in = a.in_fmt == AESK_F32_STEREO && ((b * a.N) & 1) == 0;
        const float *const xin = reinterpret_cast<const float *>(a.x) + 2 * (b * a.N);
        float4 pfx0, pfx1;                                  // next tile's 4 stereo frames of this thread
        float4 lnA[2], lnB[2];                              // next tile's feedback-delay line samples
        pfx0 = pfx1 = lnA[0] = lnA[1] = lnB[0] = lnB[1] = make_float4(0.f, 0.f, 0.f, 0.f);
        bool pf_ok = false;                                 // pfx holds the coming tile
        if (fast_in && N >= T) {
            pfx0 = aes_ldg_v4(xin + 2 * i0);
            pfx1 = aes_ldg_v4(xin + 2 * i0 + 4);
            pf_ok = true;
        }
        float vprev[2][FR];                                 // dry signal at the reverb input, tile i-1
#pragma unroll
        for (int ch = 0; ch < 2; ++ch)
#pragma unroll
            for (int j = 0; j < FR; ++j) vprev[ch][j] = 0.0f;

        int par = 0;
        for (int it = 0; it <= ntiles; ++it, par ^= 1) {
            const bool has_cur = it < ntiles, has_prev = it > 0;
            const int n0 = it * T;
            const int rem = N - n0;
            const int len = rem < T ? rem : T;
            float *const Sc = smem + par * 2 * T;           // this tile's buffer
            float *const Sp = smem + (par ^ 1) * 2 * T;     // tile i-1's
            float v[2][FR], y[2][NC][FR], e[2][NC];
            double bq_yz[2][FR], bq_e1[2], bq_e2[2];

            // ---------------- phase 1 ----------------
            if (has_prev) aesrv_allpass<TOPO, 0>(Sp, rings, wch, wj0, pos1, apg);
            if (has_cur) {
                if (pf_ok) {
                    v[0][0] = pfx0.x; v[1][0] = pfx0.y; v[0][1] = pfx0.z; v[1][1] = pfx0.w;
                    v[0][2] = pfx1.x; v[1][2] = pfx1.y; v[0][3] = pfx1.z; v[1][3] = pfx1.w;
                } else {
                    aes_load_frames<FR>(io, b, n0, len, tid, v);            // ragged / unaligned / non-f32 tiles
                }
                if constexpr (PRE == AESRV_PRE_DELAY) {
                    // feedback delay (delay.py:7-22,94-96); every delayed sample predates the tile
                    const FastStage &ds = a.st[0];
                    const float fb = ds.fb, dry = ds.dry, wet = ds.wet;
#pragma unroll
                    for (int ch = 0; ch < 2; ++ch) {
                        const FRing rg = ds.ring[ch][0];
                        const int m = ((rg.lag + 3) & ~3) - rg.lag;
                        float line[FR];
                        if (!pf_ok || it == 0) {                            // nothing was fetched ahead for this tile
                            const float *rb = gscr + rg.off;
                            int b0 = da[ch] + 4;
                            b0 = b0 >= rg.len ? b0 - rg.len : b0;
                            lnA[ch] = aes_ldg_v4(rb + da[ch]);
                            lnB[ch] = aes_ldg_v4(rb + b0);
                        }
                        if (m == 0) { line[0] = lnA[ch].x; line[1] = lnA[ch].y; line[2] = lnA[ch].z; line[3] = lnA[ch].w; }
                        else aesf_select4<FR>(lnA[ch], lnB[ch], m, line);
                        if (n0 < rg.lag) {                                  // only the first tiles of a clip: zero history
#pragma unroll
                            for (int j = 0; j < FR; ++j)
                                if (n0 + i0 + j < rg.lag) line[j] = 0.0f;
                        }
                        float nb[FR];
#pragma unroll
                        for (int j = 0; j < FR; ++j) {
                            const float x = v[ch][j];
                            nb[j] = fmaf(line[j], fb, x);
                            v[ch][j] = aes_mix_clip(dry, x, wet, line[j]);
                        }
                        aes_stv<FR>(gscr + rg.off + dw[ch], nb);
                    }
                }
                if constexpr (PRE == AESRV_PRE_BIQUAD) {
                    // transposed DF-II from zero state, outputs kept (filter.py:8-40; see aes_fast_kernel.cuh)
                    const FastStage &bs = a.st[0];
                    const double b0 = bs.bq[0], b1 = bs.bq[1], b2 = bs.bq[2], a1 = bs.bq[3], a2 = bs.bq[4];
#pragma unroll
                    for (int ch = 0; ch < 2; ++ch) {
                        double s1 = 0.0, s2 = 0.0;
#pragma unroll
                        for (int j = 0; j < FR; ++j) {
                            const double xj = (double)v[ch][j];
                            const double yy = fma(b0, xj, s1);
                            s1 = fma(b1, xj, fma(-a1, yy, s2));
                            s2 = fma(b2, xj, -a2 * yy);
                            bq_yz[ch][j] = yy;
                        }
                        bq_e1[ch] = s1; bq_e2[ch] = s2;
                    }
                    const int bq_nscan = bs.nscan;
                    for (int s = 0; s < bq_nscan; ++s) {
                        const double m0 = bs.bq_pow[s][0], m1 = bs.bq_pow[s][1], m2 = bs.bq_pow[s][2], m3 = bs.bq_pow[s][3];
#pragma unroll
                        for (int ch = 0; ch < 2; ++ch) {
                            const double u1 = __shfl_up_sync(0xffffffffu, bq_e1[ch], 1 << s);
                            const double u2 = __shfl_up_sync(0xffffffffu, bq_e2[ch], 1 << s);
                            if (lane >= (1 << s)) {
                                bq_e1[ch] = fma(m0, u1, fma(m2, u2, bq_e1[ch]));
                                bq_e2[ch] = fma(m1, u1, fma(m3, u2, bq_e2[ch]));
                            }
                        }
                    }
                    if (lane == 31) {
#pragma unroll
                        for (int ch = 0; ch < 2; ++ch) {
                            wtot[(warp * 2 + ch) * 2] = bq_e1[ch];
                            wtot[(warp * 2 + ch) * 2 + 1] = bq_e2[ch];
                        }
                    }
                }
                // damped combs (reverb.py:33-46) on u = lp/(1-h): delayed ring reads, one-pole over the
                // thread's 4 samples from zero, Kogge-Stone over the warp, lane 31 publishes the warp total
#pragma unroll
                for (int ch = 0; ch < 2; ++ch)
#pragma unroll
                    for (int cc = 0; cc < NC; ++cc) {
                        const int L = aesf_topo_comb(TOPO, ch, cc), rlen = (L + 3) & ~3;
                        aesf_comb_read<FR>(rings + aesf_topo_comb_off(TOPO, ch, cc), wb[ch][cc], rlen - L, rlen, y[ch][cc]);
                        float u = y[ch][cc][0];
#pragma unroll
                        for (int j = 1; j < FR; ++j) u = fmaf(h, u, y[ch][cc][j]);
                        e[ch][cc] = u;
                    }
#pragma unroll 1
                for (int s = 0; s < nscan; ++s) {
                    const float m = rs.hp[s];
#pragma unroll
                    for (int ch = 0; ch < 2; ++ch)
#pragma unroll
                        for (int cc = 0; cc < NC; ++cc) {
                            const float t = __shfl_up_sync(0xffffffffu, e[ch][cc], 1 << s);
                            if (lane >= (1 << s)) e[ch][cc] = fmaf(m, t, e[ch][cc]);
                        }
                }
                if (lane == 31) {
                    aes_stv<4>(wt + warp * 8, e[0]);
                    aes_stv<4>(wt + warp * 8 + 4, e[1]);
                }
                if constexpr (PM != 0) {
                    // pre-delay line (reverb.py:11-31): written here, read behind the barrier
#pragma unroll
                    for (int ch = 0; ch < 2; ++ch) aes_stv<FR>(rings + rs.pre[ch].off + pw[ch], v[ch]);
                }
            }
            __syncthreads();

            // ---------------- phase 2 ----------------
            if (has_prev) aesrv_allpass<TOPO, 1>(Sp, rings, wch, wj0, pos2, apg);
            if (has_cur) {
                if constexpr (PRE == AESRV_PRE_BIQUAD) {
                    const FastStage &bs = a.st[0];
                    const double b1 = bs.bq[1], b2 = bs.bq[2], a1 = bs.bq[3], a2 = bs.bq[4];
                    const double w0 = bs.bq_pow[5][0], w1 = bs.bq_pow[5][1], w2 = bs.bq_pow[5][2], w3 = bs.bq_pow[5][3];
                    const double *lt = reinterpret_cast<const double *>(a.lane_tab + lane * FAST_LANE_STRIDE + 4);
                    const double l0 = lt[0], l1 = lt[1], l2 = lt[2], l3 = lt[3];
                    const double *sin = bst + par * 8;
                    double *sout = bst + (par ^ 1) * 8;
#pragma unroll
                    for (int ch = 0; ch < 2; ++ch) {
                        const double cx1 = sin[4 * ch + 0], cx2 = sin[4 * ch + 1], cy1 = sin[4 * ch + 2], cy2 = sin[4 * ch + 3];
                        double C1 = b1 * cx1 + b2 * cx2 - a1 * cy1 - a2 * cy2;
                        double C2 = b2 * cx1 - a2 * cy1;
                        const int u0 = warp > bs.nxw ? warp - bs.nxw : 0;
                        if (u0 > 0) { C1 = 0.0; C2 = 0.0; }
                        for (int t = u0; t < warp; ++t) {
                            const double t1 = fma(w0, C1, fma(w2, C2, wtot[(t * 2 + ch) * 2]));
                            const double t2 = fma(w1, C1, fma(w3, C2, wtot[(t * 2 + ch) * 2 + 1]));
                            C1 = t1; C2 = t2;
                        }
                        double x1 = __shfl_up_sync(0xffffffffu, bq_e1[ch], 1), x2 = __shfl_up_sync(0xffffffffu, bq_e2[ch], 1);
                        if (lane == 0) { x1 = 0.0; x2 = 0.0; }
                        const double S1 = fma(l0, C1, fma(l2, C2, x1));
                        const double S2 = fma(l1, C1, fma(l3, C2, x2));
#pragma unroll
                        for (int j = 0; j < FR; ++j) {
                            const double yy = j == 0 ? bq_yz[ch][0] + S1 : fma(bs.bq_row[j][0], S1, fma(bs.bq_row[j][1], S2, bq_yz[ch][j]));
                            if (len == T) {
                                if (j >= FR - 2 && tid == AES_NT - 1) {
                                    sout[4 * ch + (FR - 1 - j)] = (double)v[ch][j]; sout[4 * ch + 2 + (FR - 1 - j)] = yy;
                                }
                            } else {
                                if (i0 + j == len - 1) { sout[4 * ch + 0] = (double)v[ch][j]; sout[4 * ch + 2] = yy; }
                                if (i0 + j == len - 2) { sout[4 * ch + 1] = (double)v[ch][j]; sout[4 * ch + 3] = yy; }
                            }
                            v[ch][j] = (float)yy;
                        }
                        if (len == 1 && tid == 0) { sout[4 * ch + 1] = cx1; sout[4 * ch + 3] = cy1; }
                    }
                }
                float pre[2][FR];
                if constexpr (PM == 0) {
#pragma unroll
                    for (int ch = 0; ch < 2; ++ch)
#pragma unroll
                        for (int j = 0; j < FR; ++j) pre[ch][j] = v[ch][j];
                } else {
#pragma unroll
                    for (int ch = 0; ch < 2; ++ch) {
                        const FRing rg = rs.pre[ch];
                        // (the line lives in the zero-initialised ring area: no first-lap check)
                        aesf_read<FR, 0>(rings + rg.off, pa[ch], ((rg.lag + 3) & ~3) - rg.lag, rg.len, pre[ch]);
                    }
                }
                const float *const cp = par ? cp1 : cp0;
                const float4 Ca = aes_lds_v4(cp), Cb = aes_lds_v4(cp + 4);
                const float Cin[2][NC] = { { Ca.x, Ca.y, Ca.z, Ca.w }, { Cb.x, Cb.y, Cb.z, Cb.w } };
                float *const cnext = cst + (par ^ 1) * 8;
                const int u0 = warp > nxw ? warp - nxw : 0;
                float sum[2][FR];
#pragma unroll
                for (int ch = 0; ch < 2; ++ch) {
                    float uend[NC];
#pragma unroll
                    for (int cc = 0; cc < NC; ++cc) {
                        const float ex = __shfl_up_sync(0xffffffffu, e[ch][cc], 1);
                        float C;
                        if (nxw == 1) {                                  // usual case: h^(32*FR) < 2^-32
                            C = Cin[ch][cc];
                        } else {
                            C = u0 == 0 ? cst[par * 8 + ch * 4 + cc] : 0.0f;
                            for (int t = u0; t < warp; ++t) C = fmaf(hw, C, wt[t * 8 + ch * 4 + cc]);
                        }
                        float u = fmaf(hl, C, lane == 0 ? 0.0f : ex);
                        const float gs = rs.gs[ch][cc];
                        float nb[FR];
#pragma unroll
                        for (int j = 0; j < FR; ++j) {
                            u = fmaf(h, u, y[ch][cc][j]);
                            nb[j] = fmaf(gs, u, pre[ch][j]);            // buf[n] = x + g*(1-h)*u
                            if (cc == 0) sum[ch][j] = y[ch][0][j];      // reverb.py:235-241: sum starts at 0
                            else sum[ch][j] = __fadd_rn(sum[ch][j], y[ch][cc][j]);
                        }
                        aes_stv<FR>(reinterpret_cast<float *>(reinterpret_cast<char *>(rings + aesf_topo_comb_off(TOPO, ch, cc)) + wb[ch][cc]), nb);
                        uend[cc] = u;
                    }
                    if (tid == AES_NT - 1) aes_stv<4>(cnext + ch * 4, uend);
                }
                aes_stv<FR>(Sc + i0, sum[0]);
                aes_stv<FR>(Sc + T + i0, sum[1]);
            }
            __syncthreads();

            // ---------------- phase 3 ----------------
            if (has_prev) {
                float o[2][FR];
                aes_ldv_sp<FR, 0>(Sp + i0, o[0]);
                aes_ldv_sp<FR, 0>(Sp + T + i0, o[1]);
#pragma unroll
                for (int ch = 0; ch < 2; ++ch)
#pragma unroll
                    for (int j = 0; j < FR; ++j) o[ch][j] = aes_mix_clip(rdry, vprev[ch][j], rwet, o[ch][j]);
                const int rp = N - (n0 - T);
                aes_store_frames<FR>(io, b, n0 - T, rp < T ? rp : T, tid, o);
                pos1 = aesf_adv(pos1, apI1, apL1);
                pos2 = aesf_adv(pos2, apI2, apL2);
            }
            if (has_cur) {
#pragma unroll
                for (int ch = 0; ch < 2; ++ch)
#pragma unroll
                    for (int j = 0; j < FR; ++j) vprev[ch][j] = v[ch][j];
#pragma unroll
                for (int ch = 0; ch < 2; ++ch)
#pragma unroll
                    for (int cc = 0; cc < NC; ++cc) {
                        const int rlen = (aesf_topo_comb(TOPO, ch, cc) + 3) & ~3;
                        wb[ch][cc] = aesf_adv(wb[ch][cc], 4 * (T % rlen), 4 * rlen);
                    }
                if constexpr (PM != 0) {
#pragma unroll
                    for (int ch = 0; ch < 2; ++ch) {
                        pw[ch] = aesf_adv(pw[ch], rs.pre[ch].tinc, rs.pre[ch].len);
                        pa[ch] = aesf_adv(pa[ch], rs.pre[ch].tinc, rs.pre[ch].len);
                    }
                }
                // loads for tile i+1: a whole phase (and the next tile's first all-pass walk) ahead of their use
                const bool next_full = rem >= 2 * T;
                pf_ok = fast_in && next_full;
                if (pf_ok) {
                    const float *p = xin + 2 * (n0 + T + i0);
                    pfx0 = aes_ldg_v4(p);
                    pfx1 = aes_ldg_v4(p + 4);
                }
                if constexpr (PRE == AESRV_PRE_DELAY) {
                    const FastStage &ds = a.st[0];
#pragma unroll
                    for (int ch = 0; ch < 2; ++ch) {
                        const FRing rg = ds.ring[ch][0];
                        dw[ch] = aesf_adv(dw[ch], rg.tinc, rg.len);
                        da[ch] = aesf_adv(da[ch], rg.tinc, rg.len);
                        if (pf_ok) {
                            const float *rb = gscr + rg.off;
                            lnA[ch] = aes_ldg_v4(rb + da[ch]);
                            if (((rg.lag + 3) & ~3) != rg.lag) {
                                int b0 = da[ch] + 4;
                                b0 = b0 >= rg.len ? b0 - rg.len : b0;
                                lnB[ch] = aes_ldg_v4(rb + b0);
                            }
                        }
                    }
                }
            }
        }
        if (PRE == AESRV_PRE_BIQUAD && a.state_out != nullptr) {
            __syncthreads();
            // carried scalars of the biquad (stage 0) at the end of the clip: written by the last tile
            // into the parity opposite to its own, i.e. opposite to the drain iteration's too
            if (tid < 8) a.state_out[b * 32 + tid] = bst[(ntiles & 1) * 8 + tid];
        }
        __syncthreads();                                    // the next clip re-initialises rings and carried values
    }
}

#ifndef AES_CPU_EMU
template <int TOPO, int PRE, int PM>
__global__ void __launch_bounds__(AES_NT, 2) aes_rv_kernel(const __grid_constant__ FastArgs a)
{
    aes_rv_body<TOPO, PRE, PM>(a);
}
#endif
