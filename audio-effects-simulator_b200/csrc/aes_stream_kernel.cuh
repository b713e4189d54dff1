// aes_stream_kernel.cuh -- block-streaming path: any block size, state carried between calls.
//
// The whole-clip kernels start every clip from freshly prepared delay lines.  The reference's
// blocks, however, are also driven block by block (the chain's 2 x 1024-frame warm-up,
// core.py:131-136, and the live 256-frame callback, engine.py:156-163) with ring buffers and
// filter state carried from call to call.  This kernel keeps those semantics exactly: rings in
// the reference's own layout (size, write pointer), the reference's per-sample loops in their
// original order and promoted f64 arithmetic (delay.py:7-22, reverb.py:11-67, filter.py:8-40,
// gate.py:6-42, octaver.py:17-82).  It is latency-oriented, not throughput-oriented: one CTA per
// call, one thread per channel/side (threads 0 and 32; a reverb's combs each get their own), stages
// separated by CTA barriers; a 256-frame block through a preset takes 0.1-0.5 ms against a 5.3 ms
// audio period.
//
// Stage descriptors are the aes_stage_desc records of include/aesim.h with three extra fields:
//   q[28] device pointer of the stage's state blob (float*), q[29] frames processed since
//   prepare() (every ring's write pointer is that count modulo its size), q[30] ring size
//   (delay: int(fs*max_delay_ms/1000)+1; reverb: pre-delay ring size).
// Blob layouts:  DELAY  ringL[size] ringR[size]
//                REVERB double lp[2][8] | per side: pre[size] comb_c[L_c+1].. ap_k[L_k+1]..
//                OCTAVER ring[size]
// Scalars (biquad DF-I state, gate gain, octaver phasor) travel in the descriptor and come back
// in state_out[stage*16 ..], like aes_chain_final_state.
#pragma once
#include "aes_plan.h"
#include "../../include/aesim.h"

#define AESS_MAX_FRAMES 2048
#define AESS_SMEM_FLOATS(nf) ((size_t)(nf) * (8 + 2 * AES_MAX_COMB))     // cur 2F | scratch 6F | comb outputs 2*8*F
#define AESS_NT 64

__device__ __forceinline__ long long aess_pymod(long long a, long long m)
{
    long long r = a % m;
    return r < 0 ? r + m : r;
}

__device__ __forceinline__ float aess_mixclip(float dry, float x, float wet, float w)
{
    const float v = __fadd_rn(__fmul_rn(dry, x), __fmul_rn(wet, w));
    return fminf(fmaxf(v, -1.0f), 1.0f);
}

__device__ __forceinline__ double aess_hermite(double t, float y0, float y1, float y2, float y3)
{
    const double c0 = y1;
    const double c1 = 0.5 * (double)__fadd_rn(y2, -y0);
    const double c2 = (double)y0 - 2.5 * y1 + 2.0 * y2 - 0.5 * y3;
    const double c3 = 0.5 * (double)__fadd_rn(y3, -y0) + 1.5 * (double)__fadd_rn(y1, -y2);
    return ((c3 * t + c2) * t + c1) * t + c0;
}

struct StreamArgs {
    const aes_stage_desc *stages;   // device copy
    int n_stages;
    int ci;                         // input channels (1: fan out, 2)
    const float *x;                 // (frames, ci)
    float *y;                       // (frames, 2)
    int frames;                     // <= AESS_MAX_FRAMES
    double *state_out;              // [n_stages][16]
};

__device__ void aes_stream_body(const StreamArgs &a)
{
    AES_DYN_SMEM(float, sm);
    const int tid = threadIdx.x, F = a.frames;
    float *cur = sm;                    // [2][F]
    float *tmp = sm + 2 * F;            // [2 sides][3][F] scratch (pre, t, sum)
    const int side = tid == 0 ? 0 : (tid == 32 ? 1 : -1);

    for (int i = tid; i < F; i += AESS_NT) {
        if (a.ci == 1) { const float v = a.x[i]; cur[i] = v; cur[F + i] = v; }      // core.py:147-149
        else { cur[i] = a.x[2 * i]; cur[F + i] = a.x[2 * i + 1]; }
    }
    __syncthreads();

    for (int s = 0; s < a.n_stages; ++s) {
        const aes_stage_desc &d = a.stages[s];
        double *so = a.state_out + 16 * s;
        float *blob = reinterpret_cast<float *>(d.q[28]);
        const long long n_tot = d.q[29];
        if (d.kind == AES_STAGE_DELAY) {
            if (side >= 0) {
                const long long size = d.q[30], dS = d.q[side];
                float *ring = blob + (long long)side * size;
                long long w = n_tot % size;
                const double fb = d.p[0];
                const float dry = (float)d.p[1], wet = (float)d.p[2];
                float *xc = cur + side * F;
                for (int n = 0; n < F; ++n) {
                    const float dl = ring[aess_pymod(w - dS, size)];
                    const float x = xc[n];
                    ring[w] = (float)((double)x + (double)dl * fb);
                    if (++w == size) w = 0;
                    xc[n] = aess_mixclip(dry, x, wet, dl);
                }
            }
        } else if (d.kind == AES_STAGE_REVERB) {
            // The combs of a side are independent given the pre-delayed input, so they run on their
            // own threads (thread c of warp `side`): pre-delay by the side's first thread, barrier,
            // nc comb loops side by side, barrier, then the first thread sums them IN COMB ORDER
            // (f32, as reverb.py does) and runs the all-passes and the mix.
            const int wside = tid >> 5, wl = tid & 31;              // warp 0: left, warp 1: right
            const int nc = (int)d.q[0], na = (int)d.q[1];
            const long long pre_dS = d.q[2], pre_size = d.q[30];
            double *lps = reinterpret_cast<double *>(blob) + wside * 8;
            long long off0 = 32;                                    // 16 doubles of lp state, then side 0's rings
            for (int sd = 0; sd < wside; ++sd) {
                off0 += pre_size;
                for (int c = 0; c < nc; ++c) off0 += d.q[4 + 8 * sd + c] + 1;
                for (int k = 0; k < na; ++k) off0 += d.q[20 + 4 * sd + k] + 1;
            }
            float *xc = cur + wside * F;
            float *pre = tmp + wside * 3 * F, *t1 = pre + F, *sum = pre + 2 * F;
            float *yc = sm + 8 * F + (size_t)wside * AES_MAX_COMB * F;   // [nc][F] comb outputs of this side
            if (wl == 0) {
                float *ring = blob + off0;
                long long w = n_tot % pre_size;
                for (int n = 0; n < F; ++n) {
                    const float v = xc[n];
                    pre[n] = pre_dS == 0 ? v : ring[aess_pymod(w - pre_dS, pre_size)];
                    ring[w] = v;
                    if (++w == pre_size) w = 0;
                }
            }
            __syncthreads();
            if (wl < nc) {
                const int c = wl;
                long long off = off0 + pre_size;
                for (int c2 = 0; c2 < c; ++c2) off += d.q[4 + 8 * wside + c2] + 1;
                const long long L = d.q[4 + 8 * wside + c], size = L + 1;
                float *ring = blob + off;
                long long w = n_tot % size;
                const double h = d.p[2], g = d.p[4 + 8 * wside + c];
                double lp = lps[c];
                float *yo = yc + (size_t)c * F;
                for (int n = 0; n < F; ++n) {
                    const float yv = ring[aess_pymod(w - L, size)];
                    const double damped = (1.0 - h) * (double)yv + h * lp;
                    lp = damped;
                    ring[w] = (float)((double)pre[n] + g * damped);
                    if (++w == size) w = 0;
                    yo[n] = yv;
                }
                lps[c] = lp;
            }
            __syncthreads();
            if (wl == 0) {
                for (int n = 0; n < F; ++n) {
                    float acc = 0.0f;
                    for (int c = 0; c < nc; ++c) acc = __fadd_rn(acc, yc[(size_t)c * F + n]);
                    sum[n] = acc;
                }
                long long off = off0 + pre_size;
                for (int c = 0; c < nc; ++c) off += d.q[4 + 8 * wside + c] + 1;
                const double ag = d.p[3];
                float *src = sum, *dst = t1;
                for (int k = 0; k < na; ++k) {
                    const long long L = d.q[20 + 4 * wside + k], size = L + 1;
                    float *ring = blob + off;
                    long long w = n_tot % size;
                    for (int n = 0; n < F; ++n) {
                        const float dl = ring[aess_pymod(w - L, size)];
                        const float xi = src[n];
                        const double yo = (double)dl - ag * (double)xi;
                        dst[n] = (float)yo;
                        ring[w] = (float)((double)xi + ag * yo);
                        if (++w == size) w = 0;
                    }
                    float *t = src; src = dst; dst = t;
                    off += size;
                }
                const float dry = (float)d.p[0], wet = (float)d.p[1];
                for (int n = 0; n < F; ++n) xc[n] = aess_mixclip(dry, xc[n], wet, src[n]);
            }
        } else if (d.kind == AES_STAGE_BIQUAD) {
            if (side >= 0) {
                const double b0 = d.p[0], b1 = d.p[1], b2 = d.p[2], a1 = d.p[3], a2 = d.p[4];
                float x1 = (float)d.p[8 + 4 * side], x2 = (float)d.p[9 + 4 * side];
                double y1 = d.p[10 + 4 * side], y2 = d.p[11 + 4 * side];
                float *xc = cur + side * F;
                for (int n = 0; n < F; ++n) {
                    const float x0 = xc[n];
                    const double y0 = b0 * x0 + b1 * x1 + b2 * x2 - a1 * y1 - a2 * y2;
                    xc[n] = (float)y0;
                    x2 = x1; x1 = x0; y2 = y1; y1 = y0;
                }
                so[4 * side + 0] = x1; so[4 * side + 1] = x2; so[4 * side + 2] = y1; so[4 * side + 3] = y2;
            }
        } else if (d.kind == AES_STAGE_GATE) {
            if (tid == 0) {
                const double thr = d.p[0], att = d.p[1], rel = d.p[2];
                double g = d.p[3];
                for (int n = 0; n < F; ++n) {
                    double lvl = 0.0;
                    const float al = fabsf(cur[n]), ar = fabsf(cur[F + n]);
                    if ((double)al > lvl) lvl = al;
                    if ((double)ar > lvl) lvl = ar;
                    const double target = lvl > thr ? 1.0 : 0.0;
                    if (g < target) g = (1.0 - att) * g + att * target;
                    else            g = (1.0 - rel) * g + rel * target;
                    cur[n] = (float)((double)cur[n] * g);
                    cur[F + n] = (float)((double)cur[F + n] * g);
                }
                so[0] = g;
            }
        } else if (d.kind == AES_STAGE_OCTAVER) {
            if (tid == 0) {
                const long long size = d.q[0];
                long long w = d.q[1];
                double ph = d.p[0];
                const double step = d.p[1], fsize = (double)size, two_pi = 6.283185307179586476925;
                const float wet_g = (float)d.p[2], dry_g = (float)(1.0 - d.p[2]);
                float *ring = blob;
                for (int n = 0; n < F; ++n) {
                    const float xl = cur[n], xr = cur[F + n];
                    const float mono = __fmul_rn(__fadd_rn(xl, xr), 0.5f);
                    ring[w] = mono;
                    double p2 = ph + 0.5;
                    if (p2 >= 1.0) p2 -= 1.0;
                    const double raw1 = (double)w - ph * fsize + fsize;
                    const long long k1 = (long long)raw1;
                    const double s1 = aess_hermite(raw1 - (double)k1, ring[aess_pymod(k1 - 1, size)], ring[aess_pymod(k1, size)],
                                                   ring[aess_pymod(k1 + 1, size)], ring[aess_pymod(k1 + 2, size)]);
                    const double raw2 = (double)w - p2 * fsize + fsize;
                    const long long k2 = (long long)raw2;
                    const double s2 = aess_hermite(raw2 - (double)k2, ring[aess_pymod(k2 - 1, size)], ring[aess_pymod(k2, size)],
                                                   ring[aess_pymod(k2 + 1, size)], ring[aess_pymod(k2 + 2, size)]);
                    const double g1 = 0.5 * (1.0 - cos(two_pi * ph)), g2 = 0.5 * (1.0 - cos(two_pi * p2));
                    const float wet = (float)(s1 * g1 + s2 * g2);
                    cur[n] = __fadd_rn(__fmul_rn(xl, dry_g), __fmul_rn(wet, wet_g));
                    cur[F + n] = __fadd_rn(__fmul_rn(xr, dry_g), __fmul_rn(wet, wet_g));
                    if (++w >= size) w = 0;
                    ph += step;
                    if (ph >= 1.0) ph -= 1.0;
                    else if (ph < 0.0) ph += 1.0;
                }
                so[0] = ph;
            }
        } else if (d.kind == AES_STAGE_DISTORTION) {
            const float drive = (float)d.p[0], mix = (float)d.p[1], dry = 1.0f - (float)d.p[1];
            for (int e = tid; e < 2 * F; e += AESS_NT) {
                const float v = cur[e];
                const float t = tanhf(__fmul_rn(drive, v));
                cur[e] = fminf(fmaxf(__fadd_rn(__fmul_rn(dry, v), __fmul_rn(mix, t)), -1.0f), 1.0f);
            }
        }
        __syncthreads();
    }
    for (int i = tid; i < F; i += AESS_NT) { a.y[2 * i] = cur[i]; a.y[2 * i + 1] = cur[F + i]; }
}
