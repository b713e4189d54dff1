// aes_convreverb.cuh -- IR-convolution reverb (BASELINE configs[3]); device code.
//
// The reference has no convolution reverb (its reverb is the Schroeder network,
// reverb.py:72-277): this is a NEW operator in the style of the other blocks,
//     out[:, c] = clip(mix_dry * x[:, c] + mix_wet * (x[:, c] * h[:, c]), -1, 1),
// with `*` the causal linear convolution with a stereo impulse response h (n_taps, 2).
// Parity is checked against a float64 numpy restatement in the test tree (parity unpinned by the
// reference).
//
// Uniformly partitioned overlap-save in the frequency domain, blocks of BK = N/2 frames and
// FFTs of N = 2^LOG2N points that live entirely in shared memory:
//   K1  fft_blocks : z = xL + i*xR -- an interleaved stereo frame IS a complex number -- so ONE
//                    complex FFT per block carries both channels;  Z[clip][j][.] (bit-reversed)
//   K2  mac        : W_j = sum_p  A_p * Z_{j-p} + B_p * conj(Z_{j-p}[N-k])   per bin,
//                    A_p = (HL_p + HR_p)/2, B_p = (HL_p - HR_p)/2 fold the two real IRs back into
//                    the packed spectrum; the IR partitions of a 256-bin tile are staged once
//                    per CTA in shared memory and reused for every clip and block of the CTA
//   K3  ifft_mix   : inverse FFT, keep the last BK samples (overlap-save), real = yL, imag = yR,
//                    dry/wet mix and clip, written as stereo frames.
// Forward transforms are decimation-in-frequency (natural in, bit-reversed out), the inverse
// is decimation-in-time (bit-reversed in, natural out); the MAC works in bit-reversed index
// space, where the mirror bin N-k of physical index i is a reversal inside i's power-of-two
// band: i' = 3*2^m - 1 - i for 2^m <= i < 2^(m+1) (i' = i for i < 2) -- contiguous, coalesced.
#pragma once
#include "aes_plan.h"

#define AESC_NT 1024
#define AESC_KT 256         // bins per MAC tile
#define AESC_JB 8           // output blocks per MAC chunk

struct alignas(8) cpx { float x, y; };      // one 64-bit load / store per element

__device__ __forceinline__ cpx c_mul(cpx a, cpx b) { cpx r; r.x = a.x * b.x - a.y * b.y; r.y = a.x * b.y + a.y * b.x; return r; }

#ifdef AES_CPU_EMU
static inline unsigned __brev(unsigned v)
{
    unsigned r = 0;
    for (int i = 0; i < 32; ++i) r |= ((v >> i) & 1u) << (31 - i);
    return r;
}
#endif

// mirror bin (N - k) mod N in bit-reversed index space
__device__ __forceinline__ int aesc_mirror(int i)
{
    if (i < 2) return i;
    const int m = 31 - __clz(i);
    return 3 * (1 << m) - 1 - i;
}

// In-place FFTs on N = 2^L points in shared memory; tw[q] = exp(-2*pi*i*q/N), q < N/2.
// Three radix-2 stages are fused per pass (8 elements in registers, one table twiddle per
// thread and its squares, the rest are the constants of W_8), a radix-4 or radix-2 pass takes
// the remainder: 14 stages = 5 shared-memory round trips and barriers instead of 14.
__device__ __forceinline__ cpx c_add(cpx a, cpx b) { cpx r; r.x = a.x + b.x; r.y = a.y + b.y; return r; }
__device__ __forceinline__ cpx c_sub(cpx a, cpx b) { cpx r; r.x = a.x - b.x; r.y = a.y - b.y; return r; }
__device__ __forceinline__ cpx c_sq(cpx a) { cpx r; r.x = a.x * a.x - a.y * a.y; r.y = 2.0f * a.x * a.y; return r; }
__device__ __forceinline__ cpx c_mul_mi(cpx a) { cpx r; r.x = a.y; r.y = -a.x; return r; }     // a * (-i)
__device__ __forceinline__ cpx c_mul_pi(cpx a) { cpx r; r.x = -a.y; r.y = a.x; return r; }     // a * (+i)
#define AESC_R2 0.70710678118654752f

// forward, decimation in frequency: natural order in, bit-reversed order out
template <int L>
__device__ void aesc_fft_dif(cpx *s, const cpx *__restrict__ tw, int tid)
{
    constexpr int N = 1 << L;
    int st = 0;
    for (; L - st >= 3; st += 3) {
        const int q = N >> (st + 3);                     // distances 4q, 2q, q
        for (int it = tid; it < N / 8; it += AESC_NT) {
            const int r = it & (q - 1);
            const int base = ((it - r) << 3) + r;
            cpx v[8];
#pragma unroll
            for (int m = 0; m < 8; ++m) v[m] = s[base + m * q];
            const cpx w1 = tw[r << st], w2 = c_sq(w1), w4 = c_sq(w2);
            // stage st (distance 4q): twiddle of pair m is w1 * W_8^m
            {
                const cpx d0 = c_sub(v[0], v[4]), d1 = c_sub(v[1], v[5]), d2 = c_sub(v[2], v[6]), d3 = c_sub(v[3], v[7]);
                v[0] = c_add(v[0], v[4]); v[1] = c_add(v[1], v[5]); v[2] = c_add(v[2], v[6]); v[3] = c_add(v[3], v[7]);
                cpx t1; t1.x = AESC_R2 * (d1.x + d1.y); t1.y = AESC_R2 * (d1.y - d1.x);          // d1 * (1-i)/sqrt2
                cpx t3; t3.x = AESC_R2 * (d3.y - d3.x); t3.y = -AESC_R2 * (d3.x + d3.y);         // d3 * (-1-i)/sqrt2
                v[4] = c_mul(d0, w1); v[5] = c_mul(t1, w1); v[6] = c_mul(c_mul_mi(d2), w1); v[7] = c_mul(t3, w1);
            }
            // stage st+1 (distance 2q): twiddle w2 * (-i)^(m&1)
#pragma unroll
            for (int g = 0; g < 8; g += 4) {
                const cpx d0 = c_sub(v[g], v[g + 2]), d1 = c_sub(v[g + 1], v[g + 3]);
                v[g] = c_add(v[g], v[g + 2]); v[g + 1] = c_add(v[g + 1], v[g + 3]);
                v[g + 2] = c_mul(d0, w2); v[g + 3] = c_mul(c_mul_mi(d1), w2);
            }
            // stage st+2 (distance q): twiddle w4
#pragma unroll
            for (int g = 0; g < 8; g += 2) {
                const cpx d = c_sub(v[g], v[g + 1]);
                v[g] = c_add(v[g], v[g + 1]);
                v[g + 1] = c_mul(d, w4);
            }
#pragma unroll
            for (int m = 0; m < 8; ++m) s[base + m * q] = v[m];
        }
        __syncthreads();
    }
    for (; st < L; ++st) {                                // remaining 1 or 2 plain radix-2 stages
        const int half = N >> (st + 1);
        for (int b = tid; b < N / 2; b += AESC_NT) {
            const int pos = b & (half - 1);
            const int i0 = ((b - pos) << 1) + pos, i1 = i0 + half;
            const cpx u = s[i0], v = s[i1], w = tw[pos << st];
            s[i0] = c_add(u, v);
            s[i1] = c_mul(c_sub(u, v), w);
        }
        __syncthreads();
    }
}

// inverse, decimation in time: bit-reversed order in, natural order out (unnormalised)
template <int L>
__device__ void aesc_ifft_dit(cpx *s, const cpx *__restrict__ tw, int tid)
{
    constexpr int N = 1 << L;
    int st = 0;
    for (; st < L % 3; ++st) {                            // leading 1 or 2 plain radix-2 stages
        const int half = 1 << st;
        for (int b = tid; b < N / 2; b += AESC_NT) {
            const int pos = b & (half - 1);
            const int i0 = ((b - pos) << 1) + pos, i1 = i0 + half;
            cpx w = tw[pos << (L - 1 - st)];
            w.y = -w.y;
            const cpx u = s[i0], t = c_mul(s[i1], w);
            s[i0] = c_add(u, t);
            s[i1] = c_sub(u, t);
        }
        __syncthreads();
    }
    for (; st < L; st += 3) {
        const int h = 1 << st;                            // distances h, 2h, 4h
        for (int it = tid; it < N / 8; it += AESC_NT) {
            const int r = it & (h - 1);
            const int base = ((it - r) << 3) + r;
            cpx v[8];
#pragma unroll
            for (int m = 0; m < 8; ++m) v[m] = s[base + m * h];
            cpx v4 = tw[r << (L - 3 - st)];               // exponent r*N/(8h)
            v4.y = -v4.y;
            const cpx v2 = c_sq(v4), v1 = c_sq(v2);
            // stage st (distance h): twiddle v1
#pragma unroll
            for (int g = 0; g < 8; g += 2) {
                const cpx t = c_mul(v[g + 1], v1);
                v[g + 1] = c_sub(v[g], t);
                v[g] = c_add(v[g], t);
            }
            // stage st+1 (distance 2h): twiddle v2 * (+i)^(m&1)
#pragma unroll
            for (int g = 0; g < 8; g += 4) {
                const cpx t0 = c_mul(v[g + 2], v2), t1 = c_mul_pi(c_mul(v[g + 3], v2));
                v[g + 2] = c_sub(v[g], t0); v[g] = c_add(v[g], t0);
                v[g + 3] = c_sub(v[g + 1], t1); v[g + 1] = c_add(v[g + 1], t1);
            }
            // stage st+2 (distance 4h): twiddle v4 * conj(W_8)^m
            {
                const cpx a0 = c_mul(v[4], v4), a1 = c_mul(v[5], v4), a2 = c_mul(v[6], v4), a3 = c_mul(v[7], v4);
                cpx t1; t1.x = AESC_R2 * (a1.x - a1.y); t1.y = AESC_R2 * (a1.x + a1.y);          // a1 * (1+i)/sqrt2
                const cpx t2 = c_mul_pi(a2);
                cpx t3; t3.x = -AESC_R2 * (a3.x + a3.y); t3.y = AESC_R2 * (a3.x - a3.y);         // a3 * (-1+i)/sqrt2
                v[4] = c_sub(v[0], a0); v[0] = c_add(v[0], a0);
                v[5] = c_sub(v[1], t1); v[1] = c_add(v[1], t1);
                v[6] = c_sub(v[2], t2); v[2] = c_add(v[2], t2);
                v[7] = c_sub(v[3], t3); v[3] = c_add(v[3], t3);
            }
#pragma unroll
            for (int m = 0; m < 8; ++m) s[base + m * h] = v[m];
        }
        __syncthreads();
    }
}

struct ConvArgs {
    const float *x;         // (B, Nf, 2) f32
    float *y;               // (B, Nf, 2) f32
    cpx *Z, *W;             // [B][nblk][N] spectra (bit-reversed order)
    const cpx *A, *Bc;      // [P][N] IR partition spectra (bit-reversed order), 1/N folded in
    const cpx *tw;          // [N/2]
    long long B, Nf;
    int nblk, P;
    float dry, wet;
};

// K1: forward FFT of the overlap-save frame [(j-1)*BK, (j+1)*BK) of every clip
template <int L>
__device__ void aesc_fft_blocks_body(const ConvArgs &a)
{
    constexpr int N = 1 << L, BK = N / 2;
    AES_DYN_SMEM(cpx, s);
    const int tid = threadIdx.x;
    const long long blk = blockIdx.x;
    const long long clip = blk / a.nblk;
    const int j = (int)(blk % a.nblk);
    const cpx *xc = reinterpret_cast<const cpx *>(a.x) + clip * a.Nf;
    const long long f0 = (long long)(j - 1) * BK;
    for (int i = tid; i < N; i += AESC_NT) {
        const long long f = f0 + i;
        cpx v; v.x = 0.f; v.y = 0.f;
        if (f >= 0 && f < a.Nf) v = xc[f];
        s[i] = v;
    }
    __syncthreads();
    aesc_fft_dif<L>(s, a.tw, tid);
    cpx *zo = a.Z + ((size_t)clip * a.nblk + j) * N;
    for (int i = tid; i < N; i += AESC_NT) zo[i] = s[i];
}

// K2: per-bin FIR over the block index with the IR partition spectra
template <int L>
__device__ void aesc_mac_body(const ConvArgs &a, int clips_per_cta)
{
    constexpr int N = 1 << L;
    AES_DYN_SMEM(cpx, s);                               // A tile [P][KT] | B tile [P][KT]
    const int tid = threadIdx.x;                        // AESC_KT threads
    const int ktile = blockIdx.x % (N / AESC_KT);
    const long long cgrp = blockIdx.x / (N / AESC_KT);
    const int i = ktile * AESC_KT + tid;                // physical (bit-reversed) bin of this thread
    const int im = aesc_mirror(i);
    cpx *sA = s, *sB = s + (size_t)a.P * AESC_KT;
    for (int p = 0; p < a.P; ++p) {
        sA[p * AESC_KT + tid] = a.A[(size_t)p * N + i];
        sB[p * AESC_KT + tid] = a.Bc[(size_t)p * N + i];
    }
    __syncthreads();
    for (long long clip = cgrp * clips_per_cta; clip < a.B && clip < (cgrp + 1) * clips_per_cta; ++clip) {
        const cpx *zc = a.Z + (size_t)clip * a.nblk * N;
        cpx *wc = a.W + (size_t)clip * a.nblk * N;
        for (int j0 = 0; j0 < a.nblk; j0 += AESC_JB) {
            cpx acc[AESC_JB];
#pragma unroll
            for (int u = 0; u < AESC_JB; ++u) { acc[u].x = 0.f; acc[u].y = 0.f; }
            const int qlo = j0 - a.P + 1 < 0 ? 0 : j0 - a.P + 1;
            const int qhi = j0 + AESC_JB - 1 < a.nblk - 1 ? j0 + AESC_JB - 1 : a.nblk - 1;
            for (int q = qlo; q <= qhi; ++q) {
                const cpx z = zc[(size_t)q * N + i];
                cpx zm = zc[(size_t)q * N + im];
                zm.y = -zm.y;
                const int p0 = j0 - q;                          // partition that maps block q to output j0
                if (p0 >= 0 && p0 + AESC_JB <= a.P) {
                    // interior of the window: all JB outputs take this block, partitions p0 .. p0+JB-1
                    // at constant offsets -- no per-output bounds test or index arithmetic
                    const cpx *pa = sA + p0 * AESC_KT + tid, *pb = sB + p0 * AESC_KT + tid;
#pragma unroll
                    for (int u = 0; u < AESC_JB; ++u) {
                        const cpx ca = pa[u * AESC_KT], cb = pb[u * AESC_KT];
                        acc[u].x += ca.x * z.x - ca.y * z.y + cb.x * zm.x - cb.y * zm.y;
                        acc[u].y += ca.x * z.y + ca.y * z.x + cb.x * zm.y + cb.y * zm.x;
                    }
                } else {
#pragma unroll
                    for (int u = 0; u < AESC_JB; ++u) {
                        const int p = p0 + u;
                        if (p >= 0 && p < a.P) {
                            const cpx ca = sA[p * AESC_KT + tid], cb = sB[p * AESC_KT + tid];
                            acc[u].x += ca.x * z.x - ca.y * z.y + cb.x * zm.x - cb.y * zm.y;
                            acc[u].y += ca.x * z.y + ca.y * z.x + cb.x * zm.y + cb.y * zm.x;
                        }
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < AESC_JB; ++u)
                if (j0 + u < a.nblk) wc[(size_t)(j0 + u) * N + i] = acc[u];
        }
    }
}

// K3: inverse FFT, overlap-save (keep the last BK samples), dry/wet mix and clip
template <int L>
__device__ void aesc_ifft_mix_body(const ConvArgs &a)
{
    constexpr int N = 1 << L, BK = N / 2;
    AES_DYN_SMEM(cpx, s);
    const int tid = threadIdx.x;
    const long long blk = blockIdx.x;
    const long long clip = blk / a.nblk;
    const int j = (int)(blk % a.nblk);
    const cpx *wi = a.W + ((size_t)clip * a.nblk + j) * N;
    for (int i = tid; i < N; i += AESC_NT) s[i] = wi[i];
    __syncthreads();
    aesc_ifft_dit<L>(s, a.tw, tid);
    const cpx *xc = reinterpret_cast<const cpx *>(a.x) + clip * a.Nf;
    cpx *yc = reinterpret_cast<cpx *>(a.y) + clip * a.Nf;
    const long long f0 = (long long)j * BK;
    for (int i = tid; i < BK; i += AESC_NT) {
        const long long f = f0 + i;
        if (f < a.Nf) {
            const cpx dryv = xc[f], wetv = s[BK + i];
            cpx o;
            o.x = fminf(fmaxf(__fadd_rn(__fmul_rn(a.dry, dryv.x), __fmul_rn(a.wet, wetv.x)), -1.0f), 1.0f);
            o.y = fminf(fmaxf(__fadd_rn(__fmul_rn(a.dry, dryv.y), __fmul_rn(a.wet, wetv.y)), -1.0f), 1.0f);
            yc[f] = o;
        }
    }
}

// IR preparation: FFT of partition p of (hL + i*hR), then A = (HL+HR)/2/N, B = (HL-HR)/2/N
// with HL[k] = (H[k] + conj(H[N-k]))/2, HR[k] = (H[k] - conj(H[N-k]))/(2i)
template <int L>
__device__ void aesc_ir_prep_body(const float *ir, int n_taps, cpx *A, cpx *Bc, const cpx *tw)
{
    constexpr int N = 1 << L, BK = N / 2;
    AES_DYN_SMEM(cpx, s);
    const int tid = threadIdx.x, p = blockIdx.x;
    for (int i = tid; i < N; i += AESC_NT) {
        const long long t = (long long)p * BK + i;
        cpx v; v.x = 0.f; v.y = 0.f;
        if (i < BK && t < n_taps) { v.x = ir[2 * t]; v.y = ir[2 * t + 1]; }
        s[i] = v;
    }
    __syncthreads();
    aesc_fft_dif<L>(s, tw, tid);
    const float sc = 1.0f / (float)N;
    for (int i = tid; i < N; i += AESC_NT) {
        const cpx h = s[i];
        cpx hm = s[aesc_mirror(i)];
        hm.y = -hm.y;
        cpx hl, hr;
        hl.x = 0.5f * (h.x + hm.x); hl.y = 0.5f * (h.y + hm.y);
        // (h - hm) / (2i) = (-i/2) * (h - hm) = ( (h.y - hm.y)/2 , -(h.x - hm.x)/2 )
        hr.x = 0.5f * (h.y - hm.y); hr.y = -0.5f * (h.x - hm.x);
        cpx av, bv;
        av.x = 0.5f * sc * (hl.x + hr.x); av.y = 0.5f * sc * (hl.y + hr.y);
        bv.x = 0.5f * sc * (hl.x - hr.x); bv.y = 0.5f * sc * (hl.y - hr.y);
        A[(size_t)p * N + i] = av;
        Bc[(size_t)p * N + i] = bv;
    }
}
