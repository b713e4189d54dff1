// aes_spectral.cuh -- device code of the SpectralFilter block (reference
// src/audioblocks/spectral.py:44-100): rfft of a Hann-windowed frame of M = 2*hop samples ->
// magnitude gate per bin with a temporally smoothed mask -> irfft.  M is whatever the chain's
// block size dictates (whole-file mode: M = 2N, e.g. 960 000 or 1 785 550), so the DFT is
// evaluated with Bluestein's chirp-z identity over power-of-two FFTs of length P >= 2M-1:
//     X[k] = c[k] * sum_n (x[n] c[n]) conj(c[k-n]),      c[n] = exp(-i*pi*n^2/M)
// i.e. one circular convolution with the fixed kernel v[m] = conj(c[|m|]) per transform.
// The inverse real transform is the same machinery on the conjugated Hermitian spectrum.
// The power-of-two FFT is decimation-in-frequency forward (bit-reversed out) and
// decimation-in-time inverse (bit-reversed in): stages whose butterflies span >= 1024 points
// are grid-wide radix-2 passes over global memory, the last/first 10 stages run as independent
// 1024-point transforms in shared memory (aesc_fft_dif<10> / aesc_ifft_dit<10>); the pointwise
// product with FFT(v) happens in bit-reversed order, so no permutation pass exists.
// Two real frames share one complex transform: z = f0 + i*f1 goes through the forward DFT, the
// two spectra are separated by conjugate symmetry (X0[k] = (Z[k] + conj(Z[M-k]))/2,
// X1[k] = (Z[k] - conj(Z[M-k]))/(2i)), gated each with its own mask, recombined as
// W = P0 + i*P1 (both Hermitian) and one inverse DFT returns y0 + i*y1.
#pragma once
#include "aes_convreverb.cuh"

struct SpecArgs {
    cpx *buf;               // [nb][P] work buffer (in place)
    const cpx *vhat;        // [P] FFT(v), bit-reversed order, 1/P folded in
    const cpx *chirp;       // [M] c[n]
    const cpx *twP;         // unused on the device (grid-wide twiddles come from sincospif); the emulator build keeps a table
    const cpx *tw1k;        // [512] exp(-2*pi*i*q/1024)
    const float *frames;    // [nf][M] analysis frames: already windowed, or raw when `window` is set
    const float *window;    // [M] Hann window applied on load (f32 product, as numpy's), or null
    float *mask;            // [nb][M/2+1] smoothed mask, in/out
    float *out;             // [nb][M] irfft of the processed spectrum
    long long M, P;
    int L, nb;              // nb: complex transforms = ceil(nf / 2)
    int nf;                 // real frames (frames / mask / out rows)
    float thr, red, alpha;
};

// u[n] = (f0[n] + i*f1[n]) * c[n] (n < M), zero padding to P; f1 = 0 for an unpaired last frame
__device__ void aess_load_body(const SpecArgs &a)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long total = (long long)a.nb * a.P;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        const long long b = e / a.P, n = e % a.P;
        cpx v; v.x = 0.f; v.y = 0.f;
        if (n < a.M) {
            cpx z;
            z.x = a.frames[2 * b * a.M + n];
            z.y = 2 * b + 1 < a.nf ? a.frames[(2 * b + 1) * a.M + n] : 0.0f;
            if (a.window != nullptr) { const float w = a.window[n]; z.x = __fmul_rn(z.x, w); z.y = __fmul_rn(z.y, w); }
            v = c_mul(z, a.chirp[n]);
        }
        a.buf[e] = v;
    }
}

// exp(-2*pi*i * e / 2^L) without a table: in the grid-wide passes neighbouring threads need
// twiddles 2^st entries apart, so table reads cost up to 32 cache lines per warp load and out-weigh
// the data traffic (measured: 2.3 ms per pass against 0.93 ms for the unit-stride first pass);
// the argument e / 2^(L-1) is exact in f32 and sincospif is accurate to an ulp.
__device__ __forceinline__ cpx aess_twiddle(long long e, int L)
{
    cpx w;
#ifdef AES_CPU_EMU
    const double ang = -M_PI * (double)e / (double)(1LL << (L - 1));
    w.x = (float)cos(ang); w.y = (float)sin(ang);
#else
    sincospif(-(float)e / (float)(1LL << (L - 1)), &w.y, &w.x);
#endif
    return w;
}

template <int R, bool TABLE, typename I>
__device__ __forceinline__ void aess_bfly(cpx (&p)[1 << R], I i, I d, int st, int inverse, int L,
                                          const cpx *__restrict__ tw)
{
    constexpr int K = 1 << R;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int dk = inverse ? (1 << r) : (1 << (R - 1 - r));     // butterfly distance in units of d
        const int sh = inverse ? (L - 1 - (st + r)) : (st + r);     // twiddle exponent shift of this stage
#pragma unroll
        for (int k = 0; k < K; ++k) {
            if (k & dk) continue;
            const I pos = (I)(k & (dk - 1)) * d + i;
            cpx w = TABLE ? tw[pos << sh] : aess_twiddle((long long)pos << sh, L);
            const cpx u = p[k], v = p[k + dk];
            if (!inverse) {
                p[k] = c_add(u, v);
                p[k + dk] = c_mul(c_sub(u, v), w);
            } else {
                w.y = -w.y;
                const cpx t = c_mul(v, w);
                p[k] = c_add(u, t);
                p[k + dk] = c_sub(u, t);
            }
        }
    }
}

// R (1..3) consecutive grid-wide stages in one pass over global memory: a thread owns the 2^R
// points its butterflies connect, so the data makes one round trip per pass instead of one per
// stage (11 global stages of a 2^21-point transform: 4 passes).  Forward (DIF) passes start at
// stage `st` (largest distance first); inverse (DIT) passes start at stage `st` and climb.
// IN : 0 read the work buffer | 1 first pass of the analysis round: (f0 + i*f1)*chirp straight from
//      the frames, zeros beyond M (replaces aess_load_body) | 2 first pass of the synthesis round:
//      the work buffer below M, zeros beyond (replaces aess_zero_pad_body)
// OUT: 0 write the work buffer | 1 last pass of the synthesis round: conj(chirp*Y)/M straight to
//      the two output frames (replaces aess_store_body)
template <int R, int IN, int OUT>
__device__ void aess_global_pass_body(const SpecArgs &a, int st, int inverse)
{
    constexpr int K = 1 << R;
    const long long per = a.P >> R;                             // threads' worth of work per transform
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long total = (long long)a.nb * per;
    const long long d = inverse ? (1LL << st) : (a.P >> (st + R));      // distance between a thread's points
    const float inv = 1.0f / (float)a.M;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        const long long b = e / per, rem = e % per;
        const long long i = rem & (d - 1), blk = rem >> (inverse ? st : (a.L - st - R));
        const long long n0 = blk * (d << R) + i;                 // index of point 0 inside the transform
        cpx *p0 = a.buf + b * a.P + n0;
        cpx p[K];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const long long n = n0 + k * d;
            if (IN == 0) {
                p[k] = p0[k * d];
            } else {
                cpx v; v.x = 0.f; v.y = 0.f;
                if (n < a.M) {
                    if (IN == 1) {
                        cpx z;
                        z.x = a.frames[2 * b * a.M + n];
                        z.y = 2 * b + 1 < a.nf ? a.frames[(2 * b + 1) * a.M + n] : 0.0f;
                        if (a.window != nullptr) { const float w = a.window[n]; z.x = __fmul_rn(z.x, w); z.y = __fmul_rn(z.y, w); }
                        v = c_mul(z, a.chirp[n]);
                    } else {
                        v = p0[k * d];
                    }
                }
                p[k] = v;
            }
        }
        aess_bfly<R, false, long long>(p, i, d, st, inverse, a.L, a.twP);
#pragma unroll
        for (int k = 0; k < K; ++k) {
            if (OUT == 0) {
                p0[k * d] = p[k];
            } else {
                const long long n = n0 + k * d;
                if (n < a.M) {
                    const cpx y = c_mul(a.chirp[n], p[k]);
                    a.out[2 * b * a.M + n] = y.x * inv;
                    if (2 * b + 1 < a.nf) a.out[(2 * b + 1) * a.M + n] = -y.y * inv;
                }
            }
        }
    }
}
// how many stages the next global pass takes when `remaining` are left
__host__ __device__ inline int aess_pass_radix(int remaining) { return remaining >= 3 ? 3 : remaining; }

// The 10 stages that stay inside aligned 1024-point chunks.  128 threads per chunk, 8 points per
// thread, two chunks per 256-thread CTA, persistent over the chunks: three radix-8 passes and one
// radix-2 pass, the outer passes straight from / to global memory (coalesced), the inner ones
// through shared memory (index padded by 2 per 16 so the stride-16 and stride-2 passes stay at
// the 2-wavefront minimum).  With `mul` the forward transform is multiplied by FFT(v) on the way
// out (bit-reversed order on both sides).
#define AESS_LOCAL_NT 256
#define AESS_LOCAL_SMEM_CPX (2 * 1152)
__device__ __forceinline__ int aess_pad(int idx) { return idx + 2 * (idx >> 4); }

__device__ void aess_local_body(const SpecArgs &a, int inverse, int mul)
{
    AES_DYN_SMEM(cpx, smem);
    const int g = threadIdx.x >> 7, t = threadIdx.x & 127;
    cpx *s = smem + g * 1152;
    const long long nchunks = (long long)a.nb * a.P / 1024;
    const cpx *tw = a.tw1k;
    // the chunk's 8 points per thread are fetched one trip ahead (registers q), so the global
    // latency hides behind the three shared-memory passes of the current chunk
    cpx q[8];
    auto fetch = [&](long long cpn) {
        const long long chn = 2 * cpn + g;
        if (chn >= nchunks) return;
        const cpx *dn = a.buf + chn * 1024;
        if (!inverse) {
#pragma unroll
            for (int k = 0; k < 8; ++k) q[k] = dn[t + 128 * k];
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) { const int e = 2 * (t + 128 * k); q[2 * k] = dn[e]; q[2 * k + 1] = dn[e + 1]; }
        }
    };
    if (2 * (long long)blockIdx.x < nchunks) fetch(blockIdx.x);
    for (long long cp = blockIdx.x; 2 * cp < nchunks; cp += gridDim.x) {
        const long long chunk = 2 * cp + g;
        const bool valid = chunk < nchunks;
        cpx *d = a.buf + chunk * 1024;
        const long long off = (chunk * 1024) % a.P;
        cpx p[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) p[k] = q[k];
        if (2 * (cp + gridDim.x) < nchunks) fetch(cp + gridDim.x);
        if (!inverse) {
            if (valid) {                                        // stages 0-2: points t + 128k (fetched from global)
                aess_bfly<3, true, int>(p, t, 128, 0, 0, 10, tw);
#pragma unroll
                for (int k = 0; k < 8; ++k) s[aess_pad(t + 128 * k)] = p[k];
            }
            __syncthreads();
            if (valid) {                                        // stages 3-5: points 128*hi + lo + 16k
                const int i = t & 15, base = (t >> 4) * 128 + i;
#pragma unroll
                for (int k = 0; k < 8; ++k) p[k] = s[aess_pad(base + 16 * k)];
                aess_bfly<3, true, int>(p, i, 16, 3, 0, 10, tw);
#pragma unroll
                for (int k = 0; k < 8; ++k) s[aess_pad(base + 16 * k)] = p[k];
            }
            __syncthreads();
            if (valid) {                                        // stages 6-8: points 16*hi + lo + 2k
                const int i = t & 1, base = (t >> 1) * 16 + i;
#pragma unroll
                for (int k = 0; k < 8; ++k) p[k] = s[aess_pad(base + 2 * k)];
                aess_bfly<3, true, int>(p, i, 2, 6, 0, 10, tw);
#pragma unroll
                for (int k = 0; k < 8; ++k) s[aess_pad(base + 2 * k)] = p[k];
            }
            __syncthreads();
            if (valid) {                                        // stage 9 (twiddle 1): pairs (2q, 2q+1), to global
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int e = 2 * (t + 128 * k);
                    const cpx u = s[aess_pad(e)], v = s[aess_pad(e) + 1];
                    cpx o0 = c_add(u, v), o1 = c_sub(u, v);
                    if (mul) { o0 = c_mul(o0, a.vhat[off + e]); o1 = c_mul(o1, a.vhat[off + e + 1]); }
                    d[e] = o0; d[e + 1] = o1;
                }
            }
        } else {
            if (valid) {                                        // stage 0: pairs (fetched from global)
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int e = 2 * (t + 128 * k);
                    const cpx u = p[2 * k], v = p[2 * k + 1];
                    s[aess_pad(e)] = c_add(u, v); s[aess_pad(e) + 1] = c_sub(u, v);
                }
            }
            __syncthreads();
            if (valid) {                                        // stages 1-3: distances 2, 4, 8
                const int i = t & 1, base = (t >> 1) * 16 + i;
#pragma unroll
                for (int k = 0; k < 8; ++k) p[k] = s[aess_pad(base + 2 * k)];
                aess_bfly<3, true, int>(p, i, 2, 1, 1, 10, tw);
#pragma unroll
                for (int k = 0; k < 8; ++k) s[aess_pad(base + 2 * k)] = p[k];
            }
            __syncthreads();
            if (valid) {                                        // stages 4-6: distances 16, 32, 64
                const int i = t & 15, base = (t >> 4) * 128 + i;
#pragma unroll
                for (int k = 0; k < 8; ++k) p[k] = s[aess_pad(base + 16 * k)];
                aess_bfly<3, true, int>(p, i, 16, 4, 1, 10, tw);
#pragma unroll
                for (int k = 0; k < 8; ++k) s[aess_pad(base + 16 * k)] = p[k];
            }
            __syncthreads();
            if (valid) {                                        // stages 7-9: distances 128, 256, 512, to global
#pragma unroll
                for (int k = 0; k < 8; ++k) p[k] = s[aess_pad(t + 128 * k)];
                aess_bfly<3, true, int>(p, t, 128, 7, 1, 10, tw);
#pragma unroll
                for (int k = 0; k < 8; ++k) d[t + 128 * k] = p[k];
            }
        }
        __syncthreads();                                        // the next chunk reuses the buffer
    }
}

// forward result -> the two frames' spectra -> spectral gate on each -> conjugated combined
// spectrum times chirp, ready for the second (inverse) Bluestein round.  One thread per pair and
// bin k in [0, M/2]; it alone touches elements k and M-k of the pair's buffer.
__device__ __forceinline__ cpx aess_gate_one(cpx X, float *maskp, long long k, const SpecArgs &a)
{
    const float mag = sqrtf(X.x * X.x + X.y * X.y);
    const float cur = mag > a.thr ? 1.0f : a.red;               // spectral.py:68
    const float m = a.alpha * *maskp + (1.0f - a.alpha) * cur;  // spectral.py:71
    *maskp = m;
    cpx Pk; Pk.x = X.x * m; Pk.y = X.y * m;                     // mag*mask*exp(i*phase)
    if (k == 0 || 2 * k == a.M) Pk.y = 0.f;                     // irfft ignores the imaginary part of DC / Nyquist
    return Pk;
}

__device__ void aess_gate_body(const SpecArgs &a)
{
    const long long nbins = a.M / 2 + 1;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < (long long)a.nb * nbins; e += stride) {
        const long long b = e / nbins, k = e % nbins;
        const long long km = k == 0 ? 0 : a.M - k;
        const bool two = 2 * b + 1 < a.nf;
        cpx *d = a.buf + b * a.P;
        const cpx Zk = c_mul(a.chirp[k], d[k]), Zm = c_mul(a.chirp[km], d[km]);
        cpx X0, X1;                                             // rfft bin k of frame 2b / 2b+1
        X0.x = 0.5f * (Zk.x + Zm.x); X0.y = 0.5f * (Zk.y - Zm.y);
        X1.x = 0.5f * (Zk.y + Zm.y); X1.y = 0.5f * (Zm.x - Zk.x);
        const cpx P0 = aess_gate_one(X0, a.mask + (2 * b) * nbins + k, k, a);
        cpx P1; P1.x = 0.f; P1.y = 0.f;
        if (two) P1 = aess_gate_one(X1, a.mask + (2 * b + 1) * nbins + k, k, a);
        // inverse DFT via a forward one: y0 + i*y1 = conj(DFT(conj(W))) / M with W = P0full + i*P1full,
        // W[k] = P0 + i*P1, W[M-k] = conj(P0) + i*conj(P1)
        cpx ck, cm;                                             // conj(W[k]), conj(W[M-k])
        ck.x = P0.x - P1.y; ck.y = -P0.y - P1.x;
        cm.x = P0.x + P1.y; cm.y = P0.y - P1.x;
        d[k] = c_mul(ck, a.chirp[k]);
        if (k != 0 && 2 * k != a.M) d[km] = c_mul(cm, a.chirp[km]);
    }
}

// zero the padding [M, P) again before the second convolution round
__device__ void aess_zero_pad_body(const SpecArgs &a)
{
    const long long pad = a.P - a.M;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < (long long)a.nb * pad; e += stride) {
        const long long b = e / pad, n = a.M + e % pad;
        cpx z; z.x = 0.f; z.y = 0.f;
        a.buf[b * a.P + n] = z;
    }
}

// y0[n] + i*y1[n] = conj(c[n]*Y[n]) / M
__device__ void aess_store_body(const SpecArgs &a)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    const float inv = 1.0f / (float)a.M;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < (long long)a.nb * a.M; e += stride) {
        const long long b = e / a.M, n = e % a.M;
        const cpx y = c_mul(a.chirp[n], a.buf[b * a.P + n]);
        a.out[2 * b * a.M + n] = y.x * inv;
        if (2 * b + 1 < a.nf) a.out[(2 * b + 1) * a.M + n] = -y.y * inv;
    }
}
