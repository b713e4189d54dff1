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
#pragma once
#include "aes_convreverb.cuh"

struct SpecArgs {
    cpx *buf;               // [nb][P] work buffer (in place)
    const cpx *vhat;        // [P] FFT(v), bit-reversed order, 1/P folded in
    const cpx *chirp;       // [M] c[n]
    const cpx *twP;         // [P/2] exp(-2*pi*i*q/P)
    const cpx *tw1k;        // [512] exp(-2*pi*i*q/1024)
    const float *frames;    // [nb][M] windowed analysis frames
    float *mask;            // [nb][M/2+1] smoothed mask, in/out
    float *out;             // [nb][M] irfft of the processed spectrum
    long long M, P;
    int L, nb;
    float thr, red, alpha;
};

// u[n] = frame[n]*c[n] (n < M), zero padding to P
__device__ void aess_load_body(const SpecArgs &a)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long total = (long long)a.nb * a.P;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        const long long b = e / a.P, n = e % a.P;
        cpx v; v.x = 0.f; v.y = 0.f;
        if (n < a.M) {
            const float x = a.frames[b * a.M + n];
            const cpx c = a.chirp[n];
            v.x = x * c.x; v.y = x * c.y;
        }
        a.buf[e] = v;
    }
}

// one grid-wide radix-2 stage (butterfly distance `half` >= 1024); inverse uses conjugate twiddles
__device__ void aess_global_stage_body(const SpecArgs &a, int st, int inverse)
{
    const long long halfP = a.P >> 1;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long total = (long long)a.nb * halfP;
    // forward DIF stage st: half = P >> (st+1), twiddle exponent pos << st
    // inverse DIT stage st: half = 1 << st,     twiddle exponent pos << (L-1-st), conjugated
    const long long half = inverse ? (1LL << st) : (a.P >> (st + 1));
    const int sh = inverse ? (a.L - 1 - st) : st;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        const long long b = e / halfP, bf = e % halfP;
        const long long pos = bf & (half - 1);
        const long long i0 = ((bf - pos) << 1) + pos, i1 = i0 + half;
        cpx *d = a.buf + b * a.P;
        cpx w = a.twP[pos << sh];
        const cpx u = d[i0], v = d[i1];
        if (!inverse) {
            d[i0] = c_add(u, v);
            d[i1] = c_mul(c_sub(u, v), w);
        } else {
            w.y = -w.y;
            const cpx t = c_mul(v, w);
            d[i0] = c_add(u, t);
            d[i1] = c_sub(u, t);
        }
    }
}

// the 10 stages that stay inside aligned 1024-point chunks, in shared memory; when `mul` is set
// the forward pass also multiplies by FFT(v) on the way out (bit-reversed order on both sides)
__device__ void aess_local_body(const SpecArgs &a, int inverse, int mul)
{
    AES_DYN_SMEM(cpx, s);
    const int tid = threadIdx.x;
    const long long chunk = blockIdx.x;                         // nb * P / 1024 chunks
    cpx *d = a.buf + chunk * 1024;
    const long long off = (chunk * 1024) % a.P;
    s[tid] = d[tid];
    __syncthreads();
    if (!inverse) aesc_fft_dif<10>(s, a.tw1k, tid); else aesc_ifft_dit<10>(s, a.tw1k, tid);
    cpx v = s[tid];
    if (mul) v = c_mul(v, a.vhat[off + tid]);
    d[tid] = v;
}

// forward result -> spectral gate -> conjugated Hermitian spectrum times chirp, ready for the
// second (inverse) Bluestein round.  One thread per bin k in [0, M/2].
__device__ void aess_gate_body(const SpecArgs &a)
{
    const long long nbins = a.M / 2 + 1;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < (long long)a.nb * nbins; e += stride) {
        const long long b = e / nbins, k = e % nbins;
        cpx *d = a.buf + b * a.P;
        const cpx X = c_mul(a.chirp[k], d[k]);                  // rfft bin k
        const float mag = sqrtf(X.x * X.x + X.y * X.y);
        const float cur = mag > a.thr ? 1.0f : a.red;           // spectral.py:68
        const float m = a.alpha * a.mask[e] + (1.0f - a.alpha) * cur;      // spectral.py:71
        a.mask[e] = m;
        cpx Pk; Pk.x = X.x * m; Pk.y = X.y * m;                 // mag*mask*exp(i*phase)
        if (k == 0 || 2 * k == a.M) Pk.y = 0.f;                 // irfft ignores the imaginary part of DC / Nyquist
        // inverse DFT via forward DFT: y = conj(DFT(conj(Pfull))) / M; conj(Pfull)[k] = conj(Pk), [M-k] = Pk
        cpx lo; lo.x = Pk.x; lo.y = -Pk.y;
        // in place: d[k] is read only by this thread, d[M-k] (> M/2) by nobody in this kernel
        d[k] = c_mul(lo, a.chirp[k]);
        if (k != 0 && 2 * k != a.M) d[a.M - k] = c_mul(Pk, a.chirp[a.M - k]);
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

// y[n] = Re(conj(c[n]*Y[n])) / M
__device__ void aess_store_body(const SpecArgs &a)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    const float inv = 1.0f / (float)a.M;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < (long long)a.nb * a.M; e += stride) {
        const long long b = e / a.M, n = e % a.M;
        const cpx y = c_mul(a.chirp[n], a.buf[b * a.P + n]);
        a.out[e] = y.x * inv;
    }
}
