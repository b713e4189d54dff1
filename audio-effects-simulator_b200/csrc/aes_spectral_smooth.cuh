// aes_spectral_smooth.cuh -- SpectralFilter transforms for frame lengths M = n1 * n2 whose factors are
// 2, 3 and 5 only (BASELINE's whole-file frame: 960 000 = 960 x 1000), without Bluestein.
//
// The chirp-z path (aes_spectral.cuh) serves ANY M with four power-of-two FFTs of P >= 2M-1 points per
// transform pair: 21 passes over a 16.8 MB buffer for M = 960 000.  When M is smooth the DFT is done
// directly as a four-step (n1 x n2) transform in three kernels, 46 MB of traffic per pair of clips:
//
//   K1  columns:  for a tile of C adjacent columns b, load x[a*n2 + b] (a = 0..n1-1; straight from the
//                 clips / analysis frames: window, pairing z = f0 + i*f1, zero half), n1-point DIF FFTs
//                 in shared memory, times W_M^(b*k1), store Y[k1][b]
//   K2  row pairs: rows k1 and n1-k1 together (they hold bins k and M-k): n2-point DIF FFTs, the
//                 conjugate-symmetry split into the two clips' spectra, the spectral gate
//                 (spectral.py:66-73), recombination, n2-point inverse DIT FFTs, times W_M^(-b*k1)
//                 -- the spectrum never leaves shared memory
//   K3  columns:  n1-point inverse DIT FFTs, 1/M, real part -> clip / frame 2p, imaginary -> 2p+1
//
// Shared-memory FFTs are in-place mixed radix (4, 2, 3, 5): decimation in frequency forward, which
// leaves X[k] at the digit-reversed position, and decimation in time inverse, which takes that
// order; the gate and the global stores translate indices, no permutation pass exists.  Twiddles
// come from tables built in double precision at plan creation (w_n^j per FFT length; W_M^t as a
// product of a 1024-entry and an M/1024-entry table).
#pragma once
#include "aes_spectral.cuh"

#define AESM_MAX_STAGES 12
#define AESM_NT 256                 // threads of the row-pair kernel
#define AESM_NTC 512                // threads of the forward column kernel (61 KB tiles: 3 CTAs per SM, 48 warps)
#define AESM_C 8                    // columns per K1 / K3 tile

struct SmoothDiv { unsigned mul; unsigned d; };     // x / d for x * d < 2^32:  __umulhi(x, mul)

struct SmoothFft {
    int n, ns;
    int r[AESM_MAX_STAGES];         // radices in DIF order
    int len[AESM_MAX_STAGES];       // sub-transform length entering stage i (len[0] = n)
    SmoothDiv dm[AESM_MAX_STAGES];  // division by m_i = len[i] / r[i]
    int toff[AESM_MAX_STAGES];      // stage twiddles T_i[(d-1) * m_i + p] = w_len^(p*d) start here in the plan's table
    int tsize;                      // entries of all stages (stages with m = 1 have none)
};

struct SmoothArgs {
    SmoothFft f1, f2;               // column (n1) and row (n2) transforms
    cpx *buf;                       // [np][n1][n2] work buffer
    const cpx *tw1, *tw2;           // per-stage twiddles of the two transforms (SmoothFft::toff)
    const cpx *twlo, *twhi;         // W_M^t = twhi[t >> 10] * twlo[t & 1023]
    const int *rev1, *rev2;         // position of output index k after the DIF stages (mixed-radix digit reversal)
    const float *frames;            // [nf][M] analysis frames (mode 1) or null
    const float *window;            // [M] Hann window or null
    const float *clips;             // [nf][N][2] stereo clips (mode 2: frame = [zeros(N), mean * window[N:]])
    float *mask;                    // [nf][M/2+1] smoothed mask in/out, or null: starts at ones, not kept
    float *out;                     // [nf][M] (mode 1)
    float *yclips;                  // [nf][N][2] (mode 2): first N samples, both channels
    int M;                          // frame length (n1 * n2 <= 2^23)
    int n1, n2, np, nf;             // np = ceil(nf / 2) transform pairs
    int mode;                       // 1 frames -> out, 2 clips -> yclips
    float thr, red, alpha;
    const cpx *tw10;                // aesm_rows10 twiddles [4][10][100] (960 x 1000 only), see aesm_fill_rows10
    int rows_self_only;             // aesm_rows_body takes only the rows that pair with themselves (0 and n1/2)
};

#ifdef AES_CPU_EMU
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
#endif
__device__ __forceinline__ unsigned aesm_div(unsigned x, const SmoothDiv &d) { return d.d == 1 ? x : __umulhi(x, d.mul); }

__device__ __forceinline__ cpx aesm_cmulc(cpx a, cpx w, bool conj_w)
{
    cpx r;
    if (!conj_w) { r.x = a.x * w.x - a.y * w.y; r.y = a.x * w.y + a.y * w.x; }
    else         { r.x = a.x * w.x + a.y * w.y; r.y = a.y * w.x - a.x * w.y; }
    return r;
}

// small DFTs in registers; INV conjugates the roots of unity
template <int R, bool INV> __device__ __forceinline__ void aesm_dft(cpx (&a)[R])
{
    if (R == 2) {
        const cpx u = a[0], v = a[1 % R];
        a[0] = c_add(u, v); a[1 % R] = c_sub(u, v);
    } else if (R == 4) {
        const cpx s0 = c_add(a[0], a[2 % R]), d0 = c_sub(a[0], a[2 % R]);
        const cpx s1 = c_add(a[1 % R], a[3 % R]), d1 = c_sub(a[1 % R], a[3 % R]);
        cpx jd; jd.x = d1.y; jd.y = -d1.x;                  // -i * d1
        if (INV) { jd.x = -jd.x; jd.y = -jd.y; }            // +i * d1
        a[0] = c_add(s0, s1); a[2 % R] = c_sub(s0, s1);
        a[1 % R] = c_add(d0, jd); a[3 % R] = c_sub(d0, jd);
    } else if (R == 8) {
        const float h = 0.70710678118654752440f;
        cpx b[4], c[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) { b[j] = c_add(a[j], a[(j + 4) % R]); c[j] = c_sub(a[j], a[(j + 4) % R]); }
        cpx t;                                              // c[j] *= w8^j (conjugated for the inverse)
        if (!INV) {
            t.x = h * (c[1].x + c[1].y); t.y = h * (c[1].y - c[1].x); c[1] = t;
            t.x = c[2].y; t.y = -c[2].x; c[2] = t;
            t.x = h * (c[3].y - c[3].x); t.y = -h * (c[3].x + c[3].y); c[3] = t;
        } else {
            t.x = h * (c[1].x - c[1].y); t.y = h * (c[1].x + c[1].y); c[1] = t;
            t.x = -c[2].y; t.y = c[2].x; c[2] = t;
            t.x = -h * (c[3].x + c[3].y); t.y = h * (c[3].x - c[3].y); c[3] = t;
        }
        aesm_dft<4, INV>(b);
        aesm_dft<4, INV>(c);
#pragma unroll
        for (int k = 0; k < 4; ++k) { a[(2 * k) % R] = b[k]; a[(2 * k + 1) % R] = c[k]; }
    } else if (R == 3) {
        const float s = 0.86602540378443864676f;
        const cpx t1 = c_add(a[1 % R], a[2 % R]);
        cpx t2; t2.x = a[0].x - 0.5f * t1.x; t2.y = a[0].y - 0.5f * t1.y;
        const cpx t3 = c_sub(a[1 % R], a[2 % R]);
        cpx jt; jt.x = s * t3.y; jt.y = -s * t3.x;          // -i * s * t3
        if (INV) { jt.x = -jt.x; jt.y = -jt.y; }
        a[0] = c_add(a[0], t1);
        a[1 % R] = c_add(t2, jt); a[2 % R] = c_sub(t2, jt);
    } else {
        static_assert(R == 2 || R == 3 || R == 4 || R == 5 || R == 8, "");
        const float c1 = 0.30901699437494742410f, c2 = -0.80901699437494742410f;
        const float s1 = 0.95105651629515357212f, s2 = 0.58778525229247312917f;
        const cpx t1 = c_add(a[1 % R], a[4 % R]), t2 = c_add(a[2 % R], a[3 % R]);
        const cpx t3 = c_sub(a[1 % R], a[4 % R]), t4 = c_sub(a[2 % R], a[3 % R]);
        cpx m1, m2, q1, q2;
        m1.x = a[0].x + c1 * t1.x + c2 * t2.x; m1.y = a[0].y + c1 * t1.y + c2 * t2.y;
        m2.x = a[0].x + c2 * t1.x + c1 * t2.x; m2.y = a[0].y + c2 * t1.y + c1 * t2.y;
        q1.x = s1 * t3.x + s2 * t4.x; q1.y = s1 * t3.y + s2 * t4.y;
        q2.x = s2 * t3.x - s1 * t4.x; q2.y = s2 * t3.y - s1 * t4.y;
        cpx j1, j2;                                         // -i * q
        j1.x = q1.y; j1.y = -q1.x; j2.x = q2.y; j2.y = -q2.x;
        if (INV) { j1.x = -j1.x; j1.y = -j1.y; j2.x = -j2.x; j2.y = -j2.y; }
        a[0].x += t1.x + t2.x; a[0].y += t1.y + t2.y;
        a[1 % R] = c_add(m1, j1); a[4 % R] = c_sub(m1, j1);
        a[2 % R] = c_add(m2, j2); a[3 % R] = c_sub(m2, j2);
    }
}

// One in-place stage over S sequences held in shared memory: element e of sequence q sits at
// s[q * qs + e * es].  DIF (forward):  out_d[p] = (sum_j a_j w_R^(jd)) * w_len^(p*d) -> block d of the
// sub-transform;  DIT (inverse): the transpose with conjugated roots.
template <int R, bool INV>
__device__ __forceinline__ void aesm_stage(cpx *s, int S, SmoothDiv divS, int qs, int es, const SmoothFft &f, int i, const cpx *tw)
{
    const int len = f.len[i], m = len / R, nbf = f.n / R;
    const cpx *__restrict__ T = tw + f.toff[i];
    for (unsigned e = threadIdx.x; e < (unsigned)(S * nbf); e += blockDim.x) {
        const unsigned fi = aesm_div(e, divS), q = e - fi * S;          // sequences fastest: adjacent columns, adjacent words
        const unsigned blk = aesm_div(fi, f.dm[i]), p = fi - blk * m;
        cpx *base = s + q * qs + (blk * len + p) * es;
        cpx a[R];
#pragma unroll
        for (int j = 0; j < R; ++j) a[j] = base[j * m * es];
        if (INV && m > 1) {
#pragma unroll
            for (int d = 1; d < R; ++d) a[d] = aesm_cmulc(a[d], T[(d - 1) * m + p], true);
        }
        aesm_dft<R, INV>(a);
        if (!INV && m > 1) {
#pragma unroll
            for (int d = 1; d < R; ++d) a[d] = aesm_cmulc(a[d], T[(d - 1) * m + p], false);
        }
#pragma unroll
        for (int j = 0; j < R; ++j) base[j * m * es] = a[j];
    }
}

template <bool INV>
__device__ __forceinline__ void aesm_fft(cpx *s, int S, SmoothDiv divS, int qs, int es, const SmoothFft &f, const cpx *tw)
{
    for (int k = 0; k < f.ns; ++k) {
        const int i = INV ? f.ns - 1 - k : k;
        switch (f.r[i]) {
        case 8: aesm_stage<8, INV>(s, S, divS, qs, es, f, i, tw); break;
        case 4: aesm_stage<4, INV>(s, S, divS, qs, es, f, i, tw); break;
        case 2: aesm_stage<2, INV>(s, S, divS, qs, es, f, i, tw); break;
        case 3: aesm_stage<3, INV>(s, S, divS, qs, es, f, i, tw); break;
        default: aesm_stage<5, INV>(s, S, divS, qs, es, f, i, tw); break;
        }
        __syncthreads();
    }
}

// ---- the same stages with every size a compile-time constant (BASELINE's 960 000 = 960 x 1000): the
// index arithmetic of the run-time version (about half of its instructions) folds into immediates
template <int R, bool INV, int LEN, int NTOT, int S, int QS, int ES, int NT, int TOFF>
__device__ __forceinline__ void aesm_stage_c(cpx *s, const cpx *__restrict__ tw)
{
    constexpr int m = LEN / R, nbf = NTOT / R, total = S * nbf;
    const cpx *__restrict__ T = tw + TOFF;
#pragma unroll
    for (int e0 = 0; e0 < total; e0 += NT) {
        const int e = e0 + (int)threadIdx.x;
        if (e0 + NT > total && e >= total) break;
        const int fi = e / S, q = e % S;
        const int blk = fi / m, p = fi % m;
        cpx *base = s + q * QS + (blk * LEN + p) * ES;
        cpx a[R];
#pragma unroll
        for (int j = 0; j < R; ++j) a[j] = base[j * m * ES];
        if (INV && m > 1) {
#pragma unroll
            for (int d = 1; d < R; ++d) a[d] = aesm_cmulc(a[d], T[(d - 1) * m + p], true);
        }
        aesm_dft<R, INV>(a);
        if (!INV && m > 1) {
#pragma unroll
            for (int d = 1; d < R; ++d) a[d] = aesm_cmulc(a[d], T[(d - 1) * m + p], false);
        }
#pragma unroll
        for (int j = 0; j < R; ++j) base[j * m * ES] = a[j];
    }
}

// four stages R0..R3 of an N-point transform (the order aesm_build_fft picks: 8s, 4s, 2s, 3s, 5s)
template <bool INV, int N, int S, int QS, int ES, int NT, int R0, int R1, int R2, int R3>
__device__ __forceinline__ void aesm_fft_c(cpx *s, const cpx *__restrict__ tw)
{
    constexpr int L0 = N, L1 = L0 / R0, L2 = L1 / R1, L3 = L2 / R2;
    static_assert(L3 == R3, "radices must multiply to N");
    constexpr int T0 = 0;
    constexpr int T1 = T0 + (L0 / R0 > 1 ? (R0 - 1) * (L0 / R0) : 0);
    constexpr int T2 = T1 + (L1 / R1 > 1 ? (R1 - 1) * (L1 / R1) : 0);
    constexpr int T3 = T2 + (L2 / R2 > 1 ? (R2 - 1) * (L2 / R2) : 0);
    if (!INV) {
        aesm_stage_c<R0, INV, L0, N, S, QS, ES, NT, T0>(s, tw); __syncthreads();
        aesm_stage_c<R1, INV, L1, N, S, QS, ES, NT, T1>(s, tw); __syncthreads();
        aesm_stage_c<R2, INV, L2, N, S, QS, ES, NT, T2>(s, tw); __syncthreads();
        aesm_stage_c<R3, INV, L3, N, S, QS, ES, NT, T3>(s, tw); __syncthreads();
    } else {
        aesm_stage_c<R3, INV, L3, N, S, QS, ES, NT, T3>(s, tw); __syncthreads();
        aesm_stage_c<R2, INV, L2, N, S, QS, ES, NT, T2>(s, tw); __syncthreads();
        aesm_stage_c<R1, INV, L1, N, S, QS, ES, NT, T1>(s, tw); __syncthreads();
        aesm_stage_c<R0, INV, L0, N, S, QS, ES, NT, T0>(s, tw); __syncthreads();
    }
}

// SHAPE 1: n1 = 960 = 8*8*3*5, n2 = 1000 = 8*5*5*5
#define AESM_SHAPE_960x1000 1
static inline int aesm_static_shape(const SmoothFft &f1, const SmoothFft &f2)
{
    const int r1[4] = { 8, 8, 3, 5 }, r2[4] = { 8, 5, 5, 5 };
    if (f1.n != 960 || f2.n != 1000 || f1.ns != 4 || f2.ns != 4) return 0;
    for (int i = 0; i < 4; ++i) if (f1.r[i] != r1[i] || f2.r[i] != r2[i]) return 0;
    return AESM_SHAPE_960x1000;
}

// position of output index k after the DIF stages (mixed-radix digit reversal); host side, tabulated per plan
static inline int aesm_rev(int k, const SmoothFft &f)
{
    int pos = 0, len = f.n;
    for (int i = 0; i < f.ns; ++i) {
        const int r = f.r[i], d = k % r;
        k /= r; len /= r;
        pos += d * len;
    }
    return pos;
}

// W_M^t (t in [0, M)) or its conjugate
__device__ __forceinline__ cpx aesm_wM(const SmoothArgs &a, int t, bool conj_w)
{
    const cpx w = aesm_cmulc(a.twhi[t >> 10], a.twlo[t & 1023], false);
    cpx r = w;
    if (conj_w) r.y = -r.y;
    return r;
}

// sample n of the analysis frame of real frame `fr` (or 0 when the pair has no second frame)
__device__ __forceinline__ float aesm_sample(const SmoothArgs &a, int fr, int n, float w)
{
    if (fr >= a.nf) return 0.0f;
    if (a.mode == 2) {
        const int N = a.M / 2;
        if (n < N) return 0.0f;
        const float2 f = reinterpret_cast<const float2 *>(a.clips)[(long long)fr * N + (n - N)];
        return __fmul_rn(__fmul_rn(__fadd_rn(f.x, f.y), 0.5f), w);      // np.mean, then * window, in f32
    }
    return __fmul_rn(a.frames[(long long)fr * a.M + n], w);
}

// ---- K1: column FFTs ---------------------------------------------------------------------------------
template <int SHAPE, int NT = AESM_NTC>
__device__ void aesm_cols_fwd_body(const SmoothArgs &a)
{
    AES_DYN_SMEM(cpx, s);                                   // [n1][C]
    const int n1 = SHAPE == AESM_SHAPE_960x1000 ? 960 : a.n1, n2 = SHAPE == AESM_SHAPE_960x1000 ? 1000 : a.n2, tiles = (n2 + AESM_C - 1) / AESM_C;
    SmoothDiv divC; divC.d = AESM_C; divC.mul = 0x20000000u;        // x / 8
    for (long long w = blockIdx.x; w < (long long)a.np * tiles; w += gridDim.x) {
        const int p = (int)(w / tiles), b0 = (int)(w % tiles) * AESM_C;
        // (measured, r2ah: the zero half as plain stores and the other half's loads batched four trips ahead of use --
        // 3.63 -> 4.19 ms, the batch spills at this kernel's 42 registers; the loop below already overlaps across warps)
        for (int e = threadIdx.x; e < n1 * AESM_C; e += blockDim.x) {
            const int row = e / AESM_C, c = e % AESM_C, b = b0 + c;
            cpx z; z.x = 0.f; z.y = 0.f;
            if (b < n2 && (a.mode != 2 || 2 * row >= n1)) {          // whole-clip frames: the first half is zeros
                const int n = row * n2 + b;
                const float wn = a.window != nullptr ? a.window[n] : 1.0f;
                z.x = aesm_sample(a, 2 * p, n, wn);
                z.y = aesm_sample(a, 2 * p + 1, n, wn);
            }
            s[e] = z;
        }
        __syncthreads();
        if (SHAPE == AESM_SHAPE_960x1000) aesm_fft_c<false, 960, AESM_C, 1, AESM_C, NT, 8, 8, 3, 5>(s, a.tw1);
        else aesm_fft<false>(s, AESM_C, divC, 1, AESM_C, a.f1, a.tw1);
        cpx *dst = a.buf + (long long)p * a.M;
        for (int e = threadIdx.x; e < n1 * AESM_C; e += blockDim.x) {
            const int k1 = e / AESM_C, c = e % AESM_C, b = b0 + c;
            if (b < n2) {
                const cpx v = s[a.rev1[k1] * AESM_C + c];
                dst[k1 * n2 + b] = c_mul(v, aesm_wM(a, b * k1, false));       // b * k1 < n2 * n1 = M
            }
        }
        __syncthreads();
    }
}

// ---- K2: row pairs: forward rows, gate, inverse rows -------------------------------------------------
__device__ __forceinline__ cpx aesm_gate_one(cpx X, float *maskp, bool self_conj, const SmoothArgs &a)
{
    const float mag = sqrtf(X.x * X.x + X.y * X.y);
    const float cur = mag > a.thr ? 1.0f : a.red;               // spectral.py:68
    const float prev = maskp != nullptr ? *maskp : 1.0f;
    const float m = a.alpha * prev + (1.0f - a.alpha) * cur;    // spectral.py:71
    if (maskp != nullptr) *maskp = m;
    cpx Pk; Pk.x = X.x * m; Pk.y = X.y * m;                     // mag * mask * exp(i*phase)
    if (self_conj) Pk.y = 0.f;                                  // irfft ignores the imaginary part of DC / Nyquist
    return Pk;
}

template <int SHAPE>
__device__ void aesm_rows_body(const SmoothArgs &a)
{
    AES_DYN_SMEM(cpx, s);                                   // [2][n2]: row k1, row n1 - k1
    const int n1 = SHAPE == AESM_SHAPE_960x1000 ? 960 : a.n1, n2 = SHAPE == AESM_SHAPE_960x1000 ? 1000 : a.n2, half = n1 / 2 + 1;      // row pairs per transform: k1 = 0 .. n1/2
    const int nbins = a.M / 2 + 1;
    SmoothDiv div2; div2.d = 2; div2.mul = 0x80000000u;
    SmoothDiv div1; div1.d = 1; div1.mul = 0;
    const int per = a.rows_self_only ? 2 : half;            // (the other row pairs went through aesm_rows10_body)
    for (long long w = blockIdx.x; w < (long long)a.np * per; w += gridDim.x) {
        const int p = (int)(w / per), k1 = a.rows_self_only ? (int)(w % per) * (n1 / 2) : (int)(w % per), k1b = (n1 - k1) % n1;
        const bool self = k1 == k1b;                        // rows 0 and n1/2 pair with themselves
        const int S = self ? 1 : 2;
        cpx *rowA = a.buf + (long long)p * a.M + (long long)k1 * n2;
        cpx *rowB = a.buf + (long long)p * a.M + (long long)k1b * n2;
        for (int e = threadIdx.x; e < n2; e += blockDim.x) {
            s[e] = rowA[e];
            if (!self) s[n2 + e] = rowB[e];
        }
        __syncthreads();
        if (SHAPE == AESM_SHAPE_960x1000) {
            if (self) aesm_fft_c<false, 1000, 1, 1000, 1, AESM_NT, 8, 5, 5, 5>(s, a.tw2);
            else aesm_fft_c<false, 1000, 2, 1000, 1, AESM_NT, 8, 5, 5, 5>(s, a.tw2);
        } else aesm_fft<false>(s, S, self ? div1 : div2, n2, 1, a.f2, a.tw2);
        // bins: k = k1 + n1*k2 (row A, column k2)  <->  M - k = k1b + n1*k2b (row B, column k2b)
        const bool two = 2 * p + 1 < a.nf;
        for (int k2 = threadIdx.x; k2 < n2; k2 += blockDim.x) {
            const int k = k1 + n1 * k2;
            const int k2b = k1 == 0 ? (k2 == 0 ? 0 : n2 - k2) : n2 - 1 - k2;
            const int km = k1b + n1 * k2b;                  // = (M - k) mod M
            if (self && km < k) continue;                   // each unordered pair once
            const int ia = a.rev2[k2], ib = (self ? 0 : n2) + a.rev2[k2b];
            const cpx Zk = s[ia], Zm = s[ib];
            const int kk = k <= km ? k : km;                // the rfft bin this pair is
            cpx X0, X1;                                     // rfft bin of frame 2p / 2p+1 at index k
            X0.x = 0.5f * (Zk.x + Zm.x); X0.y = 0.5f * (Zk.y - Zm.y);
            X1.x = 0.5f * (Zk.y + Zm.y); X1.y = 0.5f * (Zm.x - Zk.x);
            if (k > km) { X0.y = -X0.y; X1.y = -X1.y; }    // we hold bin M-kk: its spectrum value is the conjugate
            const bool sc = k == km;
            float *m0 = a.mask != nullptr ? a.mask + (long long)(2 * p) * nbins + kk : nullptr;
            float *m1 = a.mask != nullptr ? a.mask + (long long)(2 * p + 1) * nbins + kk : nullptr;
            const cpx P0 = aesm_gate_one(X0, m0, sc, a);
            cpx P1; P1.x = 0.f; P1.y = 0.f;
            if (two) P1 = aesm_gate_one(X1, m1, sc, a);
            // W = P0full + i*P1full: W[kk] = P0 + i*P1, W[M-kk] = conj(P0) + i*conj(P1)
            cpx Wlo, Whi;
            Wlo.x = P0.x - P1.y; Wlo.y = P0.y + P1.x;
            Whi.x = P0.x + P1.y; Whi.y = P1.x - P0.y;
            s[ia] = k <= km ? Wlo : Whi;
            if (!sc) s[ib] = k <= km ? Whi : Wlo;
        }
        __syncthreads();
        if (SHAPE == AESM_SHAPE_960x1000) {
            if (self) aesm_fft_c<true, 1000, 1, 1000, 1, AESM_NT, 8, 5, 5, 5>(s, a.tw2);
            else aesm_fft_c<true, 1000, 2, 1000, 1, AESM_NT, 8, 5, 5, 5>(s, a.tw2);
        } else aesm_fft<true>(s, S, self ? div1 : div2, n2, 1, a.f2, a.tw2);
        for (int e = threadIdx.x; e < n2; e += blockDim.x) {
            rowA[e] = c_mul(s[e], aesm_wM(a, e * k1, true));
            if (!self) rowB[e] = c_mul(s[n2 + e], aesm_wM(a, e * k1b, true));
        }
        __syncthreads();
    }
}

// ---- K2 for 960 x 1000, register-resident ------------------------------------------------------------
// ncu (profiles/r2ad): aesm_rows_kernel is bound by the shared-memory pipe (84 % of its wavefront rate, a quarter of
// the wavefronts bank conflicts of the digit-reversed gate sweep and of the radix-5 stages): 22 shared-memory
// accesses per element and transform pair (4 stages x load + store, forward and inverse, plus the gate).  Here a
// row of 1000 = 10 x 10 x 10 is transformed in three radix-10 passes held in registers, 100 threads per row with
// 10 elements each, two exchanges through shared memory per transform (8 accesses per element in all):
//
//   n = 100 n1 + 10 n2 + n3,  k = k1 + 10 k2 + 100 k3
//   pass A  thread (n2, n3) = n mod 100:  DFT-10 over n1 -> k1, times w_100^(n2 k1)
//   pass B  thread (k1, n3):              DFT-10 over n2 -> k2, times w_1000^(n3 (k1 + 10 k2))
//   pass C  thread (k1, k2):              DFT-10 over n3 -> k3:  the thread holds columns k1 + 10 k2 + 100 k3
//
// The gate pairs bin k of row k1 (column c) with bin M - k, which is column 999 - c of row n1 - k1 (k1 != 0, n1/2:
// those two rows pair with themselves and stay with aesm_rows_body).  A thread therefore runs row A as thread u and
// row B as thread 99 - u: after pass C it holds columns {kA + 100 k3} of row A and {99 - kA + 100 k3'} of row B, i.e.
// every partner pair (k3' = 9 - k3) sits in its own registers and the gate touches no memory.  The inverse runs the
// three passes backwards with conjugated roots and leaves thread (n2, n3) with x[100 n1 + n mod 100]: global loads
// and stores are 800 contiguous bytes per n1.  Exchange layouts (a thread writes 10 values at one stride and another
// reads 10 at another; both sides hit 16 distinct 8-byte banks per half-warp):
//   exchange 1  (k1, n2, n3) at 106 k1 + 10 n2 + n3      writer lanes: consecutive;  reader lanes: 106 = 10 (mod 16)
//   exchange 2  (k1, k2, n3) at n3 + 10 k1 + 113 k2      writer lanes: consecutive;  reader lanes: 113 = 1  (mod 16)
// Two buffers per row alternate, so an exchange costs one CTA barrier.  Twiddles come from four [10][100] tables
// (thread index fastest: conflict-free), built in double on the host and copied into shared memory once per CTA
// (ncu r2af, tables read through L1: 36 % of the stall samples on the long scoreboard).  The four-step factor of
// the output, W_M^(-b k1) with b = 100 j + u, is W_M^(-u k1) -- one table look-up per thread and row -- times
// W_M^(-100 j k1), ten values per row that ten threads look up and everyone reads as a broadcast.
#define AESR_P1 106
#define AESR_P2 113
#define AESR_ROWBUF 1136            // cpx per exchange buffer: max(10 * 106, 10 * 113), rounded up to 128 bytes
#define AESR_G 2                    // row pairs per CTA (default; the kernel is a template on it)
#define AESR_NT_OF(G) ((100 * (G) + 31) / 32 * 32)
#define AESR_TW_ENTRIES 4000
// exchange buffers | the four twiddle tables | W_M^(-100 j k1) of the two rows of every pair, by item parity
#define AESR_SMEM_OF2(G, PP) (((G) * 2 * (PP) * AESR_ROWBUF + AESR_TW_ENTRIES + (G) * 2 * 2 * 10) * (int)sizeof(cpx))
#define AESR_SMEM_OF(G) AESR_SMEM_OF2(G, 2)
// ... | PF: the next item's two rows, fetched by bulk copies under this item's passes, and one mbarrier per pair
#define AESR_SMEM_OF3(G, PP, PF) (AESR_SMEM_OF2(G, PP) + (PF) * ((G) * 2 * 1000 * (int)sizeof(cpx) + (G) * 8 + 8))
#define AESR_NT AESR_NT_OF(AESR_G)
#define AESR_SMEM_BYTES AESR_SMEM_OF(AESR_G)

// host: t[0][k1][u] = w_100^(n2 k1), u = 10 n2 + n3;          t[1][k2][u] = w_1000^(n3 (k1 + 10 k2)), u = 10 k1 + n3;
//       t[2][n3][u] = conj w_1000^(n3 (k1 + 10 k2)), u = 10 k1 + k2;   t[3][n2][u] = conj w_100^(n2 k1), u = 10 k1 + n3
static inline void aesm_fill_rows10(cpx *t)
{
    const double tp = 2.0 * 3.14159265358979323846;
    for (int j = 0; j < 10; ++j)
        for (int u = 0; u < 100; ++u) {
            const int hi = u / 10, lo = u % 10;
            const double ang[4] = { -tp * (double)((hi * j) % 100) / 100.0,              // n2 = hi, k1 = j
                                    -tp * (double)((lo * (hi + 10 * j)) % 1000) / 1000.0, // k1 = hi, n3 = lo, k2 = j
                                    tp * (double)((j * (hi + 10 * lo)) % 1000) / 1000.0,  // k1 = hi, k2 = lo, n3 = j
                                    tp * (double)((j * hi) % 100) / 100.0 };             // k1 = hi, n2 = j
            for (int q = 0; q < 4; ++q) {
                t[(q * 10 + j) * 100 + u].x = (float)cos(ang[q]);
                t[(q * 10 + j) * 100 + u].y = (float)sin(ang[q]);
            }
        }
}

// DFT-10 as Good-Thomas 2 x 5 (no twiddles inside): n = (5 na + 2 nb) mod 10, k = (5 ka + 6 kb) mod 10
template <bool INV> __device__ __forceinline__ void aesm_dft10(cpx (&a)[10])
{
    cpx e[5], o[5];
#pragma unroll
    for (int nb = 0; nb < 5; ++nb) {
        const cpx u = a[(2 * nb) % 10], v = a[(5 + 2 * nb) % 10];
        e[nb] = c_add(u, v); o[nb] = c_sub(u, v);
    }
    aesm_dft<5, INV>(e);
    aesm_dft<5, INV>(o);
#pragma unroll
    for (int kb = 0; kb < 5; ++kb) { a[(6 * kb) % 10] = e[kb]; a[(5 + 6 * kb) % 10] = o[kb]; }
}

// PP = 2: two exchange buffers per row alternate (one barrier per exchange);  PP = 1: one buffer, a second barrier
// between an exchange's loads and the next exchange's stores, half the exchange memory (more CTAs per SM)
// PF: the rows of the CTA's next item are fetched into shared memory by two 8000-byte bulk copies (issued by one
// thread per pair behind the first barrier, when every thread has taken this item's rows into registers) instead of
// ten 8-byte loads per thread and row at the top of the item -- ncu r2ag: 20 % of the stall samples waited there.
template <int G, int PP = 2, int PF = 0>
__device__ void aesm_rows10_body(const SmoothArgs &a)
{
    AES_DYN_SMEM(cpx, s);                                   // [G][row A, row B][PP][AESR_ROWBUF]
    constexpr int n1 = 960, n2 = 1000, NP = n1 / 2 - 1;     // row pairs k1 = 1 .. 479 per transform
    const int M = n1 * n2;
    const int tid = threadIdx.x, g = tid / 100, u = tid - 100 * g, uB = 99 - u;
    const int hi = u / 10, lo = u - 10 * hi, hiB = 9 - hi, loB = 9 - lo;
    cpx *const A0 = s + (size_t)(g < G ? g : 0) * 2 * PP * AESR_ROWBUF, *const A1 = A0 + (PP - 1) * AESR_ROWBUF;
    cpx *const B0 = A0 + PP * AESR_ROWBUF, *const B1 = B0 + (PP - 1) * AESR_ROWBUF;
    cpx *const tws = s + (size_t)G * 2 * PP * AESR_ROWBUF;
    const cpx *const t1f = tws, *const t2f = tws + 1000, *const t2i = tws + 2000, *const t1i = tws + 3000;
    cpx *const vj = tws + AESR_TW_ENTRIES + (g < G ? g : 0) * 40;       // [parity][row A, row B][10]
    cpx *const stg = tws + AESR_TW_ENTRIES + G * 40 + (g < G ? g : 0) * 2000;      // [row A, row B][1000]
    unsigned long long *const mbar = reinterpret_cast<unsigned long long *>(tws + AESR_TW_ENTRIES + G * 40 + G * 2000) + (g < G ? g : 0);
    for (int i = tid; i < AESR_TW_ENTRIES; i += blockDim.x) tws[i] = a.tw10[i];
    if (PF && g < G && u == 0) { aes_mbar_init(mbar, 1); aes_mbar_init_fence(); }
    __syncthreads();
    const int nbins = M / 2 + 1;
    const long long total = (long long)a.np * NP, groups = (total + G - 1) / G;
    // hand the rows of item `it` to the copy engine (one thread per pair)
    auto fetch = [&](long long it) {
        const int pp = (int)(it / NP), kr = 1 + (int)(it % NP);
        const cpx *ra = a.buf + (long long)pp * M + (long long)kr * n2, *rb = a.buf + (long long)pp * M + (long long)(n1 - kr) * n2;
        aes_fence_proxy_async_smem();                       // the staging rows were read with ordinary loads
        aes_mbar_expect(mbar, 2 * 1000 * (unsigned)sizeof(cpx));
        aes_bulk_g2s(stg, ra, 1000 * (unsigned)sizeof(cpx), mbar, aes_policy_evict_first());
        aes_bulk_g2s(stg + 1000, rb, 1000 * (unsigned)sizeof(cpx), mbar, aes_policy_evict_first());
#ifdef AES_CPU_EMU
        aes_mbar_complete_emu(mbar);
#endif
    };
    if (PF && g < G && u == 0 && (long long)blockIdx.x * G + g < total) fetch((long long)blockIdx.x * G + g);
    unsigned uses = 0;                                      // items this pair has taken from its staging rows
    int par = 0;
    for (long long w = blockIdx.x; w < groups; w += gridDim.x, par ^= 1) {
        const long long item = w * G + g;
        const bool live = g < G && item < total;
        const int p = live ? (int)(item / NP) : 0, k1r = live ? 1 + (int)(item % NP) : 1, k1b = n1 - k1r;
        cpx *const rowA = a.buf + (long long)p * M + (long long)k1r * n2;
        cpx *const rowB = a.buf + (long long)p * M + (long long)k1b * n2;
        cpx A[10], B[10];
        cpx baseA, baseB;
        baseA.x = baseA.y = baseB.x = baseB.y = 0.f;
        if (live) {
            if (PF) {
                aes_mbar_wait(mbar, uses);
                ++uses;
#pragma unroll
                for (int j = 0; j < 10; ++j) { A[j] = stg[100 * j + u]; B[j] = stg[1000 + 100 * j + uB]; }
            } else {
#pragma unroll
                for (int j = 0; j < 10; ++j) { A[j] = rowA[100 * j + u]; B[j] = rowB[100 * j + uB]; }
            }
            // (read behind the fourth barrier; the slot of the other parity may still be in use by a thread that has
            // not left the previous item)
            if (u < 10) {
                vj[par * 20 + u] = aesm_wM(a, 100 * u * k1r, true);
                vj[par * 20 + 10 + u] = aesm_wM(a, 100 * u * k1b, true);
            }
            baseA = aesm_wM(a, u * k1r, true);
            baseB = aesm_wM(a, uB * k1b, true);
            aesm_dft10<false>(A);
            aesm_dft10<false>(B);
#pragma unroll
            for (int k = 1; k < 10; ++k) {
                A[k] = c_mul(A[k], t1f[k * 100 + u]);
                B[k] = c_mul(B[k], t1f[k * 100 + uB]);
            }
#pragma unroll
            for (int k = 0; k < 10; ++k) { A0[AESR_P1 * k + u] = A[k]; B0[AESR_P1 * k + uB] = B[k]; }
        }
        __syncthreads();
        if (PF && g < G && u == 0 && (w + gridDim.x) * G + g < total) fetch((w + gridDim.x) * G + g);
        if (live) {
#pragma unroll
            for (int j = 0; j < 10; ++j) { A[j] = A0[AESR_P1 * hi + 10 * j + lo]; B[j] = B0[AESR_P1 * hiB + 10 * j + loB]; }
        }
        if (PP == 1) __syncthreads();                       // every thread holds its values: the buffer is free
        if (live) {
            aesm_dft10<false>(A);
            aesm_dft10<false>(B);
#pragma unroll
            for (int k = 0; k < 10; ++k) {
                A[k] = c_mul(A[k], t2f[k * 100 + u]);
                B[k] = c_mul(B[k], t2f[k * 100 + uB]);
            }
#pragma unroll
            for (int k = 0; k < 10; ++k) { A1[u + AESR_P2 * k] = A[k]; B1[uB + AESR_P2 * k] = B[k]; }
        }
        __syncthreads();
        if (live) {
#pragma unroll
            for (int j = 0; j < 10; ++j) { A[j] = A1[j + 10 * hi + AESR_P2 * lo]; B[j] = B1[j + 10 * hiB + AESR_P2 * loB]; }
        }
        if (PP == 1) __syncthreads();                       // every thread holds its values: the buffer is free
        if (live) {
            aesm_dft10<false>(A);
            aesm_dft10<false>(B);
            // A[k3] = column kA + 100 k3 of row k1r, B[k3] = column 99 - kA + 100 k3 of row n1 - k1r
            const int kA = hi + 10 * lo;
            const bool two = 2 * p + 1 < a.nf;
#pragma unroll
            for (int k3 = 0; k3 < 10; ++k3) {
                const int k = k1r + n1 * (kA + 100 * k3), km = M - k;
                const cpx Zk = A[k3], Zm = B[9 - k3];
                const int kk = k <= km ? k : km;            // the rfft bin this pair is
                cpx X0, X1;                                 // rfft bin of frame 2p / 2p+1 at index k (see aesm_rows_body)
                X0.x = 0.5f * (Zk.x + Zm.x); X0.y = 0.5f * (Zk.y - Zm.y);
                X1.x = 0.5f * (Zk.y + Zm.y); X1.y = 0.5f * (Zm.x - Zk.x);
                if (k > km) { X0.y = -X0.y; X1.y = -X1.y; }
                float *m0 = a.mask != nullptr ? a.mask + (long long)(2 * p) * nbins + kk : nullptr;
                float *m1 = a.mask != nullptr ? a.mask + (long long)(2 * p + 1) * nbins + kk : nullptr;
                const cpx P0 = aesm_gate_one(X0, m0, false, a);
                cpx P1; P1.x = 0.f; P1.y = 0.f;
                if (two) P1 = aesm_gate_one(X1, m1, false, a);
                cpx Wlo, Whi;
                Wlo.x = P0.x - P1.y; Wlo.y = P0.y + P1.x;
                Whi.x = P0.x + P1.y; Whi.y = P1.x - P0.y;
                A[k3] = k <= km ? Wlo : Whi;
                B[9 - k3] = k <= km ? Whi : Wlo;
            }
            aesm_dft10<true>(A);
            aesm_dft10<true>(B);
#pragma unroll
            for (int j = 0; j < 10; ++j) {
                A[j] = c_mul(A[j], t2i[j * 100 + u]);
                B[j] = c_mul(B[j], t2i[j * 100 + uB]);
            }
#pragma unroll
            for (int j = 0; j < 10; ++j) { A0[j + 10 * hi + AESR_P2 * lo] = A[j]; B0[j + 10 * hiB + AESR_P2 * loB] = B[j]; }
        }
        __syncthreads();
        if (live) {
#pragma unroll
            for (int j = 0; j < 10; ++j) { A[j] = A0[u + AESR_P2 * j]; B[j] = B0[uB + AESR_P2 * j]; }
        }
        if (PP == 1) __syncthreads();                       // every thread holds its values: the buffer is free
        if (live) {
            aesm_dft10<true>(A);
            aesm_dft10<true>(B);
#pragma unroll
            for (int j = 1; j < 10; ++j) {
                A[j] = c_mul(A[j], t1i[j * 100 + u]);
                B[j] = c_mul(B[j], t1i[j * 100 + uB]);
            }
#pragma unroll
            for (int j = 0; j < 10; ++j) { A1[AESR_P1 * hi + 10 * j + lo] = A[j]; B1[AESR_P1 * hiB + 10 * j + loB] = B[j]; }
        }
        __syncthreads();
        if (live) {
#pragma unroll
            for (int j = 0; j < 10; ++j) { A[j] = A1[AESR_P1 * j + u]; B[j] = B1[AESR_P1 * j + uB]; }
        }
        if (PP == 1) __syncthreads();                       // every thread holds its values: the buffer is free
        if (live) {
            aesm_dft10<true>(A);
            aesm_dft10<true>(B);
#pragma unroll
            for (int j = 0; j < 10; ++j) {
                rowA[100 * j + u] = c_mul(A[j], c_mul(vj[par * 20 + j], baseA));
                rowB[100 * j + uB] = c_mul(B[j], c_mul(vj[par * 20 + 10 + j], baseB));
            }
        }
        // (the next item's first stores go to the buffers last read ahead of the fourth barrier)
    }
}

// ---- K3: inverse column FFTs, outputs --------------------------------------------------------------------
template <int SHAPE>
__device__ void aesm_cols_inv_body(const SmoothArgs &a)
{
    AES_DYN_SMEM(cpx, s);
    const int n1 = SHAPE == AESM_SHAPE_960x1000 ? 960 : a.n1, n2 = SHAPE == AESM_SHAPE_960x1000 ? 1000 : a.n2, tiles = (n2 + AESM_C - 1) / AESM_C;
    const float inv = 1.0f / (float)a.M;
    const int N = a.M / 2;
    SmoothDiv divC; divC.d = AESM_C; divC.mul = 0x20000000u;
    for (long long w = blockIdx.x; w < (long long)a.np * tiles; w += gridDim.x) {
        const int p = (int)(w / tiles), b0 = (int)(w % tiles) * AESM_C;
        const cpx *src = a.buf + (long long)p * a.M;
        // (measured, r2ah: six trips' loads batched ahead of their stores -- 128 registers, 2 CTAs per SM, 2.64 -> 3.56 ms)
        for (int e = threadIdx.x; e < n1 * AESM_C; e += blockDim.x) {
            const int k1 = e / AESM_C, c = e % AESM_C, b = b0 + c;
            cpx z; z.x = 0.f; z.y = 0.f;
            if (b < n2) z = src[k1 * n2 + b];
            s[a.rev1[k1] * AESM_C + c] = z;
        }
        __syncthreads();
        if (SHAPE == AESM_SHAPE_960x1000) aesm_fft_c<true, 960, AESM_C, 1, AESM_C, AESM_NT, 8, 8, 3, 5>(s, a.tw1);
        else aesm_fft<true>(s, AESM_C, divC, 1, AESM_C, a.f1, a.tw1);
        const bool two = 2 * p + 1 < a.nf;
        for (int e = threadIdx.x; e < n1 * AESM_C; e += blockDim.x) {
            const int row = e / AESM_C, c = e % AESM_C, b = b0 + c;
            if (b >= n2) continue;
            const int n = row * n2 + b;
            const cpx v = s[e];
            if (a.mode == 2) {
                if (n < N) {                                // the block emits the first N samples, both channels
                    const float y0 = v.x * inv, y1 = v.y * inv;
                    reinterpret_cast<float2 *>(a.yclips)[(long long)(2 * p) * N + n] = make_float2(y0, y0);
                    if (two) reinterpret_cast<float2 *>(a.yclips)[(long long)(2 * p + 1) * N + n] = make_float2(y1, y1);
                }
            } else {
                a.out[(long long)(2 * p) * a.M + n] = v.x * inv;
                if (two) a.out[(long long)(2 * p + 1) * a.M + n] = v.y * inv;
            }
        }
        __syncthreads();
    }
}

// ---- host: factorisation -----------------------------------------------------------------------------------
static inline bool aesm_build_fft(int n, SmoothFft *f)
{
    f->n = n; f->ns = 0; f->tsize = 0;
    int rem = n, len = n;
    const int order[5] = { 8, 4, 2, 3, 5 };     // odd radices last: their unit-stride last stage is bank-conflict free
    for (int oi = 0; oi < 5; ++oi) {
        const int r = order[oi];
        while (rem % r == 0 && rem > 1) {
            if (f->ns >= AESM_MAX_STAGES) return false;
            f->r[f->ns] = r; f->len[f->ns] = len;
            const unsigned m = (unsigned)(len / r);
            f->dm[f->ns].d = m;
            f->dm[f->ns].mul = m == 1 ? 0u : (unsigned)(0x100000000ULL / m) + 1u;
            f->toff[f->ns] = f->tsize;
            if (m > 1) f->tsize += (r - 1) * (int)m;
            ++f->ns; rem /= r; len /= r;
        }
    }
    return rem == 1;
}

// stage twiddles of one transform, in double (host; a few thousand entries)
static inline void aesm_fill_twiddles(const SmoothFft &f, cpx *t)
{
    for (int i = 0; i < f.ns; ++i) {
        const int r = f.r[i], len = f.len[i], m = len / r;
        if (m == 1) continue;
        for (int d = 1; d < r; ++d)
            for (int p = 0; p < m; ++p) {
                const double ang = -2.0 * 3.14159265358979323846 * (double)(((long long)p * d) % len) / (double)len;
                t[f.toff[i] + (d - 1) * m + p].x = (float)cos(ang);
                t[f.toff[i] + (d - 1) * m + p].y = (float)sin(ang);
            }
    }
}

// n1 x n2 = M with both factors {2,3,5}-smooth, n1 even, and small enough for shared memory;
// the most balanced such split, or false
static inline bool aesm_split(long long M, int *n1, int *n2)
{
    const int max1 = 2048, max2 = 4096;
    long long best = -1;
    for (long long a = 2; a <= max1 && a <= M; a += 2) {
        if (M % a) continue;
        const long long b = M / a;
        if (b > max2 || b < 2) continue;
        SmoothFft t1, t2;
        if (!aesm_build_fft((int)a, &t1) || !aesm_build_fft((int)b, &t2)) continue;
        if ((unsigned long long)a * AESM_C * (unsigned long long)b >= 0x80000000ULL) continue;
        const long long score = a > b ? a - b : b - a;
        if (best < 0 || score < best) { best = score; *n1 = (int)a; *n2 = (int)b; }
    }
    return best >= 0;
}
