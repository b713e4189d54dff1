// aes_analysis.cuh -- the plot-side analysis of the reference, on the device:
// assets/02_custom.js:108-154 `calculateSpectrumAndChroma` (4-term Blackman-Harris window, n_fft-point
// FFT, magnitudes in dB, peak frequency above 60 Hz) and :65-106 `calculateChroma` (12 pitch classes from
// the bins between 70 Hz and 5 kHz above 15 % of the peak, weighted, normalised, cubed).  The page runs it
// on the last FFT_SIZE = 16384 samples of the original AND the processed signal (02_custom.js:179-184):
// two real signals, i.e. ONE complex transform with the L/R separation of the convolution reverb's forward
// pass (aes_convreverb.cuh) -- `a` rides in the real part, `b` in the imaginary part.
// One CTA per (original, processed) pair; the JS works in float64, here the transform is float32 and the
// note decisions (frequency -> MIDI note) are float64.  The JS cannot run in this image (no node):
// parity is checked against a float64 numpy restatement in the test tree, unpinned by the reference.
#pragma once
#include "aes_convreverb.cuh"

#ifdef AES_CPU_EMU
static inline double cospi(double x) { return cos(M_PI * x); }
#endif

struct AnalysisArgs {
    const float *a, *b;         // [n_pairs][n_samples] mono signals (b may equal a)
    float *db;                  // [n_pairs][2][n_fft/2+1]  20*log10(|X|/n_fft + 1e-9)   (02_custom.js:142-143)
    float *lin;                 // [n_pairs][2][n_fft/2+1]  |X|                            (magnitudesLin, :140)
    float *chroma;              // [n_pairs][2][12]
    float *peak_freq;           // [n_pairs][2]
    long long n_samples;
    double sample_rate;
};

// chroma / peak of one signal from its stored magnitudes; every thread of the CTA calls it
template <int NT>
__device__ __forceinline__ void aesa_reduce(const float *lin, const float *db, int n_fft, double fs, float *chroma_out,
                                            float *peak_out, int tid, double *sh /* 64 doubles */)
{
    const int n_bins = n_fft / 2 + 1;
    const double bin_hz = fs / (double)n_fft;
    // pass 1: maximum magnitude over every bin (02_custom.js:69-71); peak in dB above 60 Hz, first maximum (:145-148)
    float mx = 0.f, pk = -INFINITY;
    int pk_k = 0x7fffffff;
    for (int k = tid; k < n_bins; k += NT) {
        mx = fmaxf(mx, lin[k]);
        if ((double)k * bin_hz > 60.0) {
            const float d = db[k];
            if (d > pk) { pk = d; pk_k = k; }
        }
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        const float opk = __shfl_xor_sync(0xffffffffu, pk, o);
        const int ok = __shfl_xor_sync(0xffffffffu, pk_k, o);
        if (opk > pk || (opk == pk && ok < pk_k)) { pk = opk; pk_k = ok; }
    }
    float *shf = reinterpret_cast<float *>(sh);
    int *shi = reinterpret_cast<int *>(sh + 32);        // floats [0, 64) | ints from double 32 | chroma totals from double 48
    const int warp = tid >> 5, lane = tid & 31, nw = NT / 32;
    __syncthreads();
    if (lane == 0) { shf[warp] = mx; shf[32 + warp] = pk; shi[warp] = pk_k; }
    __syncthreads();
    mx = 0.f; pk = -INFINITY; pk_k = 0x7fffffff;
    for (int w = 0; w < nw; ++w) {
        mx = fmaxf(mx, shf[w]);
        const float opk = shf[32 + w];
        const int ok = shi[w];
        if (opk > pk || (opk == pk && ok < pk_k)) { pk = opk; pk_k = ok; }
    }
    __syncthreads();
    // pass 2: pitch classes (02_custom.js:74-97); the sums in double, in a fixed order (lane, warp)
    const float threshold = mx * 0.15f;
    double acc[12];
#pragma unroll
    for (int c = 0; c < 12; ++c) acc[c] = 0.0;
    for (int k = 1 + tid; k < n_bins; k += NT) {
        const double freq = (double)k * bin_hz;
        if (freq < 70.0 || freq > 5000.0) continue;
        float weighting = 1.0f;
        if (freq > 800.0) weighting *= 0.5f;
        if (freq > 1500.0) weighting *= 0.1f;
        const float mw = lin[k] * weighting;
        if (mw < threshold) continue;
        const double midi = 12.0 * log2(freq / 440.0) + 69.0;
        const double nearest = floor(midi + 0.5);                       // Math.round
        if (fabs(midi - nearest) > 0.50) continue;
        int pc = (int)nearest % 12;
        pc = (pc + 12) % 12;
#pragma unroll
        for (int c = 0; c < 12; ++c) acc[c] += c == pc ? (double)mw : 0.0;
    }
#pragma unroll
    for (int c = 0; c < 12; ++c) {
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], o);
    }
    double *shc = sh + 48;              // [12] totals
    if (tid < 12) shc[tid] = 0.0;
    __syncthreads();
    for (int w = 0; w < nw; ++w) {      // warps add their sums one after the other: a fixed order
        if (warp == w && lane == 0) {
#pragma unroll
            for (int c = 0; c < 12; ++c) shc[c] += acc[c];
        }
        __syncthreads();
    }
    if (tid == 0) {
        double mc = shc[0];
        for (int c = 1; c < 12; ++c) mc = fmax(mc, shc[c]);
        mc += 1e-9;
        for (int c = 0; c < 12; ++c) {
            const double v = shc[c] / mc;
            chroma_out[c] = (float)(v * v * v);
        }
        *peak_out = pk_k == 0x7fffffff ? 0.f : (float)((double)pk_k * bin_hz);
    }
    __syncthreads();
}

template <int R>
__device__ __forceinline__ void aesa_body(const AnalysisArgs &q)
{
    using G = AescGeo<R>;
    constexpr int N = G::N, NB = N / 2 + 1;
    AES_DYN_SMEM(cpx, s);
    const int tid = threadIdx.x;
    const long long pair = blockIdx.x;
    // the last n_fft samples (02_custom.js:179-181)
    const float *a = q.a + pair * q.n_samples + (q.n_samples - N);
    const float *b = q.b + pair * q.n_samples + (q.n_samples - N);
    float *lin = q.lin + pair * 2 * NB, *db = q.db + pair * 2 * NB;
    aesc_fwd<R>(s, nullptr, tid, 0.5f,
        [&](int n) {
            // 4-term Blackman-Harris (02_custom.js:113-117), in double like the page
            const double ph = (double)n / (double)(N - 1);
            const double w = 0.35875 - 0.48829 * cospi(2.0 * ph) + 0.14128 * cospi(4.0 * ph) - 0.01168 * cospi(6.0 * ph);
            cpx v;
            v.x = (float)((double)a[n] * w);
            v.y = (float)((double)b[n] * w);
            return v;
        },
        [&](int g, cpx v) {
            // separated spectra: g < N/2 -> A[g]; N/2 < g < N -> B[N-g]; g = N/2 -> A[N/2]; N, N+1 -> B[0], B[N/2] (real)
            const int sig = (g > N / 2 && g < N) || g >= N;
            const int k = g < N ? (g <= N / 2 ? g : N - g) : (g == N ? 0 : N / 2);
            const float m = sqrtf(v.x * v.x + v.y * v.y);
            lin[sig * NB + k] = m;
            db[sig * NB + k] = 20.0f * log10f(m / (float)N + 1e-9f);
        });
    __syncthreads();                    // this CTA's global stores are visible to its own threads
    double *sh = reinterpret_cast<double *>(s);
    for (int sig = 0; sig < 2; ++sig)
        aesa_reduce<G::NT>(lin + sig * NB, db + sig * NB, N, q.sample_rate, q.chroma + (pair * 2 + sig) * 12,
                           q.peak_freq + pair * 2 + sig, tid, sh);
}
