// aes_analysis.cu -- C-ABI entries of the spectrum / chromagram analysis (aes_analysis.cuh).
#include <new>

#include "aes_common.h"
#include "aes_chain_kernel.cuh"      // AES_DYN_SMEM
#include "aes_analysis.cuh"

template <int R> __global__ void __launch_bounds__(AescGeo<R>::NT, 1) aesa_kernel(const __grid_constant__ AnalysisArgs q) { aesa_body<R>(q); }

template <int R> static int analysis_launch(const AnalysisArgs &q, int64_t n_pairs, cudaStream_t st)
{
    using G = AescGeo<R>;
    const int smem = G::SMEM_CPX * (int)sizeof(cpx);
    AES_CUDA(cudaFuncSetAttribute(aesa_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    aesa_kernel<R><<<(unsigned)n_pairs, G::NT, smem, st>>>(q);
    aes_count_launch();
    AES_CUDA(cudaGetLastError());
    return 0;
}

AES_EXPORT int aes_spectrum_chroma(const float *a, const float *b, int64_t n_pairs, int64_t n_samples, int n_fft,
                                   double sample_rate, float *mag_db, float *mag_lin, float *chroma, float *peak_freq,
                                   void *stream)
{
    AES_REQUIRE(n_fft == 16384 || n_fft == 2048 || n_fft == 256, "n_fft 16384 (the page's FFT_SIZE), 2048 or 256");
    AES_REQUIRE(n_samples >= n_fft, "a signal must hold at least n_fft samples (02_custom.js:179 slices the last FFT_SIZE)");
    AES_REQUIRE(sample_rate > 0.0, "sample rate");
    if (n_pairs <= 0) return 0;
    AES_REQUIRE(n_pairs < (1LL << 31), "too many signal pairs");
    AES_REQUIRE(a && b && mag_db && mag_lin && chroma && peak_freq, "NULL device buffer");
    AnalysisArgs q;
    q.a = a; q.b = b; q.db = mag_db; q.lin = mag_lin; q.chroma = chroma; q.peak_freq = peak_freq;
    q.n_samples = n_samples; q.sample_rate = sample_rate;
    cudaStream_t st = (cudaStream_t)stream;
    return n_fft == 16384 ? analysis_launch<32>(q, n_pairs, st) : n_fft == 2048 ? analysis_launch<16>(q, n_pairs, st)
                                                                                 : analysis_launch<8>(q, n_pairs, st);
}

AES_EXPORT int aes_spectrum_chroma_host(const float *a_host, const float *b_host, int64_t n_pairs, int64_t n_samples,
                                        int n_fft, double sample_rate, float *mag_db, float *mag_lin, float *chroma,
                                        float *peak_freq)
{
    AES_REQUIRE(n_fft == 16384 || n_fft == 2048 || n_fft == 256, "n_fft 16384 (the page's FFT_SIZE), 2048 or 256");
    AES_REQUIRE(n_samples >= n_fft, "a signal must hold at least n_fft samples (02_custom.js:179 slices the last FFT_SIZE)");
    if (n_pairs <= 0) return 0;
    AES_REQUIRE(a_host && b_host && mag_db && chroma && peak_freq, "NULL host buffer");
    // only the analysed tail of every signal crosses the bus
    const size_t nb = (size_t)n_fft / 2 + 1, sig = (size_t)n_pairs * n_fft * sizeof(float), mags = (size_t)n_pairs * 2 * nb * sizeof(float);
    char *d = nullptr;
    AES_CUDA(cudaMalloc(&d, 2 * sig + 2 * mags + (size_t)n_pairs * 2 * 13 * sizeof(float)));
    float *da = (float *)d, *db_ = (float *)(d + sig), *ddb = (float *)(d + 2 * sig), *dlin = (float *)(d + 2 * sig + mags);
    float *dch = (float *)(d + 2 * sig + 2 * mags), *dpk = dch + (size_t)n_pairs * 24;
    int rc = [&]() -> int {
        const size_t row = (size_t)n_fft * sizeof(float), pitch = (size_t)n_samples * sizeof(float);
        AES_CUDA(cudaMemcpy2D(da, row, a_host + (n_samples - n_fft), pitch, row, (size_t)n_pairs, cudaMemcpyHostToDevice));
        AES_CUDA(cudaMemcpy2D(db_, row, b_host + (n_samples - n_fft), pitch, row, (size_t)n_pairs, cudaMemcpyHostToDevice));
        int r = aes_spectrum_chroma(da, db_, n_pairs, n_fft, n_fft, sample_rate, ddb, dlin, dch, dpk, nullptr);
        if (r) return r;
        AES_CUDA(cudaMemcpy(mag_db, ddb, mags, cudaMemcpyDeviceToHost));
        if (mag_lin) AES_CUDA(cudaMemcpy(mag_lin, dlin, mags, cudaMemcpyDeviceToHost));
        AES_CUDA(cudaMemcpy(chroma, dch, (size_t)n_pairs * 24 * sizeof(float), cudaMemcpyDeviceToHost));
        AES_CUDA(cudaMemcpy(peak_freq, dpk, (size_t)n_pairs * 2 * sizeof(float), cudaMemcpyDeviceToHost));
        return 0;
    }();
    cudaFree(d);
    return rc;
}
