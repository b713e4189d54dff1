// aes_convreverb.cu -- host side and kernels of the IR-convolution reverb (see
// aes_convreverb.cuh for the algorithm).
#include <algorithm>
#include <math.h>
#include <new>
#include <string.h>
#include <vector>

#include "aes_common.h"
#include "aes_chain_kernel.cuh"      // AES_DYN_SMEM
#include "aes_convreverb.cuh"

template <int R> __global__ void __launch_bounds__(AescGeo<R>::NT, 1) aesc_fwd_kernel(const __grid_constant__ ConvArgs a) { aesc_fwd_body<R>(a); }
template <int R> __global__ void __launch_bounds__(AescGeo<R>::NT, 1) aesc_inv_kernel(const __grid_constant__ ConvArgs a) { aesc_inv_body<R>(a); }
template <int R> __global__ void __launch_bounds__(AescGeo<R>::NT, 1) aesc_ir_prep_kernel(const float *ir, int n_taps, cpx *H, const cpx *tw) { aesc_ir_prep_body<R>(ir, n_taps, H, tw); }
template <int PC> __global__ void __launch_bounds__(AESC_MT, 2) aesc_mac_kernel(const __grid_constant__ ConvArgs a, int NS, int p0, int cpi, int acc) { aesc_mac_body<PC>(a, NS, p0, cpi, acc); }

struct aes_convreverb_plan {
    int L = 14, R = 32, N = 0, NS = 0, BK = 0, P = 0;   // NS: row length of the spectra; P: partitions of the impulse response
    int PC = 18, Ppad = 0;                          // partitions per MAC pass; P rounded up to a multiple of it (zero spectra)
    int sms = 148;
    long long n_taps = 0;
    cpx *d_tw = nullptr, *d_H = nullptr;
    void *d_work = nullptr;         // Z | W
    size_t work_cap = 0;
    void *d_in = nullptr, *d_out = nullptr;     // staging of the host entry
    size_t io_cap = 0;
};

#define AESC_CLIPS_PER_ITEM 4
#define AESC_WORK_LIMIT ((size_t)24 << 30)       // bytes of spectra kept at once; larger batches are chunked

static size_t mac_smem(int pc) { return (size_t)2 * pc * AESC_MT * sizeof(cpx) + 16; }

template <int R> static int conv_setup(aes_convreverb_plan *pl, const float *d_ir)
{
    using G = AescGeo<R>;
    const int fft_smem = G::SMEM_CPX * (int)sizeof(cpx);
    AES_CUDA(cudaFuncSetAttribute(aesc_fwd_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, fft_smem));
    AES_CUDA(cudaFuncSetAttribute(aesc_inv_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, fft_smem));
    AES_CUDA(cudaFuncSetAttribute(aesc_ir_prep_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, fft_smem));
    AES_CUDA(cudaFuncSetAttribute(aesc_mac_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mac_smem(4)));
    AES_CUDA(cudaFuncSetAttribute(aesc_mac_kernel<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mac_smem(9)));
    AES_CUDA(cudaFuncSetAttribute(aesc_mac_kernel<18>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mac_smem(18)));
    aesc_ir_prep_kernel<R><<<pl->P, G::NT, fft_smem>>>(d_ir, (int)pl->n_taps, pl->d_H, pl->d_tw);
    aes_count_launch();
    AES_CUDA(cudaGetLastError());
    AES_CUDA(cudaDeviceSynchronize());
    return 0;
}

AES_EXPORT int aes_convreverb_plan_create(const float *ir_host, int64_t n_taps, int block_log2,
                                          aes_convreverb_plan **out)
{
    AES_REQUIRE(out != nullptr && ir_host != nullptr, "NULL argument");
    AES_REQUIRE(n_taps >= 1 && n_taps <= (1 << 24), "impulse response of 1..16M taps");
    if (block_log2 == 0) block_log2 = 14;
    AES_REQUIRE(block_log2 == 14 || block_log2 == 11 || block_log2 == 8, "FFT size 2^8, 2^11 or 2^14");
    aes_convreverb_plan *pl = new (std::nothrow) aes_convreverb_plan();
    if (!pl) { aes_set_error("out of host memory"); return AES_ERR_NOMEM; }
    pl->L = block_log2; pl->N = 1 << pl->L; pl->NS = pl->N + 32; pl->BK = pl->N / 2;
    pl->n_taps = n_taps;
    pl->R = pl->L == 14 ? 32 : pl->L == 11 ? 16 : 8;
    pl->P = (int)((n_taps + pl->BK - 1) / pl->BK);
    pl->PC = pl->P <= 4 ? 4 : pl->P <= 9 ? 9 : 18;
    pl->Ppad = (pl->P + pl->PC - 1) / pl->PC * pl->PC;
    int rc = [&]() -> int {
        int dev = 0;
        AES_CUDA(cudaGetDevice(&dev));
        AES_CUDA(cudaDeviceGetAttribute(&pl->sms, cudaDevAttrMultiProcessorCount, dev));
        std::vector<cpx> tw((size_t)pl->N / 2);
        for (int q = 0; q < pl->N / 2; ++q) {
            const double ang = -2.0 * M_PI * (double)q / (double)pl->N;
            tw[q].x = (float)cos(ang); tw[q].y = (float)sin(ang);
        }
        float *d_ir = nullptr;
        AES_CUDA(cudaMalloc(&pl->d_tw, tw.size() * sizeof(cpx)));
        AES_CUDA(cudaMemcpy(pl->d_tw, tw.data(), tw.size() * sizeof(cpx), cudaMemcpyHostToDevice));
        const size_t hbytes = ((size_t)pl->Ppad * pl->NS + AESC_MT) * sizeof(cpx);     // a tile of slack: the last bin tile overruns its row
        AES_CUDA(cudaMalloc(&pl->d_H, hbytes));
        AES_CUDA(cudaMemset(pl->d_H, 0, hbytes));
        AES_CUDA(cudaMalloc(&d_ir, (size_t)n_taps * 2 * sizeof(float)));
        AES_CUDA(cudaMemcpy(d_ir, ir_host, (size_t)n_taps * 2 * sizeof(float), cudaMemcpyHostToDevice));
        int r2 = pl->R == 32 ? conv_setup<32>(pl, d_ir) : pl->R == 16 ? conv_setup<16>(pl, d_ir) : conv_setup<8>(pl, d_ir);
        cudaFree(d_ir);
        return r2;
    }();
    if (rc) { aes_convreverb_plan_destroy(pl); return rc; }
    *out = pl;
    return 0;
}

AES_EXPORT int aes_convreverb_plan_destroy(aes_convreverb_plan *pl)
{
    if (!pl) return 0;
    if (pl->d_tw) cudaFree(pl->d_tw);
    if (pl->d_H) cudaFree(pl->d_H);
    if (pl->d_work) cudaFree(pl->d_work);
    if (pl->d_in) cudaFree(pl->d_in);
    if (pl->d_out) cudaFree(pl->d_out);
    delete pl;
    return 0;
}

template <int PC>
static void mac_launch(aes_convreverb_plan *pl, const ConvArgs &a, cudaStream_t st)
{
    const long long ngroups = (a.B + AESC_CLIPS_PER_ITEM - 1) / AESC_CLIPS_PER_ITEM;
    const long long nitems = (long long)((pl->NS + AESC_MT - 1) / AESC_MT) * ngroups;
    const unsigned grid = (unsigned)std::min<long long>(nitems, 2LL * pl->sms);
    for (int p0 = 0; p0 < pl->Ppad && p0 < a.nblk; p0 += PC) {
        aesc_mac_kernel<PC><<<grid, AESC_MT, mac_smem(PC), st>>>(a, pl->NS, p0, AESC_CLIPS_PER_ITEM, p0 > 0);
        aes_count_launch();
    }
}

template <int R>
static int conv_launch(aes_convreverb_plan *pl, ConvArgs a, cudaStream_t st)
{
    using G = AescGeo<R>;
    const int fft_smem = G::SMEM_CPX * (int)sizeof(cpx);
    const unsigned nb = (unsigned)(a.B * a.nblk);
    aesc_fwd_kernel<R><<<nb, G::NT, fft_smem, st>>>(a);
    if (pl->PC == 18) mac_launch<18>(pl, a, st);
    else if (pl->PC == 9) mac_launch<9>(pl, a, st);
    else mac_launch<4>(pl, a, st);
    aesc_inv_kernel<R><<<nb, G::NT, fft_smem, st>>>(a);
    aes_count_launch(); aes_count_launch();
    AES_CUDA(cudaGetLastError());
    return 0;
}

// x, y: device pointers, (n_clips, n_frames, 2) float32
AES_EXPORT int aes_convreverb_run(aes_convreverb_plan *pl, const float *x, float *y, int64_t n_clips,
                                  int64_t n_frames, double mix_dry, double mix_wet, void *stream)
{
    AES_REQUIRE(pl != nullptr, "plan is NULL");
    if (n_clips <= 0 || n_frames <= 0) return 0;
    AES_REQUIRE(x != nullptr && y != nullptr, "NULL device buffer");
    const int nblk = (int)((n_frames + pl->BK - 1) / pl->BK);
    const size_t per_clip = (size_t)nblk * pl->NS * sizeof(cpx) * 2;
    int64_t chunk = std::max<int64_t>(1, (int64_t)(AESC_WORK_LIMIT / per_clip));
    chunk = std::min<int64_t>(chunk, n_clips);
    const size_t need = per_clip * (size_t)chunk;
    if (pl->work_cap < need) {
        if (pl->d_work) cudaFree(pl->d_work);
        pl->d_work = nullptr; pl->work_cap = 0;
        AES_CUDA(cudaMalloc(&pl->d_work, need));
        pl->work_cap = need;
    }
    for (int64_t b0 = 0; b0 < n_clips; b0 += chunk) {
        const int64_t nb = std::min<int64_t>(chunk, n_clips - b0);
        ConvArgs a;
        a.x = x + (size_t)b0 * n_frames * 2; a.y = y + (size_t)b0 * n_frames * 2;
        a.Z = (cpx *)pl->d_work; a.W = a.Z + (size_t)chunk * nblk * pl->NS;
        a.H = pl->d_H; a.tw = pl->d_tw;
        a.B = nb; a.Nf = n_frames; a.nblk = nblk; a.P = pl->Ppad;
        a.dry = (float)mix_dry; a.wet = (float)mix_wet;
        int rc = pl->R == 32 ? conv_launch<32>(pl, a, (cudaStream_t)stream)
               : pl->R == 16 ? conv_launch<16>(pl, a, (cudaStream_t)stream) : conv_launch<8>(pl, a, (cudaStream_t)stream);
        if (rc) return rc;
    }
    return 0;
}

AES_EXPORT int aes_convreverb_process_host(aes_convreverb_plan *pl, const float *x_host, float *y_host,
                                           int64_t n_clips, int64_t n_frames, double mix_dry, double mix_wet)
{
    AES_REQUIRE(pl != nullptr, "plan is NULL");
    if (n_clips <= 0 || n_frames <= 0) return 0;
    AES_REQUIRE(x_host != nullptr && y_host != nullptr, "NULL host buffer");
    const size_t bytes = (size_t)n_clips * n_frames * 2 * sizeof(float);
    if (pl->io_cap < bytes) {
        if (pl->d_in) cudaFree(pl->d_in);
        if (pl->d_out) cudaFree(pl->d_out);
        pl->d_in = pl->d_out = nullptr; pl->io_cap = 0;
        AES_CUDA(cudaMalloc(&pl->d_in, bytes));
        AES_CUDA(cudaMalloc(&pl->d_out, bytes));
        pl->io_cap = bytes;
    }
    AES_CUDA(cudaMemcpy(pl->d_in, x_host, bytes, cudaMemcpyHostToDevice));
    int rc = aes_convreverb_run(pl, (const float *)pl->d_in, (float *)pl->d_out, n_clips, n_frames, mix_dry, mix_wet, nullptr);
    if (rc) return rc;
    AES_CUDA(cudaMemcpy(y_host, pl->d_out, bytes, cudaMemcpyDeviceToHost));
    return 0;
}

AES_EXPORT int aes_convreverb_plan_info(const aes_convreverb_plan *pl, int *fft_size, int *partitions)
{
    AES_REQUIRE(pl != nullptr, "plan is NULL");
    if (fft_size) *fft_size = pl->N;
    if (partitions) *partitions = pl->P;
    return 0;
}
