// aes_spectral.cu -- host side and kernels of the SpectralFilter block (see aes_spectral.cuh).
#include <algorithm>
#include <math.h>
#include <new>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "aes_common.h"
#include "aes_chain_kernel.cuh"      // AES_DYN_SMEM
#include "aes_spectral.cuh"
#include "aes_spectral_smooth.cuh"

template <int SHAPE> __global__ void __launch_bounds__(AESM_NTC, 3) aesm_cols_fwd_kernel(const __grid_constant__ SmoothArgs a) { aesm_cols_fwd_body<SHAPE>(a); }
// (measured, r2ap: the 960 x 1000 forward column kernel as 512 threads x 2 CTAs / 256 x 3 CTAs, 64 / 80 registers and no
// spill: 24.16 / 25.78 ms per 2048 clips against 22.93 for 512 x 3 with its 16 spilled bytes -- the warps matter more)
template <int SHAPE> __global__ void __launch_bounds__(AESM_NT, 5) aesm_rows_kernel(const __grid_constant__ SmoothArgs a) { aesm_rows_body<SHAPE>(a); }
template <int SHAPE> __global__ void __launch_bounds__(AESM_NT) aesm_cols_inv_kernel(const __grid_constant__ SmoothArgs a) { aesm_cols_inv_body<SHAPE>(a); }
// G row pairs per CTA, CTAS resident CTAs per SM (registers: 65536 / (threads * CTAS))
template <int G, int CTAS, int PP = 2, int PF = 0>
__global__ void __launch_bounds__(AESR_NT_OF(G), CTAS) aesm_rows10_kernel(const __grid_constant__ SmoothArgs a) { aesm_rows10_body<G, PP, PF>(a); }
// tables of the smooth path, in double: W_M^j (j < 1024), W_M^(1024 j), np.hanning(M)
__global__ void aesm_tables_kernel(cpx *twlo, cpx *twhi, long long nhi, float *window, long long M)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    auto w = [](double num, double den) { double sn, cs; sincospi(-2.0 * num / den, &sn, &cs); cpx c; c.x = (float)cs; c.y = (float)sn; return c; };
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < M; e += stride) {
        window[e] = (float)(0.5 - 0.5 * cospi(2.0 * (double)e / (double)(M - 1)));
        if (e < 1024) twlo[e] = w((double)e, (double)M);
        if (e < nhi) twhi[e] = w((double)(e * 1024), (double)M);
    }
}

__global__ void aess_load_kernel(const __grid_constant__ SpecArgs a) { aess_load_body(a); }
template <int R, int IN, int OUT>
__global__ void __launch_bounds__(256) aess_global_pass_kernel(const __grid_constant__ SpecArgs a, int st, int inv) { aess_global_pass_body<R, IN, OUT>(a, st, inv); }
__global__ void __launch_bounds__(AESS_LOCAL_NT, 3) aess_local_kernel(const __grid_constant__ SpecArgs a, int inv, int mul) { aess_local_body(a, inv, mul); }
__global__ void aess_gate_kernel(const __grid_constant__ SpecArgs a) { aess_gate_body(a); }
__global__ void aess_zero_pad_kernel(const __grid_constant__ SpecArgs a) { aess_zero_pad_body(a); }
__global__ void aess_store_kernel(const __grid_constant__ SpecArgs a) { aess_store_body(a); }

// whole-clip, fresh-state helpers (spectral.py:30-42 re-initialises at hop == N):
// frame = [zeros(N), mono * hanning(2N)[N:]], output = first N samples of the irfft, both channels
__global__ void aess_frames_from_clips_kernel(const float *x, const float *window, float *frames, long long nb, long long N)
{
    const long long M = 2 * N, stride = (long long)gridDim.x * blockDim.x;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < nb * M; e += stride) {
        const long long b = e / M, n = e % M;
        float v = 0.f;
        if (n >= N) {
            const float2 f = reinterpret_cast<const float2 *>(x)[b * N + (n - N)];
            v = __fmul_rn(__fmul_rn(__fadd_rn(f.x, f.y), 0.5f), window[n]);      // np.mean then * window, f32
        }
        frames[e] = v;
    }
}
__global__ void aess_clips_from_frames_kernel(const float *out, float *y, long long nb, long long N)
{
    const long long M = 2 * N, stride = (long long)gridDim.x * blockDim.x;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < nb * N; e += stride) {
        const long long b = e / N, n = e % N;
        const float v = out[b * M + n];
        reinterpret_cast<float2 *>(y)[e] = make_float2(v, v);
    }
}
// plan tables, built on the device (on the host the 2M+P sincos calls of a whole-file plan cost
// ~150 ms per request of the file route): chirp c[n] = exp(-i*pi*n^2/M) with n^2 reduced mod 2M
// in exact integer arithmetic, the Hann window np.hanning(M) as float32, and the Bluestein kernel
// v[m] = conj(c[|m|]) / P laid out on [0, P)
__global__ void aess_tables_kernel(cpx *chirp, float *window, cpx *v, long long M, long long P)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    const double invP = 1.0 / (double)P;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < P; e += stride) {
        cpx out; out.x = 0.f; out.y = 0.f;
        const long long m = e < M ? e : (e > P - M ? P - e : -1);
        if (m >= 0) {
            const unsigned long long r = ((unsigned long long)m * (unsigned long long)m) % (unsigned long long)(2 * M);
            double sn, cs;
            sincospi(-(double)r / (double)M, &sn, &cs);
            if (e < M) {
                cpx c; c.x = (float)cs; c.y = (float)sn;
                chirp[e] = c;
                window[e] = (float)(0.5 - 0.5 * cospi(2.0 * (double)e / (double)(M - 1)));
            }
            out.x = (float)((double)(float)cs * invP);
            out.y = (float)(-(double)(float)sn * invP);
        }
        v[e] = out;
    }
}
__global__ void aess_fill_kernel(float *p, float v, long long n)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) p[e] = v;
}

struct aes_spectral_plan {
    long long M = 0, P = 0;
    int L = 0, sms = 148;
    cpx *d_vhat = nullptr, *d_chirp = nullptr, *d_tw1k = nullptr;
    float *d_window = nullptr;
    void *d_work = nullptr;          // buf | frames | mask | out | (x | y for the clip entry)
    size_t work_cap = 0;
    // frame lengths n1 * n2 with factors 2, 3, 5 only skip Bluestein (aes_spectral_smooth.cuh)
    bool smooth = false;
    int n1 = 0, n2 = 0, static_shape = 0;
    int rows10_variant = 0;
    bool rows10 = false;             // 960 x 1000: the register-resident row-pair kernel (AES_SPECTRAL_ROWS_SMEM=1: the shared-memory one)
    SmoothFft f1, f2;
    cpx *d_stw = nullptr;            // stage twiddles of f1 | of f2 | twlo [1024] | twhi [ceil(M / 1024)]
    int *d_rev = nullptr;            // rev1 [n1] | rev2 [n2]
};

static int spec_grid(const aes_spectral_plan *pl) { return pl->sms * 8; }

// Device buffers of the spectral plans come from the device's stream-ordered memory pool with the
// release threshold raised to 1 GB, so the ~150 MB a whole-file plan needs is carved out of memory kept
// from the previous request instead of going to the driver each time (the file route creates and
// drops a plan per request; cudaMalloc + cudaFree of these sizes cost 40-150 ms).
static int spec_alloc(void **p, size_t bytes)
{
    static thread_local int pool_dev = -1;
    int dev = 0;
    AES_CUDA(cudaGetDevice(&dev));
    if (pool_dev != dev) {
        cudaMemPool_t pool;
        AES_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
        unsigned long long keep = 1ULL << 30;        // a whole-file plan is ~150 MB; batch work buffers (GBs) still go back
        AES_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
        pool_dev = dev;
    }
    AES_CUDA(cudaMallocAsync(p, bytes, (cudaStream_t)0));
    return 0;
}
static void spec_free(void *p) { if (p) cudaFreeAsync(p, (cudaStream_t)0); }

// `in_mode` applies to the first pass of a forward transform, `out_mode` to the last pass of an
// inverse one (see aess_global_pass_body); both need at least one grid-wide stage (L > 10)
template <int IN, int OUT>
static void spec_pass(const aes_spectral_plan *pl, const SpecArgs &a, int s, int r, int inv, cudaStream_t st)
{
    const int g = spec_grid(pl);
    if (r == 3) aess_global_pass_kernel<3, IN, OUT><<<g, 256, 0, st>>>(a, s, inv);
    else if (r == 2) aess_global_pass_kernel<2, IN, OUT><<<g, 256, 0, st>>>(a, s, inv);
    else aess_global_pass_kernel<1, IN, OUT><<<g, 256, 0, st>>>(a, s, inv);
    aes_count_launch();
}

static int spec_fft(const aes_spectral_plan *pl, const SpecArgs &a, int inverse, int mul, cudaStream_t st,
                    int in_mode = 0, int out_mode = 0)
{
    const long long pairs = ((long long)a.nb * a.P / 1024 + 1) / 2;    // two 1024-point chunks per CTA trip
    const unsigned lgrid = (unsigned)std::min<long long>(pairs, (long long)pl->sms * 8);
    if (!inverse) {                                             // stages 0 .. L-11 span >= 1024 points
        for (int s = 0, r; s <= a.L - 11; s += r) {
            r = aess_pass_radix(a.L - 10 - s);
            if (s == 0 && in_mode == 1) spec_pass<1, 0>(pl, a, s, r, 0, st);
            else if (s == 0 && in_mode == 2) spec_pass<2, 0>(pl, a, s, r, 0, st);
            else spec_pass<0, 0>(pl, a, s, r, 0, st);
        }
        aess_local_kernel<<<lgrid, AESS_LOCAL_NT, AESS_LOCAL_SMEM_CPX * sizeof(cpx), st>>>(a, 0, mul); aes_count_launch();
    } else {
        aess_local_kernel<<<lgrid, AESS_LOCAL_NT, AESS_LOCAL_SMEM_CPX * sizeof(cpx), st>>>(a, 1, 0); aes_count_launch();
        for (int s = 10, r; s < a.L; s += r) {
            r = aess_pass_radix(a.L - s);
            if (s + r == a.L && out_mode == 1) spec_pass<0, 1>(pl, a, s, r, 1, st);
            else spec_pass<0, 0>(pl, a, s, r, 1, st);
        }
    }
    AES_CUDA(cudaGetLastError());
    return 0;
}

// frames (device, windowed) -> out (device, irfft of the gated spectrum); mask updated in place
static int spec_process(const aes_spectral_plan *pl, SpecArgs a, cudaStream_t st)
{
    const int g = spec_grid(pl);
    const bool fuse = a.L > 10;                                 // a grid-wide pass exists to fold load / pad / store into
    int rc;
    if (!fuse) { aess_load_kernel<<<g, 256, 0, st>>>(a); aes_count_launch(); }
    if ((rc = spec_fft(pl, a, 0, 1, st, fuse ? 1 : 0, 0))) return rc;
    if ((rc = spec_fft(pl, a, 1, 0, st))) return rc;
    aess_gate_kernel<<<g, 256, 0, st>>>(a); aes_count_launch();
    if (!fuse) { aess_zero_pad_kernel<<<g, 256, 0, st>>>(a); aes_count_launch(); }
    if ((rc = spec_fft(pl, a, 0, 1, st, fuse ? 2 : 0, 0))) return rc;
    if ((rc = spec_fft(pl, a, 1, 0, st, 0, fuse ? 1 : 0))) return rc;
    if (!fuse) { aess_store_kernel<<<g, 256, 0, st>>>(a); aes_count_launch(); }
    AES_CUDA(cudaGetLastError());
    return 0;
}

static SmoothArgs smooth_args(const aes_spectral_plan *pl)
{
    SmoothArgs a; memset(&a, 0, sizeof a);
    a.f1 = pl->f1; a.f2 = pl->f2; a.n1 = pl->n1; a.n2 = pl->n2; a.M = (int)pl->M;
    a.tw1 = pl->d_stw; a.tw2 = a.tw1 + pl->f1.tsize; a.twlo = a.tw2 + pl->f2.tsize; a.twhi = a.twlo + 1024;
    a.rev1 = pl->d_rev; a.rev2 = pl->d_rev + pl->n1;
    a.window = pl->d_window;
    a.tw10 = a.twhi + (pl->M + 1023) / 1024;
    return a;
}

// the three passes of the four-step path; a.buf / frames / clips / outputs set by the caller
static int smooth_process(const aes_spectral_plan *pl, const SmoothArgs &a, cudaStream_t st)
{
    const long long tiles = (long long)a.np * ((a.n2 + AESM_C - 1) / AESM_C), rows = (long long)a.np * (a.n1 / 2 + 1);
    const size_t sm_c = (size_t)a.n1 * AESM_C * sizeof(cpx), sm_r = (size_t)2 * a.n2 * sizeof(cpx);
    const int per_sm_c = (int)std::max<size_t>(1, std::min<size_t>(8, (size_t)220 * 1024 / (sm_c + 1024)));
    const int per_sm_r = (int)std::max<size_t>(1, std::min<size_t>(8, (size_t)220 * 1024 / (sm_r + 1024)));
    const unsigned gc = (unsigned)std::min<long long>(tiles, (long long)pl->sms * per_sm_c);
    const unsigned gr = (unsigned)std::min<long long>(rows, (long long)pl->sms * per_sm_r);
    if (pl->static_shape == AESM_SHAPE_960x1000) {
        aesm_cols_fwd_kernel<AESM_SHAPE_960x1000><<<gc, AESM_NTC, sm_c, st>>>(a);
        if (pl->rows10) {
            // row pairs 1 .. n1/2 - 1 in registers (aesm_rows10_body); rows 0 and n1/2 pair with themselves
            const long long items = (long long)a.np * (a.n1 / 2 - 1);
            auto launch10 = [&](auto gt, auto ct, auto pt, auto ft) {
                constexpr int G = decltype(gt)::value, CTAS = decltype(ct)::value, PP = decltype(pt)::value, PF = decltype(ft)::value;
                int per_sm = CTAS;                      // what actually fits beside the shared-memory tables
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, aesm_rows10_kernel<G, CTAS, PP, PF>, AESR_NT_OF(G), AESR_SMEM_OF3(G, PP, PF));
                if (per_sm < 1) per_sm = 1;
                const unsigned g10 = (unsigned)std::min<long long>((items + G - 1) / G, (long long)pl->sms * per_sm);
                aesm_rows10_kernel<G, CTAS, PP, PF><<<g10, AESR_NT_OF(G), AESR_SMEM_OF3(G, PP, PF), st>>>(a);
            };
            using std::integral_constant;
            using I0 = integral_constant<int, 0>;
            using I1 = integral_constant<int, 1>; using I2 = integral_constant<int, 2>; using I3 = integral_constant<int, 3>;
            using I4 = integral_constant<int, 4>;
            // Measured on 2048 clips (r2ai, ms per pass of the whole block): <1,3,2> 23.94, <2,2,2> 24.62, <1,4,2> 23.89,
            // <1,4,1> 23.70, <2,3,1> 23.98, <1,5,1> (80 registers, 48 bytes of spills) 23.47; the shared-memory
            // kernel 26.58.  With the next item's rows fetched by bulk copies (r2ak): <1,4,1> 23.34, <1,3,2> 24.19,
            // <2,2,1> 22.91 -- the shipped one.  AES_ROWS10_VARIANT = 3, 5, 8 picks <1,3,2>, <1,4,1> (no prefetch) or
            // <1,4,1> with it for re-measuring; the other shapes are no longer instantiated.
            switch (pl->rows10_variant) {
            case 3: launch10(I1{}, I3{}, I2{}, I0{}); break;
            case 5: launch10(I1{}, I4{}, I1{}, I0{}); break;
            case 8: launch10(I1{}, I4{}, I1{}, I1{}); break;
            default: launch10(I2{}, I2{}, I1{}, I1{}); break;
            }
            SmoothArgs b = a;
            b.rows_self_only = 1;
            aesm_rows_kernel<AESM_SHAPE_960x1000><<<(unsigned)std::min<long long>(2LL * a.np, gr), AESM_NT, sm_r, st>>>(b);
            aes_count_launch();
        } else {
            aesm_rows_kernel<AESM_SHAPE_960x1000><<<gr, AESM_NT, sm_r, st>>>(a);
        }
        aesm_cols_inv_kernel<AESM_SHAPE_960x1000><<<gc, AESM_NT, sm_c, st>>>(a);
    } else {
        aesm_cols_fwd_kernel<0><<<gc, AESM_NTC, sm_c, st>>>(a);
        aesm_rows_kernel<0><<<gr, AESM_NT, sm_r, st>>>(a);
        aesm_cols_inv_kernel<0><<<gc, AESM_NT, sm_c, st>>>(a);      // 512 threads: 50 registers leave 2 CTAs per SM (3.1 ms against 1.8)
    }
    aes_count_launch(); aes_count_launch(); aes_count_launch();
    AES_CUDA(cudaGetLastError());
    return 0;
}

AES_EXPORT int aes_spectral_plan_destroy(aes_spectral_plan *pl)
{
    if (!pl) return 0;
    cudaDeviceSynchronize();                        // launches on the caller's streams may still use the buffers
    spec_free(pl->d_stw);
    spec_free(pl->d_rev);
    spec_free(pl->d_vhat);
    spec_free(pl->d_chirp);
    spec_free(pl->d_tw1k);
    spec_free(pl->d_window);
    spec_free(pl->d_work);
    delete pl;
    return 0;
}

AES_EXPORT int aes_spectral_plan_create(int64_t frame_len, aes_spectral_plan **out)
{
    AES_REQUIRE(out != nullptr, "NULL argument");
    AES_REQUIRE(frame_len >= 2 && frame_len % 2 == 0 && frame_len <= (1LL << 26), "frame length: even, 2..2^26");
    aes_spectral_plan *pl = new (std::nothrow) aes_spectral_plan();
    if (!pl) { aes_set_error("out of host memory"); return AES_ERR_NOMEM; }
    const long long M = frame_len;
    long long P = 1024; int L = 10;
    while (P < 2 * M - 1) { P <<= 1; ++L; }
    pl->M = M; pl->P = P; pl->L = L;
    pl->smooth = !getenv("AES_SPECTRAL_BLUESTEIN") && aesm_split(M, &pl->n1, &pl->n2);
    int rc = [&]() -> int {
        aes_device_sm_count(&pl->sms);
        if (pl->smooth) {
            aesm_build_fft(pl->n1, &pl->f1); aesm_build_fft(pl->n2, &pl->f2);
            const long long nhi = (M + 1023) / 1024;
            int ra;
            const int nst = pl->f1.tsize + pl->f2.tsize;
            if ((ra = spec_alloc((void **)&pl->d_stw, (size_t)(nst + 1024 + nhi + AESR_TW_ENTRIES) * sizeof(cpx)))) return ra;
            std::vector<cpx> stw((size_t)nst + 1);
            aesm_fill_twiddles(pl->f1, stw.data()); aesm_fill_twiddles(pl->f2, stw.data() + pl->f1.tsize);
            AES_CUDA(cudaMemcpy(pl->d_stw, stw.data(), (size_t)nst * sizeof(cpx), cudaMemcpyHostToDevice));
            if ((ra = spec_alloc((void **)&pl->d_window, (size_t)M * sizeof(float)))) return ra;
            if ((ra = spec_alloc((void **)&pl->d_rev, (size_t)(pl->n1 + pl->n2) * sizeof(int)))) return ra;
            std::vector<int> rev((size_t)pl->n1 + pl->n2);
            for (int k = 0; k < pl->n1; ++k) rev[k] = aesm_rev(k, pl->f1);
            for (int k = 0; k < pl->n2; ++k) rev[pl->n1 + k] = aesm_rev(k, pl->f2);
            AES_CUDA(cudaMemcpy(pl->d_rev, rev.data(), rev.size() * sizeof(int), cudaMemcpyHostToDevice));
            // (kernel attributes, not plan ones: always the limit, plans of other frame lengths may be alive)
            pl->static_shape = getenv("AES_SPECTRAL_GENERIC") ? 0 : aesm_static_shape(pl->f1, pl->f2);
            pl->rows10 = pl->static_shape == AESM_SHAPE_960x1000 && !getenv("AES_SPECTRAL_ROWS_SMEM");
            if (pl->rows10) {
                std::vector<cpx> t10(AESR_TW_ENTRIES);
                aesm_fill_rows10(t10.data());
                AES_CUDA(cudaMemcpy(pl->d_stw + nst + 1024 + nhi, t10.data(), t10.size() * sizeof(cpx), cudaMemcpyHostToDevice));
                if (const char *v = getenv("AES_ROWS10_VARIANT")) pl->rows10_variant = atoi(v);
                AES_CUDA(cudaFuncSetAttribute(aesm_rows10_kernel<1, 3, 2, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, AESR_SMEM_OF3(1, 2, 0)));
                AES_CUDA(cudaFuncSetAttribute(aesm_rows10_kernel<1, 4, 1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, AESR_SMEM_OF3(1, 1, 0)));
                AES_CUDA(cudaFuncSetAttribute(aesm_rows10_kernel<1, 4, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, AESR_SMEM_OF3(1, 1, 1)));
                AES_CUDA(cudaFuncSetAttribute(aesm_rows10_kernel<2, 2, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, AESR_SMEM_OF3(2, 1, 1)));
            }
            AES_CUDA(cudaFuncSetAttribute(aesm_cols_fwd_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            AES_CUDA(cudaFuncSetAttribute(aesm_cols_inv_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            AES_CUDA(cudaFuncSetAttribute(aesm_rows_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            AES_CUDA(cudaFuncSetAttribute(aesm_cols_fwd_kernel<AESM_SHAPE_960x1000>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            AES_CUDA(cudaFuncSetAttribute(aesm_cols_inv_kernel<AESM_SHAPE_960x1000>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            AES_CUDA(cudaFuncSetAttribute(aesm_rows_kernel<AESM_SHAPE_960x1000>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            aesm_tables_kernel<<<spec_grid(pl), 256>>>(pl->d_stw + nst, pl->d_stw + nst + 1024, nhi, pl->d_window, M);
            aes_count_launch();
            AES_CUDA(cudaDeviceSynchronize());
            return 0;
        }
        std::vector<cpx> tw1k(512);
        for (int q = 0; q < 512; ++q) { const double ang = -2.0 * M_PI * q / 1024.0; tw1k[q].x = (float)cos(ang); tw1k[q].y = (float)sin(ang); }
        int ra;
        if ((ra = spec_alloc((void **)&pl->d_chirp, (size_t)M * sizeof(cpx)))) return ra;
        if ((ra = spec_alloc((void **)&pl->d_vhat, (size_t)P * sizeof(cpx)))) return ra;
        if ((ra = spec_alloc((void **)&pl->d_tw1k, tw1k.size() * sizeof(cpx)))) return ra;
        if ((ra = spec_alloc((void **)&pl->d_window, (size_t)M * sizeof(float)))) return ra;
        AES_CUDA(cudaMemcpy(pl->d_tw1k, tw1k.data(), tw1k.size() * sizeof(cpx), cudaMemcpyHostToDevice));
        AES_CUDA(cudaFuncSetAttribute(aess_local_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AESS_LOCAL_SMEM_CPX * (int)sizeof(cpx)));
        aess_tables_kernel<<<spec_grid(pl), 256>>>(pl->d_chirp, pl->d_window, pl->d_vhat, M, P);
        aes_count_launch();
        // vhat = FFT_P(v) / P, kept in bit-reversed order
        SpecArgs a; memset(&a, 0, sizeof a);
        a.buf = pl->d_vhat; a.twP = nullptr; a.tw1k = pl->d_tw1k; a.M = M; a.P = P; a.L = L; a.nb = 1;
        int r2 = spec_fft(pl, a, 0, 0, nullptr);
        if (r2) return r2;
        AES_CUDA(cudaDeviceSynchronize());
        return 0;
    }();
    if (rc) { aes_spectral_plan_destroy(pl); return rc; }
    *out = pl;
    return 0;
}

static int spec_reserve(aes_spectral_plan *pl, size_t bytes)
{
    if (pl->work_cap >= bytes) return 0;
    if (pl->d_work) { cudaDeviceSynchronize(); spec_free(pl->d_work); }
    pl->d_work = nullptr; pl->work_cap = 0;
    { int ra = spec_alloc(&pl->d_work, bytes); if (ra) return ra; }
    AES_CUDA(cudaStreamSynchronize((cudaStream_t)0));       // allocated in legacy-stream order, used on the caller's stream
    pl->work_cap = bytes;
    return 0;
}

static size_t al256(size_t v) { return (v + 255) & ~(size_t)255; }

// General entry (any block size, carried state on the host): in_buffers [nb][M] are the raw
// analysis buffers (the Hann window is applied here), mask [nb][M/2+1] is updated in place,
// y [nb][M] receives irfft(processed spectrum).  Host pointers.
AES_EXPORT int aes_spectral_frames_host(aes_spectral_plan *pl, const float *in_buffers, float *mask, float *y,
                                        int n_frames, double thresh_lin, double reduction, double alpha)
{
    AES_REQUIRE(pl != nullptr && in_buffers != nullptr && mask != nullptr && y != nullptr, "NULL argument");
    if (n_frames <= 0) return 0;
    const long long M = pl->M, nbins = M / 2 + 1;
    const size_t s_buf = al256((size_t)(pl->smooth ? M : pl->P) * sizeof(cpx)), s_fr = al256((size_t)M * 4), s_mk = al256((size_t)nbins * 4);
    int rc = spec_reserve(pl, s_buf + 2 * s_fr + s_mk);
    if (rc) return rc;
    char *w = (char *)pl->d_work;
    if (pl->smooth) {
        SmoothArgs sa = smooth_args(pl);
        float *d_fr = (float *)(w + s_buf), *d_out = (float *)(w + s_buf + s_fr), *d_mask = (float *)(w + s_buf + 2 * s_fr);
        sa.buf = (cpx *)w; sa.frames = d_fr; sa.out = d_out; sa.mask = d_mask; sa.mode = 1; sa.np = 1; sa.nf = 1;
        sa.thr = (float)thresh_lin; sa.red = (float)reduction; sa.alpha = (float)alpha;
        for (int f = 0; f < n_frames; ++f) {
            AES_CUDA(cudaMemcpy(d_fr, in_buffers + (size_t)f * M, (size_t)M * 4, cudaMemcpyHostToDevice));
            AES_CUDA(cudaMemcpy(d_mask, mask + (size_t)f * nbins, (size_t)nbins * 4, cudaMemcpyHostToDevice));
            if ((rc = smooth_process(pl, sa, nullptr))) return rc;
            AES_CUDA(cudaMemcpy(y + (size_t)f * M, d_out, (size_t)M * 4, cudaMemcpyDeviceToHost));
            AES_CUDA(cudaMemcpy(mask + (size_t)f * nbins, d_mask, (size_t)nbins * 4, cudaMemcpyDeviceToHost));
        }
        return 0;
    }
    SpecArgs a; memset(&a, 0, sizeof a);
    a.buf = (cpx *)w; a.frames = (float *)(w + s_buf); a.out = (float *)(w + s_buf + s_fr); a.mask = (float *)(w + s_buf + 2 * s_fr);
    a.vhat = pl->d_vhat; a.chirp = pl->d_chirp; a.twP = nullptr; a.tw1k = pl->d_tw1k;
    a.M = M; a.P = pl->P; a.L = pl->L; a.nb = 1; a.nf = 1;
    a.thr = (float)thresh_lin; a.red = (float)reduction; a.alpha = (float)alpha;
    a.window = pl->d_window;                        // the raw buffer is windowed on load (f32 product like numpy's)
    for (int f = 0; f < n_frames; ++f) {
        const float *src = in_buffers + (size_t)f * M;
        AES_CUDA(cudaMemcpy((void *)a.frames, src, (size_t)M * 4, cudaMemcpyHostToDevice));
        AES_CUDA(cudaMemcpy(a.mask, mask + (size_t)f * nbins, (size_t)nbins * 4, cudaMemcpyHostToDevice));
        if ((rc = spec_process(pl, a, nullptr))) return rc;
        AES_CUDA(cudaMemcpy(y + (size_t)f * M, a.out, (size_t)M * 4, cudaMemcpyDeviceToHost));
        AES_CUDA(cudaMemcpy(mask + (size_t)f * nbins, a.mask, (size_t)nbins * 4, cudaMemcpyDeviceToHost));
    }
    return 0;
}

// Whole-clip entry with the fresh state the file path produces (hop == N, M == 2N): x, y are
// DEVICE pointers (n_clips, n_frames, 2) f32; the plan must have been created with frame_len 2N.
AES_EXPORT int aes_spectral_run(aes_spectral_plan *pl, const float *x, float *y, int64_t n_clips, int64_t n_frames,
                                double thresh_lin, double reduction, double alpha, void *stream)
{
    AES_REQUIRE(pl != nullptr, "plan is NULL");
    if (n_clips <= 0 || n_frames <= 0) return 0;
    AES_REQUIRE(x != nullptr && y != nullptr, "NULL device buffer");
    AES_REQUIRE(pl->M == 2 * n_frames, "plan frame length %lld != 2 * n_frames", pl->M);
    const long long M = pl->M, nbins = M / 2 + 1;
    if (pl->smooth) {
        // one M-point complex buffer per pair of clips is all the state there is.  Chunks of up to 4 GB of it:
        // L2-sized chunks (64 MB, so that the three passes hand the spectrum over on chip) measured 15 % slower,
        // the partial last wave of every small launch costs more than the DRAM trips (AES_SPECTRAL_CHUNK_MB overrides)
        const char *env = getenv("AES_SPECTRAL_CHUNK_MB");
        const size_t budget_s = (size_t)(env ? std::max(1, atoi(env)) : 4096) << 20, pair = (size_t)M * sizeof(cpx);
        const int64_t pairs_all = (n_clips + 1) / 2;
        const int64_t cp = std::max<int64_t>(1, std::min<int64_t>(pairs_all, (int64_t)(budget_s / pair)));
        int rcs = spec_reserve(pl, pair * (size_t)cp);
        if (rcs) return rcs;
        SmoothArgs sa = smooth_args(pl);
        sa.buf = (cpx *)pl->d_work; sa.mode = 2;
        sa.thr = (float)thresh_lin; sa.red = (float)reduction; sa.alpha = (float)alpha;
        for (int64_t p0 = 0; p0 < pairs_all; p0 += cp) {
            const int64_t c0 = 2 * p0, nc = std::min<int64_t>(2 * cp, n_clips - c0);
            sa.clips = x + (size_t)c0 * n_frames * 2; sa.yclips = y + (size_t)c0 * n_frames * 2;
            sa.nf = (int)nc; sa.np = (int)((nc + 1) / 2);
            if ((rcs = smooth_process(pl, sa, (cudaStream_t)stream))) return rcs;
        }
        return 0;
    }
    // per clip: half a transform buffer (two clips share one), frame, output, mask
    const size_t per = al256((size_t)pl->P * sizeof(cpx)) / 2 + 2 * al256((size_t)M * 4) + al256((size_t)nbins * 4);
    const size_t budget = (size_t)8 << 30;
    int64_t chunk = std::max<int64_t>(2, std::min<int64_t>(n_clips + (n_clips & 1), (int64_t)(budget / per) & ~(int64_t)1));
    int rc = spec_reserve(pl, per * (size_t)chunk + al256((size_t)pl->P * sizeof(cpx)));
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int g = spec_grid(pl);
    char *w = (char *)pl->d_work;
    for (int64_t b0 = 0; b0 < n_clips; b0 += chunk) {
        const int64_t nb = std::min<int64_t>(chunk, n_clips - b0);
        SpecArgs a; memset(&a, 0, sizeof a);
        // packed (unaligned-free) layout for nb frames: buf [nb][P] | frames [nb][M] | out [nb][M] | mask [nb][nbins]
        a.buf = (cpx *)w;
        float *frames = (float *)(w + (size_t)((chunk + 1) / 2) * al256((size_t)pl->P * sizeof(cpx)));
        float *outp = frames + (size_t)chunk * M;
        float *maskp = outp + (size_t)chunk * M;
        a.frames = frames; a.out = outp; a.mask = maskp;
        a.vhat = pl->d_vhat; a.chirp = pl->d_chirp; a.twP = nullptr; a.tw1k = pl->d_tw1k;
        a.M = M; a.P = pl->P; a.L = pl->L; a.nb = (int)((nb + 1) / 2); a.nf = (int)nb;      // two clips per complex transform
        a.thr = (float)thresh_lin; a.red = (float)reduction; a.alpha = (float)alpha;
        aess_frames_from_clips_kernel<<<g, 256, 0, st>>>(x + (size_t)b0 * n_frames * 2, pl->d_window, frames, nb, n_frames);
        aess_fill_kernel<<<g, 256, 0, st>>>(maskp, 1.0f, nb * nbins);          // mask_smooth starts at ones
        aes_count_launch(); aes_count_launch();
        if ((rc = spec_process(pl, a, st))) return rc;
        aess_clips_from_frames_kernel<<<g, 256, 0, st>>>(outp, y + (size_t)b0 * n_frames * 2, nb, n_frames);
        aes_count_launch();
        AES_CUDA(cudaGetLastError());
    }
    return 0;
}

AES_EXPORT int aes_spectral_process_host(aes_spectral_plan *pl, const float *x_host, float *y_host, int64_t n_clips,
                                         int64_t n_frames, double thresh_lin, double reduction, double alpha)
{
    AES_REQUIRE(pl != nullptr && x_host != nullptr && y_host != nullptr, "NULL argument");
    if (n_clips <= 0 || n_frames <= 0) return 0;
    const size_t clip_bytes = (size_t)n_frames * 2 * sizeof(float);
    const int64_t step = std::max<int64_t>(1, std::min<int64_t>(n_clips, (int64_t)(((size_t)1 << 30) / clip_bytes)));
    float *dx = nullptr, *dy = nullptr;
    int rc = spec_alloc((void **)&dx, clip_bytes * step);
    if (rc) return rc;
    if ((rc = spec_alloc((void **)&dy, clip_bytes * step))) { spec_free(dx); return rc; }
    AES_CUDA(cudaStreamSynchronize((cudaStream_t)0));
    for (int64_t b0 = 0; b0 < n_clips && !rc; b0 += step) {
        const int64_t nb = std::min<int64_t>(step, n_clips - b0);
        if (cudaMemcpy(dx, x_host + (size_t)b0 * n_frames * 2, clip_bytes * nb, cudaMemcpyHostToDevice) != cudaSuccess) { rc = AES_ERR_CUDA; break; }
        rc = aes_spectral_run(pl, dx, dy, nb, n_frames, thresh_lin, reduction, alpha, nullptr);
        if (!rc && cudaMemcpy(y_host + (size_t)b0 * n_frames * 2, dy, clip_bytes * nb, cudaMemcpyDeviceToHost) != cudaSuccess) rc = AES_ERR_CUDA;
    }
    cudaDeviceSynchronize();
    spec_free(dx); spec_free(dy);
    if (rc == AES_ERR_CUDA) aes_set_error("CUDA copy failed: %s", cudaGetErrorString(cudaGetLastError()));
    return rc;
}
