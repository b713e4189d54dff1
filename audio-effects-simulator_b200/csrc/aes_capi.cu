// aes_capi.cu -- the remaining C-ABI entry points of include/aesim.h: error state,
// device/memory/stream helpers for hosts without their own CUDA plumbing, and the
// single-block wrappers (one-stage chains).
#include <atomic>
#include <string.h>

#include "aes_common.h"

static thread_local char g_err[1024] = "";
static std::atomic<long long> g_launches{0};

void aes_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

void aes_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

AES_EXPORT int aes_abi_version(void) { return AES_ABI_VERSION; }
AES_EXPORT const char *aes_last_error(void) { return g_err; }
AES_EXPORT int64_t aes_launch_count(void) { return g_launches.load(); }

AES_EXPORT int aes_device_count(int *count)
{
    AES_REQUIRE(count != nullptr, "count is NULL");
    AES_CUDA(cudaGetDeviceCount(count));
    return 0;
}

AES_EXPORT int aes_set_device(int device)
{
    AES_CUDA(cudaSetDevice(device));
    return 0;
}

AES_EXPORT int aes_device_sm_count(int *sms)
{
    AES_REQUIRE(sms != nullptr, "sms is NULL");
    int dev = 0;
    AES_CUDA(cudaGetDevice(&dev));
    AES_CUDA(cudaDeviceGetAttribute(sms, cudaDevAttrMultiProcessorCount, dev));
    return 0;
}

AES_EXPORT int aes_malloc(void **dptr, size_t bytes)
{
    AES_REQUIRE(dptr != nullptr, "dptr is NULL");
    AES_CUDA(cudaMalloc(dptr, bytes ? bytes : 1));
    return 0;
}

AES_EXPORT int aes_free(void *dptr)
{
    if (dptr) AES_CUDA(cudaFree(dptr));
    return 0;
}

AES_EXPORT int aes_host_alloc(void **hptr, size_t bytes)
{
    AES_REQUIRE(hptr != nullptr, "hptr is NULL");
    AES_CUDA(cudaMallocHost(hptr, bytes ? bytes : 1));
    return 0;
}

AES_EXPORT int aes_host_free(void *hptr)
{
    if (hptr) AES_CUDA(cudaFreeHost(hptr));
    return 0;
}

AES_EXPORT int aes_memcpy_h2d(void *dst, const void *src, size_t bytes, void *stream)
{
    AES_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return 0;
}

AES_EXPORT int aes_memcpy_d2h(void *dst, const void *src, size_t bytes, void *stream)
{
    AES_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return 0;
}

AES_EXPORT int aes_memset(void *dst, int value, size_t bytes, void *stream)
{
    AES_CUDA(cudaMemsetAsync(dst, value, bytes, (cudaStream_t)stream));
    return 0;
}

AES_EXPORT int aes_stream_create(void **stream)
{
    AES_REQUIRE(stream != nullptr, "stream is NULL");
    cudaStream_t s;
    AES_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    *stream = (void *)s;
    return 0;
}

AES_EXPORT int aes_stream_destroy(void *stream)
{
    if (stream) AES_CUDA(cudaStreamDestroy((cudaStream_t)stream));
    return 0;
}

AES_EXPORT int aes_stream_sync(void *stream)
{
    AES_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return 0;
}

// ---- single blocks as one-stage chains ---------------------------------------------------
static int run_once(const aes_stage_desc *d, int n, const void *x, void *y, int64_t B, int64_t N, void *stream)
{
    aes_chain_plan *pl = nullptr;
    int rc = aes_chain_plan_create(d, n, 48000, &pl);
    if (rc) return rc;
    rc = aes_chain_run(pl, x, AES_FMT_F32_STEREO, y, AES_FMT_F32_STEREO, B, N, stream);
    if (!rc && cudaStreamSynchronize((cudaStream_t)stream) != cudaSuccess) {
        aes_set_error("chain kernel failed: %s", cudaGetErrorString(cudaGetLastError()));
        rc = AES_ERR_CUDA;
    }
    aes_chain_plan_destroy(pl);
    return rc;
}

AES_EXPORT int aes_delay_f32(const float *x, float *y, int64_t n_clips, int64_t n_frames, int64_t dS_L,
                             int64_t dS_R, double feedback, double mix_dry, double mix_wet, void *stream)
{
    aes_stage_desc d;
    memset(&d, 0, sizeof d);
    d.kind = AES_STAGE_DELAY;
    d.q[0] = dS_L; d.q[1] = dS_R;
    d.p[0] = feedback; d.p[1] = mix_dry; d.p[2] = mix_wet;
    return run_once(&d, 1, x, y, n_clips, n_frames, stream);
}

AES_EXPORT int aes_biquad_cascade_f32(const float *x, float *y, int64_t n_clips, int64_t n_frames,
                                      int n_stages, const double *coeffs5, void *stream)
{
    AES_REQUIRE(n_stages >= 1 && n_stages <= 16 && coeffs5 != nullptr, "1..16 biquad stages");
    aes_stage_desc d[16];
    memset(d, 0, sizeof d);
    for (int s = 0; s < n_stages; ++s) {
        d[s].kind = AES_STAGE_BIQUAD;
        for (int i = 0; i < 5; ++i) d[s].p[i] = coeffs5[5 * s + i];
    }
    return run_once(d, n_stages, x, y, n_clips, n_frames, stream);
}

// engine.py:104-105: clip to [-1,1], *32767, truncate toward zero
__global__ void aes_quantize_kernel(const float *__restrict__ x, short *__restrict__ q, long long n)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float v = fminf(fmaxf(x[i], -1.0f), 1.0f);
        q[i] = (short)__float2int_rz(__fmul_rn(v, 32767.0f));
    }
}

AES_EXPORT int aes_quantize_i16(const float *x, int16_t *q, int64_t n_values, void *stream)
{
    if (n_values <= 0) return 0;
    AES_REQUIRE(x != nullptr && q != nullptr, "NULL buffer");
    int sms = 148;
    aes_device_sm_count(&sms);
    aes_quantize_kernel<<<sms * 8, 256, 0, (cudaStream_t)stream>>>(x, q, n_values);
    aes_count_launch();
    AES_CUDA(cudaGetLastError());
    return 0;
}
