/*
 * aes_oracle.c -- CPU ORACLE (test infrastructure, NOT the product).
 *
 * Plain-C restatement of the per-sample loops of the reference's offline
 * effect-chain path (javierdrp/audio-effects-simulator, src/audioblocks).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product (the CUDA
 * library under audio-effects-simulator_b200/csrc) never links or calls it.
 *
 * Parity status: the reference ships no tests or golden vectors, so this
 * oracle is pinned against outputs of the reference itself, generated in the
 * build container by tests/golden/make_golden.py (which imports the Python
 * reference from /root/reference) and committed under tests/golden/.
 *
 * Typing follows what numba infers for the reference kernels (f32 storage,
 * f64 arithmetic wherever a Python float coefficient takes part).  Built
 * twice from this one source: strict (-O2 -ffp-contract=off, the checker)
 * and fast (-O3 -ffast-math, the timed CPU baseline: the reference kernels
 * are numba fastmath=True).
 *
 * All 2-D signals are (frames, channels) float32, row-major, addressed with
 * an explicit element stride so that column views x[:, c:c+1] of an
 * interleaved buffer can be passed like the reference does.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

/* Python's % for a possibly negative numerator and positive modulus. */
static inline int64_t pymod(int64_t a, int64_t m)
{
    int64_t r = a % m;
    return r < 0 ? r + m : r;
}

/* ---- delay.py:7-22  delay_kernel ------------------------------------- */
ORC_API int64_t orc_delay_kernel(float *buf, int64_t w, int64_t size,
                                 const float *x, int64_t xs,
                                 float *wet, int64_t ws,
                                 int64_t n_frames, int64_t dS, double feedback)
{
    for (int64_t n = 0; n < n_frames; ++n) {
        int64_t r = pymod(w - dS, size);
        float delayed = buf[r];
        wet[n * ws] = delayed;
        /* f32 + f32*f64 -> f64, stored to the f32 ring */
        buf[w] = (float)((double)x[n * xs] + (double)delayed * feedback);
        if (++w == size) w = 0;
    }
    return w;
}

/* ---- filter.py:8-40  biquad_kernel (Direct Form I) ------------------- */
/* state: (channels, 4) f32 = [x1, x2, y1, y2]; y1/y2 are f64 inside a call */
ORC_API void orc_biquad_kernel(const float *x, float *y, int64_t frames,
                               int64_t channels, double b0, double b1,
                               double b2, double a1, double a2, float *state)
{
    for (int64_t c = 0; c < channels; ++c) {
        float x1 = state[c * 4 + 0], x2 = state[c * 4 + 1];
        double y1 = state[c * 4 + 2], y2 = state[c * 4 + 3];
        for (int64_t i = 0; i < frames; ++i) {
            float x0 = x[i * channels + c];
            double y0 = b0 * x0 + b1 * x1 + b2 * x2 - a1 * y1 - a2 * y2;
            y[i * channels + c] = (float)y0;
            x2 = x1; x1 = x0; y2 = y1; y1 = y0;
        }
        state[c * 4 + 0] = x1; state[c * 4 + 1] = x2;
        state[c * 4 + 2] = (float)y1; state[c * 4 + 3] = (float)y2;
    }
}

/* ---- octaver.py:9-15  cubic_interp (4-point Hermite) ----------------- */
static inline double hermite4(double t, float y0, float y1, float y2, float y3)
{
    double c0 = y1;
    double c1 = 0.5 * (double)(float)(y2 - y0);           /* f32 difference */
    double c2 = (double)y0 - 2.5 * y1 + 2.0 * y2 - 0.5 * y3;
    double c3 = 0.5 * (double)(float)(y3 - y0) + 1.5 * (double)(float)(y1 - y2);
    return ((c3 * t + c2) * t + c1) * t + c0;
}

/* ---- octaver.py:17-82  pitch_shift_kernel_cubic ---------------------- */
ORC_API void orc_pitch_shift_kernel(float *buf, int64_t *w_io, int64_t size,
                                    const float *x, int64_t xs,
                                    float *out, int64_t os, int64_t frames,
                                    double *phasor_io, double step)
{
    const double two_pi = 2.0 * M_PI;
    int64_t w = *w_io;
    double ph = *phasor_io;
    const double fsize = (double)size;
    for (int64_t i = 0; i < frames; ++i) {
        buf[w] = x[i * xs];
        double p1 = ph, p2 = ph + 0.5;
        if (p2 >= 1.0) p2 -= 1.0;

        double raw1 = (double)w - p1 * fsize + fsize;
        int64_t k1 = (int64_t)raw1;
        double f1 = raw1 - (double)k1;
        double s1 = hermite4(f1, buf[pymod(k1 - 1, size)], buf[pymod(k1, size)],
                             buf[pymod(k1 + 1, size)], buf[pymod(k1 + 2, size)]);

        double raw2 = (double)w - p2 * fsize + fsize;
        int64_t k2 = (int64_t)raw2;
        double f2 = raw2 - (double)k2;
        double s2 = hermite4(f2, buf[pymod(k2 - 1, size)], buf[pymod(k2, size)],
                             buf[pymod(k2 + 1, size)], buf[pymod(k2 + 2, size)]);

        double g1 = 0.5 * (1.0 - cos(two_pi * p1));
        double g2 = 0.5 * (1.0 - cos(two_pi * p2));
        out[i * os] = (float)(s1 * g1 + s2 * g2);

        if (++w >= size) w = 0;
        ph += step;
        if (ph >= 1.0) ph -= 1.0;
        else if (ph < 0.0) ph += 1.0;
    }
    *w_io = w;
    *phasor_io = ph;
}

/* ---- reverb.py:11-31  pure_delay_kernel ------------------------------ */
ORC_API int64_t orc_pure_delay_kernel(float *buf, int64_t w, int64_t size,
                                      const float *x, int64_t xs,
                                      float *y, int64_t ys, int64_t n_frames,
                                      int64_t dS)
{
    for (int64_t n = 0; n < n_frames; ++n) {
        float v = x[n * xs];
        y[n * ys] = dS == 0 ? v : buf[pymod(w - dS, size)];
        buf[w] = v;
        if (++w == size) w = 0;
    }
    return w;
}

/* ---- reverb.py:33-46  comb_damped_kernel ----------------------------- */
ORC_API int64_t orc_comb_damped_kernel(float *buf, int64_t w, int64_t size,
                                       const float *x, int64_t xs,
                                       float *y, int64_t ys, int64_t n_frames,
                                       int64_t dS, double g, double h,
                                       double *lp_io)
{
    double lp = *lp_io;
    for (int64_t n = 0; n < n_frames; ++n) {
        float d = buf[pymod(w - dS, size)];
        double damped = (1.0 - h) * (double)d + h * lp;
        lp = damped;
        y[n * ys] = d;
        buf[w] = (float)((double)x[n * xs] + g * damped);
        if (++w == size) w = 0;
    }
    *lp_io = lp;
    return w;
}

/* ---- reverb.py:48-67  allpass_kernel --------------------------------- */
ORC_API int64_t orc_allpass_kernel(float *buf, int64_t w, int64_t size,
                                   const float *x, int64_t xs,
                                   float *y, int64_t ys, int64_t n_frames,
                                   int64_t dS, double a)
{
    for (int64_t n = 0; n < n_frames; ++n) {
        float d = buf[pymod(w - dS, size)];
        float xi = x[n * xs];
        double yo = (double)d - a * (double)xi;
        y[n * ys] = (float)yo;
        buf[w] = (float)((double)xi + a * yo);
        if (++w == size) w = 0;
    }
    return w;
}

/* ---- gate.py:6-42  gate_kernel --------------------------------------- */
ORC_API double orc_gate_kernel(const float *x, float *y, int64_t frames,
                               int64_t channels, double gain, double thresh,
                               double att, double rel)
{
    for (int64_t i = 0; i < frames; ++i) {
        double lvl = 0.0;
        for (int64_t c = 0; c < channels; ++c) {
            float a = fabsf(x[i * channels + c]);
            if ((double)a > lvl) lvl = a;
        }
        double target = lvl > thresh ? 1.0 : 0.0;
        if (gain < target)
            gain = (1.0 - att) * gain + att * target;
        else
            gain = (1.0 - rel) * gain + rel * target;
        for (int64_t c = 0; c < channels; ++c)
            y[i * channels + c] = (float)((double)x[i * channels + c] * gain);
    }
    return gain;
}

/* ==== numpy glue of the effect wrappers, f32 elementwise =============== */

/* delay.py:94-96 / reverb.py:275-277:
 *   out[:, c] = clip(f32(dry)*x[:, c] + f32(wet)*wetsig, -1, 1)
 * numpy evaluates the two products and the sum as separate f32 ufuncs. */
ORC_API void orc_mix_clip(const float *x, int64_t xs, const float *wetsig,
                          int64_t ws, float *out, int64_t os, int64_t n_frames,
                          float dry, float wet)
{
    for (int64_t n = 0; n < n_frames; ++n) {
        float a = dry * x[n * xs];
        float b = wet * wetsig[n * ws];
        float v = a + b;
        v = v < -1.0f ? -1.0f : (v > 1.0f ? 1.0f : v);
        out[n * os] = v;
    }
}

/* octaver.py:124-126  np.mean(x, axis=1) over 2 f32 channels */
ORC_API void orc_mono_mean2(const float *x, float *mono, int64_t n_frames)
{
    for (int64_t n = 0; n < n_frames; ++n) {
        float s = x[2 * n] + x[2 * n + 1];
        mono[n] = s / 2.0f;
    }
}

/* octaver.py:146-150  out[:, ch] = x[:, ch]*f32(1-mix) + wet*f32(mix) */
ORC_API void orc_octaver_mix(const float *x, const float *wetsig, float *out,
                             int64_t n_frames, int64_t channels,
                             float dry_gain, float wet_gain)
{
    for (int64_t n = 0; n < n_frames; ++n) {
        float b = wetsig[n] * wet_gain;
        for (int64_t c = 0; c < channels; ++c) {
            float a = x[n * channels + c] * dry_gain;
            out[n * channels + c] = a + b;
        }
    }
}

/* reverb.py:241,261  sum += tmp   (f32 accumulate) */
ORC_API void orc_accumulate(float *acc, const float *v, int64_t n)
{
    for (int64_t i = 0; i < n; ++i) acc[i] = acc[i] + v[i];
}

/* engine.py:104-105  clip(-1,1); (x*32767).astype(int16)  (trunc toward 0) */
ORC_API void orc_quantize_i16(const float *x, int16_t *q, int64_t n)
{
    for (int64_t i = 0; i < n; ++i) {
        float v = x[i];
        v = v < -1.0f ? -1.0f : (v > 1.0f ? 1.0f : v);
        float s = v * 32767.0f;
        q[i] = (int16_t)s;
    }
}

/* ==== whole-chain driver used for the timed CPU baseline ================
 * One clip through  delay -> reverb  style chains entirely in C, so that the
 * multi-threaded baseline (OpenMP over clips) does not sit behind the GIL.
 * The per-block semantics are the same restated loops as above; the Python
 * wrapper (oracle/oracle.py) resolves every parameter and lays out `ops`.
 *
 * op layout (double p[32], int64 q[32]) is documented in oracle/oracle.py.
 */
enum { ORC_OP_DELAY = 1, ORC_OP_REVERB = 2, ORC_OP_BIQUAD = 3, ORC_OP_GATE = 4,
       ORC_OP_OCTAVER = 5, ORC_OP_DISTORTION = 6 };

typedef struct {
    int32_t kind;
    int32_t pad;
    double p[32];
    int64_t q[32];
} orc_op;

#include <stdlib.h>

static void run_delay(const orc_op *op, const float *x, float *y, int64_t N)
{
    /* q[0]=size q[1]=dS_L q[2]=dS_R ; p[0]=fb p[1]=dry p[2]=wet */
    int64_t size = op->q[0];
    float *ring = (float *)calloc((size_t)size, sizeof(float));
    float *wet = (float *)malloc((size_t)N * sizeof(float));
    for (int c = 0; c < 2; ++c) {
        memset(ring, 0, (size_t)size * sizeof(float));
        orc_delay_kernel(ring, 0, size, x + c, 2, wet, 1, N, op->q[1 + c], op->p[0]);
        orc_mix_clip(x + c, 2, wet, 1, y + c, 2, N, (float)op->p[1], (float)op->p[2]);
    }
    free(ring); free(wet);
}

static void run_reverb(const orc_op *op, const float *x, float *y, int64_t N)
{
    /* q[0]=ncomb q[1]=nap q[2]=pre_size q[3]=pre_dS
     * q[4+8*s+c]=comb L (side s, comb c<=7) ; q[20+4*s+k]=allpass L (k<=3)
     * p[0]=dry p[1]=wet p[2]=h p[3]=a ; p[4+8*s+c]=comb gain g */
    int64_t nc = op->q[0], na = op->q[1], pre_size = op->q[2], pre_dS = op->q[3];
    float *pre = (float *)malloc((size_t)N * sizeof(float));
    float *t1 = (float *)malloc((size_t)N * sizeof(float));
    float *t2 = (float *)malloc((size_t)N * sizeof(float));
    float *sum = (float *)malloc((size_t)N * sizeof(float));
    for (int s = 0; s < 2; ++s) {
        float *ring = (float *)calloc((size_t)pre_size, sizeof(float));
        orc_pure_delay_kernel(ring, 0, pre_size, x + s, 2, pre, 1, N, pre_dS);
        free(ring);
        memset(sum, 0, (size_t)N * sizeof(float));
        for (int64_t c = 0; c < nc; ++c) {
            int64_t L = op->q[4 + 8 * s + c];
            double lp = 0.0;
            ring = (float *)calloc((size_t)(L + 1), sizeof(float));
            orc_comb_damped_kernel(ring, 0, L + 1, pre, 1, t1, 1, N, L,
                                   op->p[4 + 8 * s + c], op->p[2], &lp);
            free(ring);
            orc_accumulate(sum, t1, N);
        }
        float *src = sum, *dst = t1;
        for (int64_t k = 0; k < na; ++k) {
            int64_t L = op->q[20 + 4 * s + k];
            ring = (float *)calloc((size_t)(L + 1), sizeof(float));
            orc_allpass_kernel(ring, 0, L + 1, src, 1, dst, 1, N, L, op->p[3]);
            free(ring);
            float *t = src; src = dst; dst = t;
        }
        (void)t2;
        orc_mix_clip(x + s, 2, src, 1, y + s, 2, N, (float)op->p[0], (float)op->p[1]);
    }
    free(pre); free(t1); free(t2); free(sum);
}

static void run_biquad(const orc_op *op, const float *x, float *y, int64_t N)
{
    float st[8] = {0};
    orc_biquad_kernel(x, y, N, 2, op->p[0], op->p[1], op->p[2], op->p[3], op->p[4], st);
}

static void run_gate(const orc_op *op, const float *x, float *y, int64_t N)
{
    /* p[0]=thresh p[1]=att p[2]=rel p[3]=initial gain */
    orc_gate_kernel(x, y, N, 2, op->p[3], op->p[0], op->p[1], op->p[2]);
}

static void run_octaver(const orc_op *op, const float *x, float *y, int64_t N)
{
    /* q[0]=size q[1]=w0 ; p[0]=phasor0 p[1]=step p[2]=mix */
    int64_t size = op->q[0], w = op->q[1];
    double ph = op->p[0];
    float *ring = (float *)calloc((size_t)size, sizeof(float));
    float *mono = (float *)malloc((size_t)N * sizeof(float));
    float *wet = (float *)calloc((size_t)N, sizeof(float));
    orc_mono_mean2(x, mono, N);
    orc_pitch_shift_kernel(ring, &w, size, mono, 1, wet, 1, N, &ph, op->p[1]);
    orc_octaver_mix(x, wet, y, N, 2, (float)(1.0 - op->p[2]), (float)op->p[2]);
    free(ring); free(mono); free(wet);
}

/* Distortion is NOT in the reference: y = clip((1-mix)*x +
 * mix*tanh(drive*x)) in f32, our own definition -- parity unpinned. */
ORC_API void orc_distortion(const float *x, float *y, int64_t n, float drive, float mix)
{
    float dry = 1.0f - mix;
    for (int64_t i = 0; i < n; ++i) {
        float d = drive * x[i];
        float t = tanhf(d);
        float a = dry * x[i];
        float b = mix * t;
        float v = a + b;
        y[i] = v < -1.0f ? -1.0f : (v > 1.0f ? 1.0f : v);
    }
}

/* Fresh-state chain over one stereo clip: x,y are (N,2) f32; tmp is (N,2). */
ORC_API void orc_chain_clip(const orc_op *ops, int32_t n_ops, const float *x,
                            float *y, float *tmp, int64_t N)
{
    const float *src = x;
    float *bufs[2] = { y, tmp };
    /* arrange the ping-pong so the last op lands in y */
    int which = (n_ops % 2 == 1) ? 0 : 1;
    if (n_ops == 0) { memcpy(y, x, (size_t)N * 2 * sizeof(float)); return; }
    for (int32_t k = 0; k < n_ops; ++k) {
        float *dst = bufs[which];
        switch (ops[k].kind) {
        case ORC_OP_DELAY:      run_delay(&ops[k], src, dst, N); break;
        case ORC_OP_REVERB:     run_reverb(&ops[k], src, dst, N); break;
        case ORC_OP_BIQUAD:     run_biquad(&ops[k], src, dst, N); break;
        case ORC_OP_GATE:       run_gate(&ops[k], src, dst, N); break;
        case ORC_OP_OCTAVER:    run_octaver(&ops[k], src, dst, N); break;
        case ORC_OP_DISTORTION: orc_distortion(src, dst, 2 * N, (float)ops[k].p[0], (float)ops[k].p[1]); break;
        default: memcpy(dst, src, (size_t)N * 2 * sizeof(float));
        }
        src = dst;
        which ^= 1;
    }
}

/* Batch of clips, pthreads over clips with a shared atomic clip counter (the
 * reference itself is one thread per clip; this is "all the host threads it
 * can use" for bench --impl reference). */
#include <pthread.h>
typedef struct {
    const orc_op *ops; int32_t n_ops; const float *x; float *y;
    int64_t B, N; int64_t *next;
} batch_job;

static void *batch_worker(void *arg)
{
    batch_job *j = (batch_job *)arg;
    float *tmp = (float *)malloc((size_t)j->N * 2 * sizeof(float));
    for (;;) {
        int64_t b = __atomic_fetch_add(j->next, 1, __ATOMIC_RELAXED);
        if (b >= j->B) break;
        orc_chain_clip(j->ops, j->n_ops, j->x + b * j->N * 2, j->y + b * j->N * 2, tmp, j->N);
    }
    free(tmp);
    return NULL;
}

ORC_API void orc_chain_batch(const orc_op *ops, int32_t n_ops, const float *x,
                             float *y, int64_t B, int64_t N, int32_t threads)
{
    int64_t next = 0;
    batch_job job = { ops, n_ops, x, y, B, N, &next };
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    pthread_t tid[256];
    for (int t = 1; t < threads; ++t) pthread_create(&tid[t], NULL, batch_worker, &job);
    batch_worker(&job);
    for (int t = 1; t < threads; ++t) pthread_join(tid[t], NULL);
}
