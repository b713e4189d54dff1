/*
 * aesim.h -- C ABI of libaesim.so, the B200-native (sm_100a) implementation of the
 * offline effect-chain hot path of javierdrp/audio-effects-simulator.
 *
 * This is the drop-in boundary: plain C, pointers and sizes only.  The reference
 * has no FFI of its own for this path (it is Python + numba); the entry points
 * below are what a binding for each reference call site binds.  The Python
 * package audio-effects-simulator_b200/audioblocks (same names and signatures as
 * the reference's src/audioblocks) calls them through ctypes; INTEGRATION.md shows
 * the stub.  Citations are file:line in the reference repository.
 *
 * Conventions: every function returns 0 on success and a negative aes_status on
 * failure; aes_last_error() returns a thread-local message for the last failure.
 * "device pointer" = CUDA global memory owned by the caller (e.g. a torch
 * tensor's data_ptr()); `stream` is a cudaStream_t passed as void* (NULL = the
 * default stream).  Kernels never allocate; a plan owns its scratch.
 * Audio buffers are frame-major: clip b, frame n, channel c at
 * ((b*N + n)*C + c), exactly the reference's (frames, channels) float32 arrays
 * (core.py:81-86) with a leading batch axis.
 */
#ifndef AESIM_H
#define AESIM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AES_ABI_VERSION 1

typedef enum aes_status {
    AES_OK = 0,
    AES_ERR_INVALID = -1,      /* bad argument / unsupported parameter combination */
    AES_ERR_CUDA = -2,         /* CUDA runtime error, text in aes_last_error() */
    AES_ERR_NOMEM = -3,
    AES_ERR_UNSUPPORTED = -4   /* valid in the reference, not (yet) on this path */
} aes_status;

/* Stage kinds: one per reference Effect subclass on the path. */
typedef enum aes_stage_kind {
    AES_STAGE_DELAY = 1,       /* delay.py:43-96   StereoDelayEffect  (kernel delay.py:7-22)      */
    AES_STAGE_REVERB = 2,      /* reverb.py:72-277 ReverbEffect       (kernels reverb.py:11-67)   */
    AES_STAGE_BIQUAD = 3,      /* filter.py:42-113 FilterEffect       (kernel filter.py:8-40)     */
    AES_STAGE_GATE = 4,        /* gate.py:45-90    NoiseGateEffect    (kernel gate.py:6-42)       */
    AES_STAGE_OCTAVER = 5,     /* octaver.py:84-150 OctaverEffect     (kernel octaver.py:17-82)   */
    AES_STAGE_DISTORTION = 6   /* no reference block: clip((1-mix)x + mix*tanh(drive*x)) */
} aes_stage_kind;

/*
 * One resolved effect block.  The host (Python) evaluates the reference's own
 * parameter expressions -- the integer truncations int(fs*ms/1000.0) etc. must be
 * bit-identical to the reference, so they are NOT redone here -- and passes the
 * results:
 *
 *  DELAY      q[0]=dS_L q[1]=dS_R                       (delay.py:38-40,84)
 *             p[0]=feedback p[1]=mix_dry p[2]=mix_wet
 *  REVERB     q[0]=n_comb(<=8) q[1]=n_allpass(<=4) q[2]=pre_dS   (reverb.py:223-225)
 *             q[4+8*s+c]=comb length, q[20+4*s+k]=all-pass length, side s in {0:L,1:R}
 *             p[0]=mix_dry p[1]=mix_wet p[2]=damp(h) p[3]=allpass_gain(a)
 *             p[4+8*s+c]=comb feedback gain g            (reverb.py:205-206)
 *  BIQUAD     p[0..4]=b0,b1,b2,a1,a2 (normalised, filter.py:62-98)
 *             p[8+4*c+{0,1,2,3}]=x1,x2,y1,y2 initial DF-I state of channel c
 *  GATE       p[0]=thresh_lin p[1]=attack_coeff p[2]=release_coeff p[3]=initial gain
 *  OCTAVER    q[0]=ring size q[1]=w0 ; p[0]=phasor0 p[1]=step p[2]=mix
 *  DISTORTION p[0]=drive p[1]=mix
 */
typedef struct aes_stage_desc {
    int32_t kind;
    int32_t flags;
    double  p[32];
    int64_t q[32];
} aes_stage_desc;

/* Sample formats at the two ends of a chain run. */
typedef enum aes_format {
    AES_FMT_F32_STEREO = 0,    /* (B,N,2) float32                                       */
    AES_FMT_F32_MONO = 1,      /* (B,N,1) float32, fanned out to L=R (core.py:147-149)  */
    AES_FMT_I16_STEREO_DOWNMIX = 2, /* (B,N,2) int16 PCM -> /32768 -> mean -> L=R (engine.py:78-84); input only */
    AES_FMT_I16_STEREO = 3     /* output: clip, *32767, truncate (engine.py:104-105)    */
} aes_format;

typedef struct aes_chain_plan aes_chain_plan;

/* ---- library / device ------------------------------------------------------------ */
int         aes_abi_version(void);
const char *aes_last_error(void);
int aes_device_count(int *count);
int aes_set_device(int device);
int aes_device_sm_count(int *sms);

/* ---- memory and streams, for hosts that do not bring their own (no torch needed) -- */
int aes_malloc(void **dptr, size_t bytes);
int aes_free(void *dptr);
int aes_host_alloc(void **hptr, size_t bytes);      /* pinned */
int aes_host_free(void *hptr);
int aes_memcpy_h2d(void *dst, const void *src, size_t bytes, void *stream);
int aes_memcpy_d2h(void *dst, const void *src, size_t bytes, void *stream);
int aes_memset(void *dst, int value, size_t bytes, void *stream);
int aes_stream_create(void **stream);
int aes_stream_destroy(void *stream);
int aes_stream_sync(void *stream);

/* ---- fused effect chain: replaces EffectsChain.process (core.py:138-161) over a
 *      whole clip, as engine.py:101-102 calls it, for a batch of B clips ---------- */
int aes_chain_plan_create(const aes_stage_desc *stages, int n_stages, int sample_rate,
                          aes_chain_plan **plan);
int aes_chain_plan_destroy(aes_chain_plan *plan);
/* x, y: device pointers.  Fresh block state at the start of every clip (what the
 * chain's re-prepare at the file's frame count produces), except the scalar state
 * carried in the descriptors (octaver phase, gate gain, biquad state).
 * A plan owns the working memory of its launches (the feedback-delay lines of its CTAs, the scan's
 * records): launches of ONE plan must be ordered with each other -- the same stream, or events
 * between streams.  Launches that are to overlap take a plan each (the chain of core.py:123-161 is
 * not re-entrant either). */
int aes_chain_run(aes_chain_plan *plan, const void *x, int in_fmt, void *y, int out_fmt,
                  int64_t n_clips, int64_t n_frames, void *stream);
/* Same call with HOST buffers: stages through pinned memory in sub-batches on
 * internal streams (H2D, kernel, D2H overlapped) and returns when y is complete. */
int aes_chain_process_host(aes_chain_plan *plan, const void *x_host, int in_fmt, void *y_host,
                           int out_fmt, int64_t n_clips, int64_t n_frames);
/* Carried scalars of stage `stage` at the end of the clip, after the last
 * aes_chain_process_host call with n_clips == 1 (what the reference keeps in the
 * effect object between calls): BIQUAD out16[4*c+{0,1,2,3}] = x1,x2,y1,y2 of channel c
 * (filter.py:35-40); GATE out16[0] = gain (gate.py:42). */
int aes_chain_final_state(aes_chain_plan *plan, int stage, double *out16);
/* Introspection for the benchmark / tests. */
int aes_chain_plan_info(const aes_chain_plan *plan, int *tile_frames, int *smem_bytes,
                        int *ctas_per_sm, int64_t *scratch_bytes_per_cta);
/* Which kernel the plan launches ("aes_fast_kernel<...>" or the generic interpreter). */
const char *aes_chain_plan_kernel_name(const aes_chain_plan *plan);
/* Number of kernel launches issued by this library since load (bench "gpu_launches"). */
int64_t aes_launch_count(void);

/* ---- block streaming with carried state (the chain's warm-up blocks core.py:131-136 and the
 *      live 256-frame callback engine.py:156-163): any block size, reference ring layout, the
 *      reference's per-sample loops in f64.  Extra descriptor fields: q[28] = device pointer of the
 *      stage's state blob (DELAY ringL[size] ringR[size]; REVERB double lp[2][8] then per side
 *      pre[size], comb rings L+1, all-pass rings L+1; OCTAVER ring[size]), q[29] = frames since
 *      prepare(), q[30] = ring size (delay / reverb pre-delay).  `stages` is updated in place
 *      (q[29], carried scalars) so the caller can keep streaming.  Host buffers. */
int aes_stream_process_host(aes_stage_desc *stages, int n_stages, const float *x_host, int channels_in,
                            float *y_host, int64_t frames);

/* The host-buffer entries keep their pipeline resources (streams, device and pinned staging,
 * line scratch) per calling host thread and reuse them across plans -- the file route builds a
 * new chain, hence a new plan, for every request (engine.py:86-99).  This releases the calling
 * thread's cache; a thread that made host-buffer calls should call it before it ends. */
int aes_release_host_cache(void);

/* ---- single blocks, device pointers (one-stage chains; same semantics) ---------- */
int aes_delay_f32(const float *x, float *y, int64_t n_clips, int64_t n_frames,
                  int64_t dS_L, int64_t dS_R, double feedback, double mix_dry, double mix_wet,
                  void *stream);
int aes_biquad_cascade_f32(const float *x, float *y, int64_t n_clips, int64_t n_frames,
                           int n_stages, const double *coeffs5, void *stream);
int aes_quantize_i16(const float *x, int16_t *q, int64_t n_values, void *stream);

/* ---- file-route reply serialisation (host buffers, host threads; engine.py:107-123) ----------
 *      The reference's `file_processed` message carries both signals as JSON float lists,
 *      `json.dumps(mono.flatten().tolist())` and `json.dumps(processed.mean(axis=1).flatten().tolist())`
 *      -- 2.05 s of the 2.3 s request on the shipped clip.  These write the identical
 *      text ("[a, b, ...]", every element printed as Python prints float(np.float32)) from the
 *      float32 buffers.  `out` must hold aes_json_float_list_bound(n) bytes; the return value is
 *      the number of bytes written (no terminator) or a negative aes_status.  threads <= 0: all
 *      host threads.  The stereo variant averages each (L, R) frame the way numpy's float32
 *      mean(axis=1) does before printing. */
int64_t aes_json_float_list_bound(int64_t n_values);
int64_t aes_json_float_list(const float *x_host, int64_t n_values, char *out, int64_t cap, int threads);
int64_t aes_json_stereo_mean_list(const float *xy_host, int64_t n_frames, char *out, int64_t cap, int threads);

/* ---- IR-convolution reverb (BASELINE configs[3]; no counterpart in the reference, whose
 *      reverb.py:72-277 is a Schroeder network): out = clip(dry*x + wet*(x (*) ir)) per channel.
 *      ir_host: (n_taps, 2) float32 host array; block_log2: FFT size 2^14 (0 = default), 2^11, 2^8. */
typedef struct aes_convreverb_plan aes_convreverb_plan;
int aes_convreverb_plan_create(const float *ir_host, int64_t n_taps, int block_log2, aes_convreverb_plan **plan);
int aes_convreverb_plan_destroy(aes_convreverb_plan *plan);
int aes_convreverb_run(aes_convreverb_plan *plan, const float *x, float *y, int64_t n_clips, int64_t n_frames,
                       double mix_dry, double mix_wet, void *stream);           /* device pointers */
int aes_convreverb_process_host(aes_convreverb_plan *plan, const float *x_host, float *y_host, int64_t n_clips,
                                int64_t n_frames, double mix_dry, double mix_wet);
int aes_convreverb_plan_info(const aes_convreverb_plan *plan, int *fft_size, int *partitions);

/* ---- SpectralFilter (spectral.py:44-100): rfft of a Hann-windowed frame of M = 2*hop samples,
 *      per-bin magnitude gate with a smoothed mask, irfft.  M is arbitrary (even): Bluestein. */
typedef struct aes_spectral_plan aes_spectral_plan;
int aes_spectral_plan_create(int64_t frame_len, aes_spectral_plan **plan);
int aes_spectral_plan_destroy(aes_spectral_plan *plan);
/* one call of SpectralFilter.process_into per frame (spectral.py:60-77), host pointers:
 * in_buffers [n][M] raw analysis buffers, mask [n][M/2+1] in/out, y [n][M] = irfft(processed) */
int aes_spectral_frames_host(aes_spectral_plan *plan, const float *in_buffers, float *mask, float *y, int n_frames,
                             double thresh_lin, double reduction, double alpha);
/* whole clips from the freshly re-initialised state (hop == n_frames, M == 2*n_frames): device pointers */
int aes_spectral_run(aes_spectral_plan *plan, const float *x, float *y, int64_t n_clips, int64_t n_frames,
                     double thresh_lin, double reduction, double alpha, void *stream);
int aes_spectral_process_host(aes_spectral_plan *plan, const float *x_host, float *y_host, int64_t n_clips,
                              int64_t n_frames, double thresh_lin, double reduction, double alpha);

/* ---- plot-side analysis (assets/02_custom.js:65-154): what the page computes for every
 *      plot refresh from the last FFT_SIZE = 16384 samples of the original and of the processed signal
 *      (02_custom.js:179-184) -- 4-term Blackman-Harris window, FFT, 20*log10(|X|/n_fft + 1e-9) for the
 *      n_fft/2+1 bins, the 12-bin chromagram of calculateChroma and the peak frequency above 60 Hz.
 *      a, b: [n_pairs][n_samples] mono float32 signals (the last n_fft samples of each are analysed; b may
 *      equal a); outputs mag_db / mag_lin [n_pairs][2][n_fft/2+1], chroma [n_pairs][2][12],
 *      peak_freq [n_pairs][2], index [.][0] = a, [.][1] = b.  n_fft: 16384, 2048 or 256.  Both signals of a
 *      pair ride through ONE complex transform.  Device pointers / host pointers (mag_lin may be NULL in the
 *      host entry). */
int aes_spectrum_chroma(const float *a, const float *b, int64_t n_pairs, int64_t n_samples, int n_fft,
                        double sample_rate, float *mag_db, float *mag_lin, float *chroma, float *peak_freq,
                        void *stream);
int aes_spectrum_chroma_host(const float *a_host, const float *b_host, int64_t n_pairs, int64_t n_samples, int n_fft,
                             double sample_rate, float *mag_db, float *mag_lin, float *chroma, float *peak_freq);

#ifdef __cplusplus
}
#endif
#endif /* AESIM_H */
