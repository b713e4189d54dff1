// aes_biquad_seq.cuh -- biquad cascades over a BATCH of clips: one thread per (clip, time segment) runs the
// reference's own per-sample recurrence (filter.py:8-40: y0 = b0*x0 + b1*x1 + b2*x2 - a1*y1 - a2*y2 in f64,
// every stage's output stored as f32), both channels and up to four stages interleaved for instruction-level
// parallelism.  No scan, no shuffles, no CTA barrier: 5 FP64 operations per sample and stage, one of them on the
// recurrence's critical path, against ~15 for the time-parallel formulations (aes_fast_kernel's biquad stage,
// aes_biquad_scan.cuh), which stay for what this one cannot fill the machine with -- few clips.
//
// Parallelism = clips x segments.  A stable biquad forgets: a segment that does not start the clip runs `warm`
// frames ahead of its first output from zero state and throws those outputs away; `warm` is the cascade's
// memory (the look-back depths of aes_biquad_build.h: |A^n| < 2^-44 per stage), so the state it arrives with
// is the sequential one to ~1e-13 -- below the f32 rounding of every stage's output.  In-place calls
// (x == y) are not segmented (a warm-up would read what the previous segment already overwrote).
//
// Memory: a thread walks its own stretch of its own clip, so a warp touches 32 distant places at once.  Every
// lane stages its stretch itself in chunks of AESQ_CH frames (256 bytes): cp.async.bulk global -> shared into a
// ring of AESQ_NBUF row buffers with one mbarrier per lane and buffer, computes in place (LDS.128 / STS.128 on
// rows padded to a 4-bank skew: conflict-free), and sends the chunk back with a bulk shared -> global copy.
// Whole 32-byte sectors both ways, nothing coalesced and nothing wasted.  Warps never meet.
#pragma once
#include <type_traits>
#include "aes_fast_kernel.cuh"      // cp.async.bulk / mbarrier helpers

#define AESQ_CH 32                              // frames per chunk (16: more warps fit, yet 114 against 167 Gsamples/s)
#define AESQ_NBUF 3                             // chunk buffers per lane (4: 6 warps per SM, 241 against 320 Gsamples/s; 2: 163)
#define AESQ_ROW (AESQ_CH * 2 + 4)              // floats per row: 256 bytes + 16 bytes of skew
#define AESQ_WARPS 2                            // warps per CTA (independent of each other)
#define AESQ_MAX_STAGES 4
#define AESQ_SMEM_BYTES (AESQ_WARPS * (AESQ_NBUF * 32 * AESQ_ROW * 4 + AESQ_NBUF * 32 * 8))

struct BqSeqArgs {
    const float *x;
    float *y;
    long long B, N;                             // clips, frames per clip (even)
    long long seg, warm;                        // frames per segment and warm-up frames, multiples of AESQ_CH
    int K;                                      // segments per clip (the last one takes the remainder)
    int n_stages;
    double bq[AESQ_MAX_STAGES][5];              // b0 b1 b2 a1 a2
    double init[AESQ_MAX_STAGES][8];            // [4*ch + {0,1,2,3}] = x1 x2 y1 y2 at the clip start (filter.py:16-19)
};

#ifndef AES_CPU_EMU
__device__ __forceinline__ void aes_bulk_s2g(void *gdst, const void *ssrc, unsigned bytes)
{
    const unsigned s = (unsigned)__cvta_generic_to_shared(ssrc);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(s), "r"(bytes) : "memory");
}
__device__ __forceinline__ void aes_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// at most N of this thread's bulk groups may still be READING their shared-memory source
template <int N> __device__ __forceinline__ void aes_bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N> __device__ __forceinline__ void aes_bulk_wait_all() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
#else
static inline void aes_bulk_s2g(void *gdst, const void *ssrc, unsigned bytes) { std::memcpy(gdst, ssrc, bytes); }
static inline void aes_bulk_commit() {}
template <int N> static inline void aes_bulk_wait_read() {}
template <int N> static inline void aes_bulk_wait_all() {}
#endif

template <int NS>
__device__ __forceinline__ void aes_biquad_seq_body(const BqSeqArgs &a)
{
    AES_DYN_SMEM(float, smem);
    constexpr int D = NS - 1;                   // steps an output pair lags its input pair (stage pipeline below)
    constexpr int PC = AESQ_CH / 2;             // pairs of frames per chunk
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *const rows = smem + (size_t)warp * (AESQ_NBUF * 32 * AESQ_ROW) + lane * AESQ_ROW;     // + buf * 32 * AESQ_ROW
    unsigned long long *const bars = reinterpret_cast<unsigned long long *>(smem + (size_t)AESQ_WARPS * AESQ_NBUF * 32 * AESQ_ROW)
                                     + (warp * 32 + lane) * AESQ_NBUF;
#pragma unroll
    for (int b = 0; b < AESQ_NBUF; ++b) aes_mbar_init(bars + b, 1);
    aes_mbar_init_fence();
    __syncwarp();
    const unsigned long long once = aes_policy_evict_first();
    const long long n_items = a.B * a.K, n_threads = (long long)gridDim.x * (AESQ_WARPS * 32);
    unsigned g = 0;                             // chunks this thread has put through its buffer ring; never reset

    // persistent threads: item = (clip, segment)
    for (long long idx = (long long)blockIdx.x * (AESQ_WARPS * 32) + threadIdx.x; idx < n_items; idx += n_threads) {
        const long long clip = idx / a.K;
        const int k = (int)(idx % a.K);
        const long long f_out = (long long)k * a.seg;                       // first frame this item writes
        const long long f_end = k == a.K - 1 ? a.N : f_out + a.seg;
        long long f0 = f_out - a.warm;                                      // first frame it reads
        const bool from_start = f0 <= 0;
        if (f0 < 0) f0 = 0;
        const int nchunks = (int)((f_end - f0 + AESQ_CH - 1) / AESQ_CH);
        const int p_out = (int)((f_out - f0) / 2), p_end = (int)((f_end - f0) / 2);   // output window, in pairs from f0
        const float *const xc = a.x + 2 * (clip * a.N + f0);
        float *const yc = a.y + 2 * (clip * a.N + f0);

        // DF-I history per stage and channel: the clip's own at frame 0, zero ahead of a warm-up
        double x1[NS][2], x2[NS][2], y1[NS][2], y2[NS][2];
        float mid[NS][2][2];                                                // input pair of stage s at the current step
#pragma unroll
        for (int s = 0; s < NS; ++s)
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
                x1[s][ch] = from_start ? a.init[s][4 * ch + 0] : 0.0;
                x2[s][ch] = from_start ? a.init[s][4 * ch + 1] : 0.0;
                y1[s][ch] = from_start ? a.init[s][4 * ch + 2] : 0.0;
                y2[s][ch] = from_start ? a.init[s][4 * ch + 3] : 0.0;
                mid[s][ch][0] = mid[s][ch][1] = 0.0f;
            }
        auto chunk_bytes = [&](int c) -> unsigned {
            const long long left = f_end - (f0 + (long long)c * AESQ_CH);
            return (unsigned)(left < AESQ_CH ? left : AESQ_CH) * 8u;
        };
        auto load = [&](int c) {                                            // chunk c of this item -> buffer (g0 + c) % NBUF
            if (c < nchunks) {
                const unsigned b = (g + (unsigned)c) % AESQ_NBUF;
                aes_fence_proxy_async_smem();                               // the row was last touched by ordinary accesses
                aes_mbar_expect(bars + b, chunk_bytes(c));
                aes_bulk_g2s(rows + b * 32 * AESQ_ROW, xc + 2 * (long long)c * AESQ_CH, chunk_bytes(c), bars + b, once);
#ifdef AES_CPU_EMU
                aes_mbar_complete_emu(bars + b);
#endif
            }
        };
        // the previous item's last copies out (the flush chunk sits in the buffer chunk 0 goes to) are through with their rows
        aes_bulk_wait_read<0>();
#pragma unroll
        for (int c = 0; c < AESQ_NBUF - 1; ++c) load(c);

        // Stage pipeline inside the thread: at step i stage s works on the pair of frames i - s, fed with what stage
        // s - 1 produced at step i - 1.  The NS x 2 (stage, channel) recurrences of a step are independent of each
        // other -- that is the instruction-level parallelism -- and the output pair lands D = NS - 1 positions late
        // in the row, which the copy out undoes.  One flush pseudo-chunk of D steps drains the pipeline.
        const int ntot = nchunks + (D > 0 ? 1 : 0);
        for (int c = 0; c < ntot; ++c) {
            const unsigned gc = g + (unsigned)c;
            float *row = rows + (gc % AESQ_NBUF) * 32 * AESQ_ROW;
            const bool flush = c >= nchunks;
            if (!flush) aes_mbar_wait(bars + gc % AESQ_NBUF, gc / AESQ_NBUF);
            const int nsteps = flush ? D : PC;
            // one step; GUARD (the first D steps of an item): stage s has no input yet at step j < s and must not
            // touch its history -- a carried state would decay over dummy input
            auto step = [&](auto guard, const int j) {
                constexpr bool GUARD = decltype(guard)::value;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (!flush) v = aes_lds_v4(row + 4 * j);
                mid[0][0][0] = v.x; mid[0][1][0] = v.y; mid[0][0][1] = v.z; mid[0][1][1] = v.w;
                float nxt[NS][2][2];
#pragma unroll
                for (int s = 0; s < NS; ++s) {
                    if (GUARD && j < s) {
#pragma unroll
                        for (int ch = 0; ch < 2; ++ch) nxt[s][ch][0] = nxt[s][ch][1] = 0.0f;
                        continue;
                    }
#pragma unroll
                    for (int ch = 0; ch < 2; ++ch)
#pragma unroll
                        for (int f = 0; f < 2; ++f) {
                            const double xin = (double)mid[s][ch][f];
                            // everything but the y1 term first: one dependent DFMA per sample on the recurrence
                            const double t = fma(a.bq[s][0], xin, fma(a.bq[s][1], x1[s][ch], fma(a.bq[s][2], x2[s][ch], -a.bq[s][4] * y2[s][ch])));
                            const double yy = fma(-a.bq[s][3], y1[s][ch], t);
                            x2[s][ch] = x1[s][ch]; x1[s][ch] = xin;
                            y2[s][ch] = y1[s][ch]; y1[s][ch] = yy;
                            nxt[s][ch][f] = (float)yy;                      // a block's output is float32 (filter.py:29)
                        }
                }
#pragma unroll
                for (int s = 1; s < NS; ++s)
#pragma unroll
                    for (int ch = 0; ch < 2; ++ch) { mid[s][ch][0] = nxt[s - 1][ch][0]; mid[s][ch][1] = nxt[s - 1][ch][1]; }
                *reinterpret_cast<float4 *>(row + 4 * j) = make_float4(nxt[NS - 1][0][0], nxt[NS - 1][1][0], nxt[NS - 1][0][1], nxt[NS - 1][1][1]);
            };
            int j = 0;
            if (D > 0 && c == 0)
                for (; j < D && j < nsteps; ++j) step(std::true_type{}, j);
#pragma unroll 2                // (four: 322 against 332 Gsamples/s on 8192 clips, r2as)
            for (; j < nsteps; ++j) step(std::false_type{}, j);
            // row positions [0, nsteps) now hold the output pairs [c*PC - D, c*PC - D + nsteps) of this item
            const int q0 = c * PC - D;
            const int lo = q0 > p_out ? q0 : p_out, hi = q0 + nsteps < p_end ? q0 + nsteps : p_end;
            if (hi > lo) {
                aes_fence_proxy_async_smem();                               // ordinary stores -> the bulk copy's reads
                aes_bulk_s2g(yc + 4 * (long long)lo, row + 4 * (lo - q0), (unsigned)(hi - lo) * 16u);
            }
            aes_bulk_commit();
            // chunk c + NBUF - 1 goes into the buffer of chunk c - 1: its copy out must be through with the row
            aes_bulk_wait_read<1>();
            load(c + AESQ_NBUF - 1);
        }
        g += (unsigned)nchunks;                 // (the flush pseudo-chunk took no copy in: its buffer's mbarrier has not moved)
    }
    aes_bulk_wait_all<0>();
}

#ifndef AES_CPU_EMU
// Launch bounds with the 4 CTAs per SM the staging rows allow: 256 threads per SM, so ptxas may take up to 255
// registers.  Without the second argument it stopped at 128 and spilled 16 bytes into the step loop (NS = 4);
// with it: 138 registers, no spill, BASELINE's cascade on 8192 clips 319.8 -> 332.3 Gsamples/s, one filter 517 -> 563.
template <int NS>
__global__ void __launch_bounds__(AESQ_WARPS * 32, 4) aes_biquad_seq_kernel(const __grid_constant__ BqSeqArgs a)
{
    aes_biquad_seq_body<NS>(a);
}
#endif
