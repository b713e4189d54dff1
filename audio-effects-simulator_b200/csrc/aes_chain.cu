// aes_chain.cu -- host side of the fused effect chain: plan objects, kernel launch,
// and the host-buffer entry that pipelines H2D / kernel / D2H over internal streams.
#include <algorithm>
#include <math.h>
#include <new>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>

#include "aes_common.h"
#include "aes_plan_build.h"
#include "aes_chain_kernel.cuh"
#include "aes_fast_build.h"
#include "aes_rv_build.h"
#include "aes_biquad_build.h"
#include "aes_biquad_seq.cuh"

#define AES_HOST_SLOTS 4

struct HostSlot {
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
    void *d_in = nullptr, *d_out = nullptr;     // device staging for one sub-batch
    void *h_in = nullptr, *h_out = nullptr;     // pinned staging (only for pageable user buffers)
    void *scratch = nullptr;                    // global-memory delay lines of the CTAs of one launch
    size_t in_cap = 0, out_cap = 0, hin_cap = 0, hout_cap = 0, scratch_cap = 0;
    // pending copy-out of a finished sub-batch from pinned staging to the user's buffer
    void *pend_dst = nullptr;
    size_t pend_bytes = 0;
};

// The host pipeline's slots (streams, events, device and pinned staging, line scratch) belong to
// the calling host thread, not to a plan: the file route builds a fresh plan per request
// (engine.py:86-99 re-creates its chain every time), and allocating ~200 MB of staging per request
// cost 40-200 ms of cudaMalloc / cudaHostAlloc / cudaFree against 3 ms of copies and kernel.
// Buffers only grow; aes_release_host_cache() frees them.
struct SlotCache {
    HostSlot slot[AES_HOST_SLOTS];
    bool ready = false;
    int device = -1;
};
static thread_local SlotCache t_cache;

// ---- shape-specialised kernels (aes_fast_kernel.cuh): launch table -----------------------
typedef void (*fast_kernel_t)(const FastArgs);
struct FastShape { int c[AESF_MAX_STAGES]; int topo; fast_kernel_t fn; };
#define X(c0, c1, c2, c3, mp) { { c0, c1, c2, c3 }, mp, aes_fast_kernel<4, c0, c1, c2, c3, mp> },
#define X3(c0, c1, c2, c3, mp) { { c0, c1, c2, c3 }, mp, aes_fast_kernel<4, c0, c1, c2, c3, mp, 3> },
static const FastShape g_fast_shapes[] = { AESF_SHAPES(X) AESF_SHAPES_3CTA(X3) };
#undef X
#undef X3

// ---- pipelined reverb-chain kernels (aes_rv_kernel.cuh): launch table ------------------------
struct RvShape { int topo, pre, pm; fast_kernel_t fn; };
#define X(topo, pre, pm) { topo, pre, pm, aes_rv_kernel<topo, pre, pm> },
static const RvShape g_rv_shapes[] = { AESRV_SHAPES(X) };
#undef X

struct aes_chain_plan {
    DevPlan host;
    // time-parallel biquad cascade (aes_biquad_scan.cuh) for few long clips
    bool bq_ok = false;
    BqArgs bq;
    cudaEvent_t bq_done = nullptr;              // last scan launch: the scratch below is shared, so scan launches of one
                                                // plan are chained by this event even when they sit on different streams
    cudaStream_t bq_stream = nullptr;           // stream of the last scan launch
    size_t bq_scan_recs = 0;                    // record count the scan scratch is laid out (and zeroed) for
    unsigned bq_ticket_base = 0;                // tickets handed out by earlier launches (the counter is never reset)
    unsigned long long bq_epoch = 0;            // tag of the last launch's look-back records
    double *d_bq_tab = nullptr;                 // lane_pw [8][32][4] | tile_pw [8][256][4]
    void *d_bq_scan = nullptr;                  // agg | inc | flag | ticket, grown on demand
    size_t bq_scan_cap = 0;
    // sequential recurrence per (clip, segment) for batches of biquad cascades (aes_biquad_seq.cuh)
    bool bqs_ok = false;
    BqSeqArgs bqs;
    long long bqs_warm = -1;                    // frames a segment runs ahead of its first output; -1: the cascade
                                                // remembers too long, clips are not segmented
    FastArgs fast;                              // flattened descriptors when a specialised kernel fits
    fast_kernel_t fast_fn = nullptr;
    bool rv = false;                            // fast_fn is an aes_rv_kernel instantiation
    float *d_lane_tab = nullptr;
    DevPlan *dev = nullptr;
    int fs = 0, device = 0, sm_count = 0, ctas_per_sm = 0, grid_max = 0;
    size_t smem_bytes = 0;
    void *scratch = nullptr;                    // lines of the CTAs of an aes_chain_run launch (grown on demand)
    size_t scratch_cap = 0;
    double *d_state = nullptr;                  // final carried scalars of a single-clip host call
    double h_state[16 * AES_MAX_STAGES];
    bool state_valid = false;
    // A chain no kernel was specialised for, cut at stage boundaries into runs that are (the signal is
    // f32 between stages in every kernel, so the cut changes no bit): launched back to back, in place
    // on the output buffer.  Empty when the whole chain has its own kernel or no cut helps.
    std::vector<aes_chain_plan *> seg;
    std::vector<int> seg_first;                 // first stage of each segment
    // A chain WITH a kernel of its own that starts with biquads ahead of a pipelined reverb chain (Guitar Filter):
    // on a batch large enough for the sequential biquad kernel, [biquads] + [the rest] in two launches is faster
    // than the fused kernel (12.5 against 13.9 ms on 2368 clips x 10 s) -- the f64 scan stage costs the fused kernel
    // more than a second trip through HBM.  Chosen per call in launch_chain.
    std::vector<aes_chain_plan *> alt;
    std::vector<int> alt_first;
};

static int plan_create(const aes_stage_desc *stages, int n_stages, int sample_rate, aes_chain_plan **out, bool allow_split);

template <int K>
static int configure_kernel(aes_chain_plan *pl)
{
    // (the attribute belongs to the kernel, not to the plan: always the limit, so that a later, smaller plan
    //  does not take a live plan's launch size away)
    AES_CUDA(cudaFuncSetAttribute(aes_chain_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, AES_SMEM_LIMIT));
    int occ = 0;
    AES_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, aes_chain_kernel<K>, AES_NT, pl->smem_bytes));
    if (occ < 1) { aes_set_error("chain kernel does not fit on an SM"); return AES_ERR_UNSUPPORTED; }
    pl->ctas_per_sm = occ;
    return 0;
}

// Segments per clip for the sequential batch kernel (persistent threads, one (clip, segment) item at a time): the
// K that wastes least -- whole rounds of the resident threads, warm-up short against the segment.  0: the kernel
// does not serve this call.
static int bqseq_segments(const aes_chain_plan *pl, long long decide_B, long long N, bool in_place)
{
    if (!pl->bqs_ok || (N & 1) || getenv("AES_NO_BQSEQ")) return 0;
    const long long resident = (long long)pl->sm_count * (227 * 1024 / AESQ_SMEM_BYTES) * AESQ_WARPS * 32;
    const long long least = (long long)pl->sm_count * 2 * 32;
    long long kmax = 1;
    if (!in_place && pl->bqs_warm > 0) kmax = std::max<long long>(1, std::min<long long>(N / (2 * pl->bqs_warm), 64));
    long long best = 1;
    double best_eff = 0.0;
    for (long long K = 1; K <= kmax; ++K) {
        const double rounds = (double)(decide_B * K) / (double)resident;
        const double seg = (double)N / (double)K;
        const double eff = rounds / ceil(rounds) * (K > 1 ? seg / (seg + (double)pl->bqs_warm) : 1.0);
        if (eff > best_eff + 1e-9) { best_eff = eff; best = K; }
    }
    return decide_B * best >= least ? (int)best : 0;
}

template <int NS> static void bqseq_launch(const aes_chain_plan *pl, const BqSeqArgs &q, cudaStream_t st)
{
    static bool once = [] { return cudaFuncSetAttribute(aes_biquad_seq_kernel<NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, AESQ_SMEM_BYTES) == cudaSuccess; }();
    (void)once;
    const long long need = (q.B * q.K + AESQ_WARPS * 32 - 1) / (AESQ_WARPS * 32);
    const long long resident = (long long)pl->sm_count * (227 * 1024 / AESQ_SMEM_BYTES);
    aes_biquad_seq_kernel<NS><<<(unsigned)std::min(need, resident), AESQ_WARPS * 32, AESQ_SMEM_BYTES, st>>>(q);
}

// `decide_B`: the batch size kernel choices are made on.  The host pipeline passes the size of the WHOLE batch,
// not of the sub-batch it launches -- the time-parallel scan, the sequential batch kernel and the tile kernels
// agree to an ulp, not bit for bit, and a clip's result must not depend on where in a batch it sits.
static int launch_chain(aes_chain_plan *pl, const void *x, int in_fmt, void *y, int out_fmt,
                        long long B, long long N, float *scratch, cudaStream_t st, double *state_out = nullptr,
                        long long decide_B = -1)
{
    if (B <= 0 || N <= 0) return 0;
    if (decide_B < 0) decide_B = B;
    const bool allow_scan = decide_B < pl->grid_max;
    const bool use_alt = !pl->alt.empty() && state_out == nullptr && in_fmt == AES_FMT_F32_STEREO &&
                         bqseq_segments(pl->alt[0], decide_B, N, x == y) > 0;
    if (!pl->seg.empty() || use_alt) {
        // segment k reads what segment k-1 wrote: f32 stereo, in the output buffer itself when that is the
        // caller's format (every kernel reads a tile before it writes it), else in a stream-ordered temporary
        const std::vector<aes_chain_plan *> &segs = use_alt ? pl->alt : pl->seg;
        const std::vector<int> &firsts = use_alt ? pl->alt_first : pl->seg_first;
        float *mid = (float *)y;
        if (out_fmt != AES_FMT_F32_STEREO) AES_CUDA(cudaMallocAsync((void **)&mid, (size_t)B * N * 2 * sizeof(float), st));
        int rc = 0;
        for (size_t k = 0; k < segs.size() && !rc; ++k) {
            const bool first = k == 0, last = k + 1 == segs.size();
            rc = launch_chain(segs[k], first ? x : mid, first ? in_fmt : AES_FMT_F32_STEREO, last ? y : mid,
                              last ? out_fmt : AES_FMT_F32_STEREO, B, N, scratch, st,
                              state_out ? state_out + 16 * firsts[k] : nullptr, decide_B);
        }
        if (mid != (float *)y) cudaFreeAsync(mid, st);
        return rc;
    }
    if (state_out == nullptr && in_fmt == AES_FMT_F32_STEREO && out_fmt == AES_FMT_F32_STEREO) {
        if (const int K = bqseq_segments(pl, decide_B, N, x == y)) {
            // a batch of biquad cascades: the plain recurrence, one thread per (clip, segment)
            BqSeqArgs q = pl->bqs;
            q.x = (const float *)x; q.y = (float *)y; q.B = B; q.N = N; q.K = K;
            q.warm = K > 1 ? pl->bqs_warm : 0;
            q.seg = (N / K + AESQ_CH - 1) / AESQ_CH * AESQ_CH;
            switch (q.n_stages) {
            case 1: bqseq_launch<1>(pl, q, st); break;
            case 2: bqseq_launch<2>(pl, q, st); break;
            case 3: bqseq_launch<3>(pl, q, st); break;
            default: bqseq_launch<4>(pl, q, st); break;
            }
            aes_count_launch();
            AES_CUDA(cudaGetLastError());
            return 0;
        }
    }
    if (pl->bq_ok && allow_scan && (B < pl->grid_max || getenv("AES_FORCE_SCAN")) && in_fmt == AES_FMT_F32_STEREO && out_fmt == AES_FMT_F32_STEREO &&
        !getenv("AES_NO_SCAN")) {
        // fewer clips than resident CTAs: go parallel in time (one CTA per 1024-frame tile)
        const long long nt = (N + AESB_T - 1) / AESB_T;
        if (!pl->bq_done) AES_CUDA(cudaEventCreateWithFlags(&pl->bq_done, cudaEventDisableTiming));
        else if (st != pl->bq_stream) AES_CUDA(cudaStreamWaitEvent(st, pl->bq_done, 0));
        pl->bq_stream = st;
        const size_t recs = (size_t)B * pl->bq.n_stages * nt;
        const size_t need = recs * (sizeof(BqRec) * 4 + 8 * sizeof(double) + sizeof(int)) + 64;
        if (pl->bq_scan_cap < need) {
            if (pl->d_bq_scan) cudaFree(pl->d_bq_scan);
            pl->d_bq_scan = nullptr; pl->bq_scan_cap = 0;
            AES_CUDA(cudaMalloc(&pl->d_bq_scan, need));
            pl->bq_scan_cap = need;
            pl->bq_scan_recs = 0;
        }
        if (pl->bq_scan_recs != recs) {             // new layout: epoch records and the ticket start at 0, once
            AES_CUDA(cudaMemsetAsync(pl->d_bq_scan, 0, need, st));
            pl->bq_scan_recs = recs;
            pl->bq_ticket_base = 0;
        }
        BqArgs a = pl->bq;
        a.x = (const float *)x; a.y = (float *)y; a.N = N; a.n_tiles = nt; a.B = B; a.dbg_skip = 0;
        if (getenv("AES_SCAN_CHAINED"))             // tests: the chained look-back, which otherwise only slowly forgetting filters take
            for (int s = 0; s < a.n_stages; ++s) a.st[s].lb_k = 0;
        // the ticket lives at the FRONT so that it keeps its place when recs changes under the same capacity
        a.ticket = (unsigned *)pl->d_bq_scan;
        a.ticket_base = pl->bq_ticket_base;
        pl->bq_ticket_base += (unsigned)(B * nt);
        a.epoch = ++pl->bq_epoch;
        a.rec16 = (BqRec *)((char *)pl->d_bq_scan + 64);
        a.agg = (double *)(a.rec16 + recs * 4);
        a.inc = a.agg + recs * 4;
        a.flag = (int *)(a.inc + recs * 4);
        a.lane_pw = pl->d_bq_tab;
        a.tile_pw = pl->d_bq_tab + (size_t)AESB_MAX_STAGES * 128;
        a.final_state = state_out;
        bool chained = false;
        for (int s = 0; s < a.n_stages; ++s) chained |= a.st[s].lb_k == 0;
        if (chained) AES_CUDA(cudaMemsetAsync(a.flag, 0, recs * sizeof(int), st));
        aes_biquad_scan_kernel<<<(unsigned)(B * nt), AESB_NT, AESB_SMEM_DOUBLES * sizeof(double), st>>>(a);
        AES_CUDA(cudaEventRecord(pl->bq_done, st));
        aes_count_launch();
        AES_CUDA(cudaGetLastError());
        return 0;
    }
    ChainArgs a{ pl->dev, x, y, B, N, scratch, in_fmt, out_fmt, state_out };
    const unsigned grid = (unsigned)std::min<long long>(B, pl->grid_max);
    if (pl->fast_fn) {
        // the specialised kernels index frames inside a clip with 32 bits (12 h of audio at 48 kHz)
        AES_REQUIRE(N < (1LL << 31) - 8192, "clips of 2^31 frames or more are not supported by this chain's kernel");
        FastArgs fa = pl->fast;
        fa.x = x; fa.y = y; fa.B = B; fa.N = N; fa.scratch = scratch; fa.lane_tab = pl->d_lane_tab;
        fa.state_out = state_out; fa.in_fmt = in_fmt; fa.out_fmt = out_fmt;
        pl->fast_fn<<<grid, pl->rv ? AESRV_NT : AES_NT, pl->smem_bytes, st>>>(fa);
        aes_count_launch();
        AES_CUDA(cudaGetLastError());
        return 0;
    }
    switch (pl->host.FR) {
    case 4: aes_chain_kernel<4><<<grid, AES_NT, pl->smem_bytes, st>>>(a); break;
    case 2: aes_chain_kernel<2><<<grid, AES_NT, pl->smem_bytes, st>>>(a); break;
    default: aes_chain_kernel<1><<<grid, AES_NT, pl->smem_bytes, st>>>(a); break;
    }
    aes_count_launch();
    AES_CUDA(cudaGetLastError());
    return 0;
}

static size_t in_frame_bytes(int fmt)
{
    return fmt == AES_FMT_F32_STEREO ? 8 : 4;       // f32 mono and i16 stereo are both 4 bytes/frame
}
static size_t out_frame_bytes(int fmt) { return fmt == AES_FMT_I16_STEREO ? 4 : 8; }

static int check_formats(int in_fmt, int out_fmt)
{
    AES_REQUIRE(in_fmt == AES_FMT_F32_STEREO || in_fmt == AES_FMT_F32_MONO || in_fmt == AES_FMT_I16_STEREO_DOWNMIX,
                "unsupported input format %d", in_fmt);
    AES_REQUIRE(out_fmt == AES_FMT_F32_STEREO || out_fmt == AES_FMT_I16_STEREO, "unsupported output format %d", out_fmt);
    return 0;
}

AES_EXPORT int aes_chain_plan_create(const aes_stage_desc *stages, int n_stages, int sample_rate,
                                     aes_chain_plan **out)
{
    return plan_create(stages, n_stages, sample_rate, out, true);
}

static int plan_create(const aes_stage_desc *stages, int n_stages, int sample_rate, aes_chain_plan **out, bool allow_split)
{
    AES_REQUIRE(out != nullptr, "plan out-pointer is NULL");
    AES_REQUIRE(n_stages == 0 || stages != nullptr, "stages is NULL");
    aes_chain_plan *pl = new (std::nothrow) aes_chain_plan();
    if (!pl) { aes_set_error("out of host memory"); return AES_ERR_NOMEM; }
    char err[512];
    int rc = aes_build_devplan(stages, n_stages, sample_rate, &pl->host, err, sizeof err);
    if (rc) { aes_set_error("%s", err); delete pl; return rc; }
    pl->fs = sample_rate;
    pl->smem_bytes = aes_plan_smem_bytes(pl->host);
    rc = [&]() -> int {
        AES_CUDA(cudaGetDevice(&pl->device));
        AES_CUDA(cudaDeviceGetAttribute(&pl->sm_count, cudaDevAttrMultiProcessorCount, pl->device));
        // a specialised kernel, when the chain's shape is one of the pre-instantiated ones
        std::vector<float> lane_tab_v((size_t)AESF_MAX_STAGES * 32 * FAST_LANE_STRIDE);   // (not static: plans may be built concurrently)
        float *lane_tab = lane_tab_v.data();
        const size_t lane_tab_bytes = lane_tab_v.size() * sizeof(float);
        int codes[AESF_MAX_STAGES];
        if (!getenv("AES_NO_FAST") && aes_fast_build(pl->host, &pl->fast, codes, lane_tab)) {
            const size_t fast_smem = aes_fast_smem_bytes(pl->host);
            const int topo = aes_fast_topo(pl->host);
            // chains ending in the default reverb: the software-pipelined kernel
            int rv_pre = 0, rv_pm = 0;
            if (!getenv("AES_NO_RV") && aes_rv_shape(pl->fast, codes, topo, &rv_pre, &rv_pm)) {
                const size_t rv_smem = aes_rv_smem_bytes(pl->host.smem_floats, rv_pre);
                for (const RvShape &sh : g_rv_shapes) {
                    if (sh.topo != topo || sh.pre != rv_pre || sh.pm != rv_pm || rv_smem > AES_SMEM_LIMIT) continue;
                    AES_CUDA(cudaFuncSetAttribute(sh.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, AES_SMEM_LIMIT));
                    int occ = 0;
                    AES_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, sh.fn, AESRV_NT, rv_smem));
                    if (occ < 1) break;
                    pl->smem_bytes = rv_smem;
                    AES_CUDA(cudaMalloc(&pl->d_lane_tab, lane_tab_bytes));
                    AES_CUDA(cudaMemcpy(pl->d_lane_tab, lane_tab, lane_tab_bytes, cudaMemcpyHostToDevice));
                    pl->fast_fn = sh.fn;
                    pl->rv = true;
                    pl->ctas_per_sm = occ;
                    break;
                }
            }
            if (!pl->fast_fn)
            for (const FastShape &sh : g_fast_shapes) {      // compile-time topologies come first
                if (memcmp(sh.c, codes, sizeof codes) != 0 || fast_smem > AES_SMEM_LIMIT) continue;
                if (sh.topo != AESF_TOPO_NONE && sh.topo != topo) continue;
                AES_CUDA(cudaFuncSetAttribute(sh.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, AES_SMEM_LIMIT));
                int occ = 0;
                AES_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, sh.fn, AES_NT, fast_smem));
                if (occ < 1) break;
                pl->smem_bytes = fast_smem;
                AES_CUDA(cudaMalloc(&pl->d_lane_tab, lane_tab_bytes));
                AES_CUDA(cudaMemcpy(pl->d_lane_tab, lane_tab, lane_tab_bytes, cudaMemcpyHostToDevice));
                pl->fast_fn = sh.fn;
                pl->ctas_per_sm = occ;
                break;
            }
        }
        int r2 = 0;
        if (!pl->fast_fn)
        switch (pl->host.FR) {
        case 4: r2 = configure_kernel<4>(pl); break;
        case 2: r2 = configure_kernel<2>(pl); break;
        default: r2 = configure_kernel<1>(pl); break;
        }
        if (r2) return r2;
        pl->grid_max = pl->sm_count * pl->ctas_per_sm;
        AES_CUDA(cudaMalloc(&pl->dev, sizeof(DevPlan)));
        AES_CUDA(cudaMemcpy(pl->dev, &pl->host, sizeof(DevPlan), cudaMemcpyHostToDevice));
        // all-biquad chains also get the time-parallel scan kernel
        bool all_bq = n_stages >= 1 && n_stages <= AESB_MAX_STAGES;
        for (int s2 = 0; s2 < n_stages && all_bq; ++s2) all_bq = stages[s2].kind == AES_STAGE_BIQUAD;
        if (all_bq) {
            std::vector<double> tab_v((size_t)AESB_MAX_STAGES * 128 + (size_t)AESB_MAX_STAGES * AESB_LBW * 4);
            double *tab = tab_v.data();
            const size_t tab_bytes = tab_v.size() * sizeof(double);
            double co[5 * AESB_MAX_STAGES], dfi[AESB_MAX_STAGES * 8];
            for (int s2 = 0; s2 < n_stages; ++s2) {
                for (int i = 0; i < 5; ++i) co[5 * s2 + i] = stages[s2].p[i];
                for (int i = 0; i < 8; ++i) dfi[8 * s2 + i] = stages[s2].p[8 + i];
            }
            aes_biquad_build(n_stages, co, dfi, &pl->bq, tab, tab + AESB_MAX_STAGES * 128);
            AES_CUDA(cudaMalloc(&pl->d_bq_tab, tab_bytes));
            AES_CUDA(cudaMemcpy(pl->d_bq_tab, tab, tab_bytes, cudaMemcpyHostToDevice));
            pl->bq_ok = true;
            if (n_stages <= AESQ_MAX_STAGES) {
                memset(&pl->bqs, 0, sizeof pl->bqs);
                pl->bqs.n_stages = n_stages;
                pl->bqs_warm = 0;
                for (int s2 = 0; s2 < n_stages; ++s2) {
                    for (int i = 0; i < 5; ++i) pl->bqs.bq[s2][i] = co[5 * s2 + i];
                    for (int i = 0; i < 8; ++i) pl->bqs.init[s2][i] = dfi[8 * s2 + i];
                    // a stage's memory: the look-back depth of the scan, in tiles (0: longer than its window)
                    if (pl->bq.st[s2].lb_k == 0) pl->bqs_warm = -1;
                    else if (pl->bqs_warm >= 0) pl->bqs_warm += (long long)pl->bq.st[s2].lb_k * AESB_T;
                }
                pl->bqs_ok = true;
            }
        }
        AES_CUDA(cudaMalloc(&pl->d_state, sizeof pl->h_state));
        AES_CUDA(cudaMemset(pl->d_state, 0, sizeof pl->h_state));
        // no kernel for the whole chain: cut it into the longest runs that have one
        if (allow_split && !pl->fast_fn && !pl->bq_ok && n_stages > 1 && !getenv("AES_NO_SPLIT")) {
            int i = 0, n_fast = 0;
            while (i < n_stages) {
                aes_chain_plan *best = nullptr;
                int blen = 0;
                for (int len = std::min<int>(AESF_MAX_STAGES, n_stages - i); len >= 1 && !best; --len) {
                    aes_chain_plan *sp = nullptr;
                    if (plan_create(stages + i, len, sample_rate, &sp, false) != 0) continue;
                    if (sp->fast_fn || len == 1) { best = sp; blen = len; }
                    else aes_chain_plan_destroy(sp);
                }
                if (!best) break;
                n_fast += best->fast_fn != nullptr;
                pl->seg.push_back(best); pl->seg_first.push_back(i);
                i += blen;
            }
            // each launch is another trip through HBM, yet four of them still run a 5-stage chain twice as fast
            // as the interpreter (profiles/README.md, r2q)
            const char *env = getenv("AES_SPLIT_MAX");
            const size_t max_seg = env ? (size_t)atoi(env) : 8;
            if (i < n_stages || n_fast == 0 || pl->seg.size() > max_seg) {
                for (aes_chain_plan *sp : pl->seg) aes_chain_plan_destroy(sp);
                pl->seg.clear(); pl->seg_first.clear();
            } else {
                for (aes_chain_plan *sp : pl->seg) {    // one scratch and one grid bound serve every segment
                    pl->host.scratch_floats = std::max(pl->host.scratch_floats, sp->host.scratch_floats);
                    pl->grid_max = std::max(pl->grid_max, sp->grid_max);
                }
            }
        }
        // biquads ahead of a pipelined reverb chain: keep the two-launch alternative beside the fused kernel
        if (allow_split && pl->fast_fn && !getenv("AES_NO_SPLIT")) {
            int nb = 0;
            while (nb < n_stages && nb < AESQ_MAX_STAGES && stages[nb].kind == AES_STAGE_BIQUAD) ++nb;
            if (nb >= 1 && nb < n_stages) {
                aes_chain_plan *pa = nullptr, *pb = nullptr;
                if (plan_create(stages, nb, sample_rate, &pa, false) == 0 &&
                    plan_create(stages + nb, n_stages - nb, sample_rate, &pb, false) == 0 && pa->bqs_ok && pb->rv) {
                    pl->alt = { pa, pb };
                    pl->alt_first = { 0, nb };
                    for (aes_chain_plan *sp : pl->alt) {
                        pl->host.scratch_floats = std::max(pl->host.scratch_floats, sp->host.scratch_floats);
                        pl->grid_max = std::max(pl->grid_max, sp->grid_max);
                    }
                } else {
                    if (pa) aes_chain_plan_destroy(pa);
                    if (pb) aes_chain_plan_destroy(pb);
                }
            }
        }
        return 0;
    }();
    if (rc) { aes_chain_plan_destroy(pl); return rc; }
    *out = pl;
    return 0;
}

AES_EXPORT int aes_chain_plan_destroy(aes_chain_plan *pl)
{
    if (!pl) return 0;
    for (aes_chain_plan *sp : pl->seg) aes_chain_plan_destroy(sp);
    for (aes_chain_plan *sp : pl->alt) aes_chain_plan_destroy(sp);
    if (pl->dev) cudaFree(pl->dev);
    if (pl->scratch) cudaFree(pl->scratch);
    if (pl->d_state) cudaFree(pl->d_state);
    if (pl->d_lane_tab) cudaFree(pl->d_lane_tab);
    if (pl->d_bq_tab) cudaFree(pl->d_bq_tab);
    if (pl->bq_done) cudaEventDestroy(pl->bq_done);
    if (pl->d_bq_scan) cudaFree(pl->d_bq_scan);
    delete pl;
    return 0;
}

AES_EXPORT int aes_chain_plan_info(const aes_chain_plan *pl, int *tile_frames, int *smem_bytes,
                                   int *ctas_per_sm, int64_t *scratch_bytes_per_cta)
{
    AES_REQUIRE(pl != nullptr, "plan is NULL");
    if (tile_frames) *tile_frames = pl->host.T;
    if (smem_bytes) *smem_bytes = (int)pl->smem_bytes;
    if (ctas_per_sm) *ctas_per_sm = pl->ctas_per_sm;
    if (scratch_bytes_per_cta) *scratch_bytes_per_cta = (int64_t)pl->host.scratch_floats * 4;
    return 0;
}

AES_EXPORT const char *aes_chain_plan_kernel_name(const aes_chain_plan *pl)
{
    if (!pl) return "";
    if (!pl->seg.empty()) {
        static thread_local std::string name;
        name = "split:";
        for (size_t k = 0; k < pl->seg.size(); ++k) {
            name += k ? " | " : " ";
            name += std::to_string(pl->seg[k]->host.n_stages) + "x " + aes_chain_plan_kernel_name(pl->seg[k]);
        }
        return name.c_str();
    }
    if (pl->rv) return "aes_rv_kernel<pipelined reverb chain>";
    return pl->fast_fn ? "aes_fast_kernel<FR=4, shape-specialised>" : "aes_chain_kernel<generic interpreter>";
}

static int grow(void **p, size_t *cap, size_t need, bool host)
{
    if (*cap >= need) return 0;
    if (*p) { if (host) cudaFreeHost(*p); else cudaFree(*p); *p = nullptr; *cap = 0; }
    if (host) AES_CUDA(cudaMallocHost(p, need)); else AES_CUDA(cudaMalloc(p, need));
    *cap = need;
    return 0;
}

static void release_cache(SlotCache &c)
{
    for (HostSlot &s : c.slot) {
        if (s.stream) cudaStreamSynchronize(s.stream);
        if (s.d_in) cudaFree(s.d_in);
        if (s.d_out) cudaFree(s.d_out);
        if (s.h_in) cudaFreeHost(s.h_in);
        if (s.h_out) cudaFreeHost(s.h_out);
        if (s.scratch) cudaFree(s.scratch);
        if (s.done) cudaEventDestroy(s.done);
        if (s.stream) cudaStreamDestroy(s.stream);
        s = HostSlot();
    }
    c.ready = false;
    c.device = -1;
}

// the calling thread's slots, created on first use and re-created when it switched device
static int ensure_slots(SlotCache &c)
{
    int dev = 0;
    AES_CUDA(cudaGetDevice(&dev));
    if (c.ready && c.device == dev) return 0;
    if (c.ready) {                                  // buffers of another device: release them there
        cudaSetDevice(c.device);
        release_cache(c);
        cudaSetDevice(dev);
    }
    for (HostSlot &s : c.slot) {
        AES_CUDA(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
        AES_CUDA(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
    }
    c.ready = true;
    c.device = dev;
    return 0;
}

// Frees the calling thread's cached pipeline buffers (device and pinned staging, line scratch,
// streams).  Nothing else frees them: a host thread that made host-buffer calls should call this
// before it ends (no destructor runs CUDA calls at thread or process exit on purpose).
AES_EXPORT int aes_release_host_cache(void)
{
    if (t_cache.ready) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaSetDevice(t_cache.device);
        release_cache(t_cache);
        cudaSetDevice(dev);
    }
    return 0;
}

AES_EXPORT int aes_chain_run(aes_chain_plan *pl, const void *x, int in_fmt, void *y, int out_fmt,
                             int64_t n_clips, int64_t n_frames, void *stream)
{
    AES_REQUIRE(pl != nullptr, "plan is NULL");
    AES_REQUIRE(n_clips >= 0 && n_frames >= 0, "negative size");
    if (n_clips == 0 || n_frames == 0) return 0;
    AES_REQUIRE(x != nullptr && y != nullptr, "NULL device buffer");
    AES_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 15) == 0, "buffers must be 16-byte aligned");
    int rc = check_formats(in_fmt, out_fmt);
    if (rc) return rc;
    // line scratch for the CTAs of this launch (kept by the plan: launches of one plan on one stream
    // are ordered; growing it waits for the device, which only happens on the first, largest call)
    const size_t need = (size_t)std::min<int64_t>(n_clips, pl->grid_max) * pl->host.scratch_floats * sizeof(float);
    if (pl->scratch_cap < need) {
        AES_CUDA(cudaDeviceSynchronize());
        if ((rc = grow(&pl->scratch, &pl->scratch_cap, need, false))) return rc;
    }
    return launch_chain(pl, x, in_fmt, y, out_fmt, n_clips, n_frames, (float *)pl->scratch, (cudaStream_t)stream);
}

static bool is_pinned(const void *p)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost;
}

static int drain_slot(HostSlot &s)
{
    AES_CUDA(cudaEventSynchronize(s.done));
    if (s.pend_dst) { memcpy(s.pend_dst, s.h_out, s.pend_bytes); s.pend_dst = nullptr; }
    return 0;
}

// Host buffers in, host buffers out.  The batch is cut into sub-batches of whole clips
// (about one full grid of CTAs each); sub-batch k runs on stream k mod 3 as
// H2D -> chain kernel -> D2H, so both copy engines and the SMs overlap.
AES_EXPORT int aes_chain_process_host(aes_chain_plan *pl, const void *x_host, int in_fmt, void *y_host,
                                      int out_fmt, int64_t n_clips, int64_t n_frames)
{
    AES_REQUIRE(pl != nullptr, "plan is NULL");
    AES_REQUIRE(n_clips >= 0 && n_frames >= 0, "negative size");
    if (n_clips == 0 || n_frames == 0) return 0;
    AES_REQUIRE(x_host != nullptr && y_host != nullptr, "NULL host buffer");
    int rc = check_formats(in_fmt, out_fmt);
    if (rc) return rc;
    SlotCache &cache = t_cache;
    if ((rc = ensure_slots(cache))) return rc;

    const size_t in_clip = (size_t)n_frames * in_frame_bytes(in_fmt);
    const size_t out_clip = (size_t)n_frames * out_frame_bytes(out_fmt);
    // sub-batch: one grid's worth of clips, capped at ~1 GiB of input per slot
    int64_t per = std::max<int64_t>(1, pl->grid_max);
    const size_t cap_bytes = (size_t)1 << 30;
    if ((size_t)per * in_clip > cap_bytes) per = std::max<int64_t>(1, (int64_t)(cap_bytes / in_clip));
    per = std::min<int64_t>(per, n_clips);
    // The PCIe copies dominate (a sub-batch's kernel is ~10x shorter than either of its copies), so
    // the pipeline's fill/drain bubble -- one H2D at the start, one D2H at the end that overlap
    // nothing -- decides the end-to-end rate: cut the batch into ~AES_HOST_CHUNKS sub-batches, but
    // keep every copy at 32 MiB or more.
    {
        int64_t chunks = 24;
        if (const char *e = getenv("AES_HOST_CHUNKS")) chunks = std::max<int64_t>(1, atoll(e));
        const int64_t min_clips = std::max<int64_t>(1, (int64_t)((((size_t)32 << 20) + in_clip - 1) / in_clip));
        const int64_t fine = std::max<int64_t>((n_clips + chunks - 1) / chunks, min_clips);
        per = std::min<int64_t>(per, fine);
    }
    const bool pin_in = is_pinned(x_host), pin_out = is_pinned(y_host);

    // (a lambda so that every early error return passes through the clean-up below)
    auto pipeline = [&]() -> int {
    int k = 0;
    for (int64_t b0 = 0; b0 < n_clips; b0 += per, ++k) {
        const int64_t nb = std::min<int64_t>(per, n_clips - b0);
        HostSlot &s = cache.slot[k % AES_HOST_SLOTS];
        if (k >= AES_HOST_SLOTS && (rc = drain_slot(s))) return rc;
        if ((rc = grow(&s.scratch, &s.scratch_cap,
                       (size_t)std::min<int64_t>(per, pl->grid_max) * pl->host.scratch_floats * sizeof(float), false))) return rc;
        if ((rc = grow(&s.d_in, &s.in_cap, (size_t)per * in_clip, false))) return rc;
        if ((rc = grow(&s.d_out, &s.out_cap, (size_t)per * out_clip, false))) return rc;
        const char *src = (const char *)x_host + (size_t)b0 * in_clip;
        char *dst = (char *)y_host + (size_t)b0 * out_clip;
        if (!pin_in) {
            if ((rc = grow(&s.h_in, &s.hin_cap, (size_t)per * in_clip, true))) return rc;
            memcpy(s.h_in, src, (size_t)nb * in_clip);
            src = (const char *)s.h_in;
        }
        AES_CUDA(cudaMemcpyAsync(s.d_in, src, (size_t)nb * in_clip, cudaMemcpyHostToDevice, s.stream));
        double *st_out = n_clips == 1 ? pl->d_state : nullptr;
        if ((rc = launch_chain(pl, s.d_in, in_fmt, s.d_out, out_fmt, nb, n_frames, (float *)s.scratch, s.stream, st_out,
                               n_clips))) return rc;
        if (st_out)
            AES_CUDA(cudaMemcpyAsync(pl->h_state, pl->d_state, (size_t)pl->host.n_state * sizeof(double),
                                     cudaMemcpyDeviceToHost, s.stream));
        if (!pin_out) {
            if ((rc = grow(&s.h_out, &s.hout_cap, (size_t)per * out_clip, true))) return rc;
            AES_CUDA(cudaMemcpyAsync(s.h_out, s.d_out, (size_t)nb * out_clip, cudaMemcpyDeviceToHost, s.stream));
            s.pend_dst = dst;
            s.pend_bytes = (size_t)nb * out_clip;
        } else {
            AES_CUDA(cudaMemcpyAsync(dst, s.d_out, (size_t)nb * out_clip, cudaMemcpyDeviceToHost, s.stream));
        }
        AES_CUDA(cudaEventRecord(s.done, s.stream));
    }
    const int used = std::min(k, AES_HOST_SLOTS);
    for (int i = 0; i < used; ++i)
        if ((rc = drain_slot(cache.slot[i]))) return rc;
    pl->state_valid = (n_clips == 1);
    return 0;
    };
    rc = pipeline();
    if (rc) {
        // an error left copies in flight into the caller's buffers and staged results pending: wait for every
        // slot and forget the pending host copies, so that a later call never writes through a stale pointer
        for (HostSlot &s : cache.slot) {
            if (s.stream) cudaStreamSynchronize(s.stream);
            s.pend_dst = nullptr;
            s.pend_bytes = 0;
        }
        cudaGetLastError();
        pl->state_valid = false;
    }
    return rc;
}

// Carried scalars of stage `stage` after the last single-clip aes_chain_process_host call:
// BIQUAD out[4*c+{0..3}] = x1,x2,y1,y2 of channel c; GATE out[0] = gain.
AES_EXPORT int aes_chain_final_state(aes_chain_plan *pl, int stage, double *out16)
{
    AES_REQUIRE(pl != nullptr && out16 != nullptr, "NULL argument");
    AES_REQUIRE(stage >= 0 && stage < pl->host.n_stages, "stage index out of range");
    AES_REQUIRE(pl->state_valid, "final state is only kept for single-clip host calls");
    memcpy(out16, pl->h_state + 16 * stage, 16 * sizeof(double));
    return 0;
}
