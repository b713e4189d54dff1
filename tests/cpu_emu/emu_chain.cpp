// emu_chain.cpp -- TEST INFRASTRUCTURE ONLY.  Runs the product's fused chain kernel
// source (csrc/aes_chain_kernel.cuh) and plan compiler (csrc/aes_plan_build.h) on
// the CPU through the cuda_emu.h shim, so kernel logic can be checked against the
// oracle on a machine without a GPU.  Never linked into libaesim.so.
#include "cuda_emu.h"
#include "../../audio-effects-simulator_b200/csrc/aes_plan_build.h"
#include "../../audio-effects-simulator_b200/csrc/aes_chain_kernel.cuh"
#include "../../audio-effects-simulator_b200/csrc/aes_fast_build.h"
#include "../../audio-effects-simulator_b200/csrc/aes_rv_build.h"
#include "../../audio-effects-simulator_b200/csrc/aes_biquad_build.h"
#include "../../audio-effects-simulator_b200/csrc/aes_biquad_seq.cuh"
#include "../../audio-effects-simulator_b200/csrc/aes_convreverb.cuh"
#include "../../audio-effects-simulator_b200/csrc/aes_analysis.cuh"
#include "../../audio-effects-simulator_b200/csrc/aes_spectral.cuh"

static char g_err[512];

template <int K> static void entry(void *p) { aes_chain_body<K>(*reinterpret_cast<ChainArgs *>(p)); }
template <int C0, int C1, int C2, int C3, int MP> static void fentry(void *p)
{
    aes_fast_body<4, C0, C1, C2, C3, MP>(*reinterpret_cast<FastArgs *>(p));
}
struct FastShape { int c[4]; int topo; void (*fn)(void *); };
#define X(c0, c1, c2, c3, mp) { { c0, c1, c2, c3 }, mp, fentry<c0, c1, c2, c3, mp> },
static const FastShape g_shapes[] = { AESF_SHAPES(X) AESF_SHAPES_3CTA(X) };
#undef X
template <int TOPO, int PRE, int PM> static void rventry(void *p) { aes_rv_body<TOPO, PRE, PM>(*reinterpret_cast<FastArgs *>(p)); }
struct RvShape { int topo, pre, pm; void (*fn)(void *); };
#define X(topo, pre, pm) { topo, pre, pm, rventry<topo, pre, pm> },
static const RvShape g_rv_shapes[] = { AESRV_SHAPES(X) };
#undef X
static int g_last_fast = 0, g_last_topo = 0;
extern "C" __attribute__((visibility("default"))) int emu_last_topo() { return g_last_topo; }
extern "C" __attribute__((visibility("default"))) int emu_last_was_fast() { return g_last_fast; }

extern "C" __attribute__((visibility("default")))
const char *emu_last_error() { return g_err; }

extern "C" __attribute__((visibility("default")))
int emu_chain_run(const aes_stage_desc *stages, int n, int fs, const void *x, int in_fmt, void *y,
                  int out_fmt, long long B, long long N, int grid, double *state_out)
{
    static DevPlan plan;
    int rc = aes_build_devplan(stages, n, fs, &plan, g_err, sizeof g_err);
    if (rc) return rc;
    if (grid < 1) grid = 1;
    std::vector<float> scratch((size_t)grid * plan.scratch_floats, 1e30f);   // poison
    ChainArgs a{ &plan, x, y, B, N, scratch.data(), in_fmt, out_fmt, state_out };
    size_t smem = aes_plan_smem_bytes(plan);
    g_last_fast = 0;
    static FastArgs fa;
    static float lane_tab[AESF_MAX_STAGES * 32 * FAST_LANE_STRIDE];
    int codes[4];
    if (!getenv("AES_NO_FAST") && aes_fast_build(plan, &fa, codes, lane_tab)) {
        const int topo = aes_fast_topo(plan);
        int rv_pre = 0, rv_pm = 0;
        if (!getenv("AES_NO_RV") && aes_rv_shape(fa, codes, topo, &rv_pre, &rv_pm)) {
            for (const RvShape &sh : g_rv_shapes) {
                if (sh.topo != topo || sh.pre != rv_pre || sh.pm != rv_pm) continue;
                g_last_topo = sh.topo;
                fa.x = x; fa.y = y; fa.B = B; fa.N = N; fa.scratch = scratch.data(); fa.lane_tab = lane_tab;
                fa.state_out = state_out; fa.in_fmt = in_fmt; fa.out_fmt = out_fmt;
                emu::launch(sh.fn, &fa, grid, AESRV_NT, aes_rv_smem_bytes(plan.smem_floats, rv_pre));
                g_last_fast = 2;
                return 0;
            }
        }
        for (const FastShape &sh : g_shapes) {
            if (memcmp(sh.c, codes, sizeof codes) != 0) continue;
            if (sh.topo != AESF_TOPO_NONE && sh.topo != topo) continue;
            g_last_topo = sh.topo;
            fa.x = x; fa.y = y; fa.B = B; fa.N = N; fa.scratch = scratch.data(); fa.lane_tab = lane_tab;
            fa.state_out = state_out; fa.in_fmt = in_fmt; fa.out_fmt = out_fmt;
            emu::launch(sh.fn, &fa, grid, AES_NT, aes_fast_smem_bytes(plan));
            g_last_fast = 1;
            return 0;
        }
    }
    switch (plan.FR) {
    case 4: emu::launch(entry<4>, &a, grid, AES_NT, smem); break;
    case 2: emu::launch(entry<2>, &a, grid, AES_NT, smem); break;
    case 1: emu::launch(entry<1>, &a, grid, AES_NT, smem); break;
    default: snprintf(g_err, sizeof g_err, "bad FR"); return -1;
    }
    return 0;
}

extern "C" __attribute__((visibility("default")))
int emu_plan_info(const aes_stage_desc *stages, int n, int fs, int *T, int *smem_bytes, long long *scratch_floats)
{
    static DevPlan plan;
    int rc = aes_build_devplan(stages, n, fs, &plan, g_err, sizeof g_err);
    if (rc) return rc;
    *T = plan.T; *smem_bytes = (int)aes_plan_smem_bytes(plan); *scratch_floats = plan.scratch_floats;
    return 0;
}

static void bq_entry(void *p) { aes_biquad_scan_body(*reinterpret_cast<BqArgs *>(p)); }

// time-parallel biquad cascade (aes_biquad_scan.cuh) on the emulator
extern "C" __attribute__((visibility("default")))
int emu_biquad_scan(const float *x, float *y, long long B, long long N, int n_stages, const double *coeffs5,
                    const double *dfi_state, int dbg_skip)
{
    if (n_stages < 1 || n_stages > AESB_MAX_STAGES) return -1;
    static BqArgs a;
    std::vector<double> lane_pw((size_t)n_stages * 128), tile_pw((size_t)n_stages * AESB_LBW * 4);
    aes_biquad_build(n_stages, coeffs5, dfi_state, &a, lane_pw.data(), tile_pw.data());
    const long long nt = (N + AESB_T - 1) / AESB_T;
    const size_t recs = (size_t)(B * n_stages * nt);
    std::vector<double> agg(recs * 4, 1e300), inc(recs * 4, 1e300);
    std::vector<int> flag(recs, 0);
    unsigned ticket = 40;                                        // the counter is never reset: a launch starts at its base
    std::vector<BqRec> rec16(recs * 4);
    for (auto &r : rec16) { r.v = 1e300; r.tag = 6; }            // records of an earlier launch
    a.rec16 = rec16.data(); a.epoch = 7; a.ticket_base = 40;
    a.x = x; a.y = y; a.N = N; a.n_tiles = nt; a.B = B; a.dbg_skip = dbg_skip;
    a.agg = agg.data(); a.inc = inc.data(); a.flag = flag.data(); a.ticket = &ticket;
    a.lane_pw = lane_pw.data(); a.tile_pw = tile_pw.data(); a.final_state = nullptr;
    emu::launch(bq_entry, &a, (unsigned)(B * nt), AESB_NT, AESB_SMEM_DOUBLES * sizeof(double));
    return 0;
}

// ---- sequential batch biquad kernel (aes_biquad_seq.cuh): K segments per clip, `warm` frames of warm-up
template <int NS> static void bqseq_entry(void *p) { aes_biquad_seq_body<NS>(*reinterpret_cast<BqSeqArgs *>(p)); }
extern "C" __attribute__((visibility("default")))
int emu_biquad_seq(const float *x, float *y, long long B, long long N, int n_stages, const double *coeffs5,
                   const double *dfi_state, int K, long long warm, int max_ctas)
{
    if (n_stages < 1 || n_stages > AESQ_MAX_STAGES || (N & 1) || K < 1 || warm % AESQ_CH) return -1;
    static BqSeqArgs q;
    memset(&q, 0, sizeof q);
    q.x = x; q.y = y; q.B = B; q.N = N; q.K = K; q.warm = K > 1 ? warm : 0; q.n_stages = n_stages;
    q.seg = (N / K + AESQ_CH - 1) / AESQ_CH * AESQ_CH;
    for (int s = 0; s < n_stages; ++s) {
        for (int i = 0; i < 5; ++i) q.bq[s][i] = coeffs5[5 * s + i];
        for (int i = 0; i < 8; ++i) q.init[s][i] = dfi_state ? dfi_state[8 * s + i] : 0.0;
    }
    unsigned grid = (unsigned)((B * K + AESQ_WARPS * 32 - 1) / (AESQ_WARPS * 32));
    if (max_ctas > 0 && grid > (unsigned)max_ctas) grid = (unsigned)max_ctas;        // persistent threads take several items
    switch (n_stages) {
    case 1: emu::launch(bqseq_entry<1>, &q, grid, AESQ_WARPS * 32, AESQ_SMEM_BYTES); break;
    case 2: emu::launch(bqseq_entry<2>, &q, grid, AESQ_WARPS * 32, AESQ_SMEM_BYTES); break;
    case 3: emu::launch(bqseq_entry<3>, &q, grid, AESQ_WARPS * 32, AESQ_SMEM_BYTES); break;
    default: emu::launch(bqseq_entry<4>, &q, grid, AESQ_WARPS * 32, AESQ_SMEM_BYTES); break;
    }
    return 0;
}

// look-back depth the host tables choose per stage (aes_biquad_build.h)
extern "C" __attribute__((visibility("default")))
int emu_biquad_lookback_depth(int n_stages, const double *coeffs5, int *out)
{
    if (n_stages < 1 || n_stages > AESB_MAX_STAGES) return -1;
    static BqArgs a;
    std::vector<double> lane_pw((size_t)n_stages * 128), tile_pw((size_t)n_stages * AESB_LBW * 4);
    aes_biquad_build(n_stages, coeffs5, nullptr, &a, lane_pw.data(), tile_pw.data());
    for (int s = 0; s < n_stages; ++s) out[s] = a.st[s].lb_k;
    return 0;
}

// ---- IR-convolution reverb (aes_convreverb.cuh) on the emulator, FFT size 2^8 / 2^11 / 2^14 -----
struct ConvLaunch { ConvArgs a; int NS, p0, cpi, acc; const float *ir; int n_taps; cpx *H; const cpx *tw; };
template <int R> static void conv_k0(void *p) { ConvLaunch *l = reinterpret_cast<ConvLaunch *>(p); aesc_ir_prep_body<R>(l->ir, l->n_taps, l->H, l->tw); }
template <int R> static void conv_k1(void *p) { aesc_fwd_body<R>(reinterpret_cast<ConvLaunch *>(p)->a); }
template <int PC> static void conv_k2(void *p) { ConvLaunch *l = reinterpret_cast<ConvLaunch *>(p); aesc_mac_body<PC>(l->a, l->NS, l->p0, l->cpi, l->acc); }
template <int R> static void conv_k3(void *p) { aesc_inv_body<R>(reinterpret_cast<ConvLaunch *>(p)->a); }

template <int R>
static int emu_conv_run(const float *ir, long long n_taps, const float *x, float *y, long long B, long long Nf,
                        float dry, float wet, int pc, int mac_grid)
{
    using G = AescGeo<R>;
    constexpr int N = G::N, NS = G::NS, BK = G::BK;
    const int P = (int)((n_taps + BK - 1) / BK);
    const int Ppad = (P + pc - 1) / pc * pc;
    const int nblk = (int)((Nf + BK - 1) / BK);
    std::vector<cpx> tw(N / 2), H((size_t)Ppad * NS + AESC_MT), Z((size_t)B * nblk * NS), W((size_t)B * nblk * NS);
    for (auto &h : H) { h.x = 0.f; h.y = 0.f; }
    for (auto &w : W) { w.x = 7.f; w.y = 7.f; }
    for (int q = 0; q < N / 2; ++q) { double ang = -2.0 * M_PI * q / N; tw[q].x = (float)cos(ang); tw[q].y = (float)sin(ang); }
    ConvLaunch l;
    l.ir = ir; l.n_taps = (int)n_taps; l.H = H.data(); l.tw = tw.data(); l.NS = NS; l.cpi = 2;
    const size_t fft_smem = (size_t)G::SMEM_CPX * sizeof(cpx);
    emu::launch(conv_k0<R>, &l, P, G::NT, fft_smem);
    l.a.x = x; l.a.y = y; l.a.Z = Z.data(); l.a.W = W.data(); l.a.H = H.data(); l.a.tw = tw.data();
    l.a.B = B; l.a.Nf = Nf; l.a.nblk = nblk; l.a.P = Ppad; l.a.dry = dry; l.a.wet = wet;
    emu::launch(conv_k1<R>, &l, (unsigned)(B * nblk), G::NT, fft_smem);
    for (int p0 = 0; p0 < Ppad && p0 < nblk; p0 += pc) {
        l.p0 = p0; l.acc = p0 > 0;
        emu::launch(pc == 18 ? conv_k2<18> : pc == 9 ? conv_k2<9> : conv_k2<4>, &l, (unsigned)mac_grid, AESC_MT,
                    (size_t)2 * pc * AESC_MT * sizeof(cpx) + 16);
    }
    emu::launch(conv_k3<R>, &l, (unsigned)(B * nblk), G::NT, fft_smem);
    return 0;
}

// log2n: 8, 11 or 14; pc: partitions per MAC pass (4, 9, 18); mac_grid: persistent MAC CTAs
extern "C" __attribute__((visibility("default")))
int emu_convreverb(const float *ir, long long n_taps, const float *x, float *y, long long B, long long Nf,
                   float dry, float wet, int log2n, int pc, int mac_grid)
{
    if (pc != 4 && pc != 9 && pc != 18) return -1;
    if (log2n == 8) return emu_conv_run<8>(ir, n_taps, x, y, B, Nf, dry, wet, pc, mac_grid);
    if (log2n == 11) return emu_conv_run<16>(ir, n_taps, x, y, B, Nf, dry, wet, pc, mac_grid);
    if (log2n == 14) return emu_conv_run<32>(ir, n_taps, x, y, B, Nf, dry, wet, pc, mac_grid);
    return -1;
}

// ---- spectrum / chromagram analysis (aes_analysis.cuh) on the emulator --------------------------
template <int R> static void ana_k(void *p) { aesa_body<R>(*reinterpret_cast<AnalysisArgs *>(p)); }
extern "C" __attribute__((visibility("default")))
int emu_spectrum_chroma(const float *a, const float *b, long long n_pairs, long long n_samples, int n_fft, double fs,
                        float *db, float *lin, float *chroma, float *peak)
{
    AnalysisArgs q;
    q.a = a; q.b = b; q.db = db; q.lin = lin; q.chroma = chroma; q.peak_freq = peak; q.n_samples = n_samples; q.sample_rate = fs;
    if (n_fft == 256) emu::launch(ana_k<8>, &q, (unsigned)n_pairs, AescGeo<8>::NT, AescGeo<8>::SMEM_CPX * sizeof(cpx));
    else if (n_fft == 2048) emu::launch(ana_k<16>, &q, (unsigned)n_pairs, AescGeo<16>::NT, AescGeo<16>::SMEM_CPX * sizeof(cpx));
    else if (n_fft == 16384) emu::launch(ana_k<32>, &q, (unsigned)n_pairs, AescGeo<32>::NT, AescGeo<32>::SMEM_CPX * sizeof(cpx));
    else return -1;
    return 0;
}

// ---- SpectralFilter kernels (aes_spectral.cuh) on the emulator ---------------------------------
struct SpecLaunch { SpecArgs a; int st, inv, mul; };
static void sp_load(void *p) { aess_load_body(reinterpret_cast<SpecLaunch *>(p)->a); }
template <int R, int IN, int OUT> static void sp_pass(void *p) { SpecLaunch *l = reinterpret_cast<SpecLaunch *>(p); aess_global_pass_body<R, IN, OUT>(l->a, l->st, l->inv); }
template <int IN, int OUT> static void emu_spec_pass(SpecLaunch &l, int s, int r, int inv)
{
    l.st = s; l.inv = inv;
    emu::launch(r == 3 ? sp_pass<3, IN, OUT> : r == 2 ? sp_pass<2, IN, OUT> : sp_pass<1, IN, OUT>, &l, 4, 256, 0);
}
static void sp_local(void *p) { SpecLaunch *l = reinterpret_cast<SpecLaunch *>(p); aess_local_body(l->a, l->inv, l->mul); }
static void sp_gate(void *p) { aess_gate_body(reinterpret_cast<SpecLaunch *>(p)->a); }
static void sp_zero(void *p) { aess_zero_pad_body(reinterpret_cast<SpecLaunch *>(p)->a); }
static void sp_store(void *p) { aess_store_body(reinterpret_cast<SpecLaunch *>(p)->a); }

static void emu_spec_fft(SpecLaunch &l, int inverse, int mul, int in_mode = 0, int out_mode = 0)
{
    const unsigned chunks = (unsigned)std::min<long long>(((long long)l.a.nb * l.a.P / 1024 + 1) / 2, 3);   // persistent: 3 CTAs
    if (!inverse) {
        for (int s = 0, r; s <= l.a.L - 11; s += r) {
            r = aess_pass_radix(l.a.L - 10 - s);
            if (s == 0 && in_mode == 1) emu_spec_pass<1, 0>(l, s, r, 0);
            else if (s == 0 && in_mode == 2) emu_spec_pass<2, 0>(l, s, r, 0);
            else emu_spec_pass<0, 0>(l, s, r, 0);
        }
        l.inv = 0; l.mul = mul; emu::launch(sp_local, &l, chunks, AESS_LOCAL_NT, AESS_LOCAL_SMEM_CPX * sizeof(cpx));
    } else {
        l.inv = 1; l.mul = 0; emu::launch(sp_local, &l, chunks, AESS_LOCAL_NT, AESS_LOCAL_SMEM_CPX * sizeof(cpx));
        for (int s = 10, r; s < l.a.L; s += r) {
            r = aess_pass_radix(l.a.L - s);
            if (s + r == l.a.L && out_mode == 1) emu_spec_pass<0, 1>(l, s, r, 1);
            else emu_spec_pass<0, 0>(l, s, r, 1);
        }
    }
}

// frames: [nb][M] windowed analysis frames; mask [nb][M/2+1] in/out; y [nb][M] out
extern "C" __attribute__((visibility("default")))
int emu_spectral_frames(const float *frames, float *mask, float *y, long long M, int nb, float thr, float red, float alpha)
{
    long long P = 1024; int L = 10;
    while (P < 2 * M - 1) { P <<= 1; ++L; }
    std::vector<cpx> chirp((size_t)M), v((size_t)P), twP((size_t)P / 2), tw1k(512), buf((size_t)((nb + 1) / 2) * P);
    for (long long n = 0; n < M; ++n) {
        const long long r = (long long)(((unsigned long long)n * (unsigned long long)n) % (unsigned long long)(2 * M));
        const double ang = -M_PI * (double)r / (double)M;
        chirp[n].x = (float)cos(ang); chirp[n].y = (float)sin(ang);
    }
    for (auto &e : v) { e.x = 0.f; e.y = 0.f; }
    for (long long m = 0; m < M; ++m) { cpx c; c.x = chirp[m].x / (float)P; c.y = -chirp[m].y / (float)P; v[m] = c; if (m > 0) v[P - m] = c; }
    for (long long q = 0; q < P / 2; ++q) { const double ang = -2.0 * M_PI * (double)q / (double)P; twP[q].x = (float)cos(ang); twP[q].y = (float)sin(ang); }
    for (int q = 0; q < 512; ++q) { const double ang = -2.0 * M_PI * q / 1024.0; tw1k[q].x = (float)cos(ang); tw1k[q].y = (float)sin(ang); }
    SpecLaunch l; memset(&l, 0, sizeof l);
    l.a.twP = twP.data(); l.a.tw1k = tw1k.data(); l.a.M = M; l.a.P = P; l.a.L = L;
    l.a.buf = v.data(); l.a.nb = 1; l.a.nf = 1;
    emu_spec_fft(l, 0, 0);                               // vhat = FFT(v)/P in place
    l.a.buf = buf.data(); l.a.vhat = v.data(); l.a.chirp = chirp.data(); l.a.frames = frames; l.a.mask = mask; l.a.out = y;
    l.a.nb = (nb + 1) / 2; l.a.nf = nb; l.a.thr = thr; l.a.red = red; l.a.alpha = alpha;
    const bool fuse = L > 10;                            // same orchestration as spec_process (aes_spectral.cu)
    if (!fuse) emu::launch(sp_load, &l, 4, 256, 0);
    emu_spec_fft(l, 0, 1, fuse ? 1 : 0, 0); emu_spec_fft(l, 1, 0);
    emu::launch(sp_gate, &l, 4, 256, 0);
    if (!fuse) emu::launch(sp_zero, &l, 4, 256, 0);
    emu_spec_fft(l, 0, 1, fuse ? 2 : 0, 0); emu_spec_fft(l, 1, 0, 0, fuse ? 1 : 0);
    if (!fuse) emu::launch(sp_store, &l, 4, 256, 0);
    return 0;
}

// ---- smooth-length SpectralFilter path (aes_spectral_smooth.cuh) on the emulator ------------------
#include "../../audio-effects-simulator_b200/csrc/aes_spectral_smooth.cuh"
template <int SHAPE> static void sm_k1(void *p) { aesm_cols_fwd_body<SHAPE>(*reinterpret_cast<SmoothArgs *>(p)); }
template <int SHAPE> static void sm_k2(void *p) { aesm_rows_body<SHAPE>(*reinterpret_cast<SmoothArgs *>(p)); }
template <int SHAPE> static void sm_k3(void *p) { aesm_cols_inv_body<SHAPE>(*reinterpret_cast<SmoothArgs *>(p)); }
static void sm_k2_rows10(void *p) { aesm_rows10_body<1, 2>(*reinterpret_cast<SmoothArgs *>(p)); }
static void sm_k2_rows10_one_buffer(void *p) { aesm_rows10_body<2, 1, 1>(*reinterpret_cast<SmoothArgs *>(p)); }    // + bulk-copy prefetch

// mode 1: frames [nf][M] raw (window applied when `window` != null), mask [nf][M/2+1] or null, y [nf][M]
// mode 2: frames = clips [nf][M/2][2], y = [nf][M/2][2]; returns 1 when M has no smooth split
extern "C" __attribute__((visibility("default")))
int emu_spectral_smooth(int mode, const float *frames, const float *window, float *mask, float *y, long long M, int nf,
                        float thr, float red, float alpha, int *n1_out, int *n2_out)
{
    int n1 = 0, n2 = 0;
    if (!aesm_split(M, &n1, &n2)) return 1;
    if (n1_out) *n1_out = n1;
    if (n2_out) *n2_out = n2;
    SmoothArgs a; memset(&a, 0, sizeof a);
    aesm_build_fft(n1, &a.f1); aesm_build_fft(n2, &a.f2);
    const long long nhi = (M + 1023) / 1024;
    std::vector<cpx> tw1((size_t)a.f1.tsize + 1), tw2((size_t)a.f2.tsize + 1), twlo(1024), twhi((size_t)nhi), buf((size_t)((nf + 1) / 2) * M);
    auto fill = [](std::vector<cpx> &t, double step, double denom) {
        for (size_t j = 0; j < t.size(); ++j) { const double ang = -2.0 * M_PI * (double)j * step / denom; t[j].x = (float)cos(ang); t[j].y = (float)sin(ang); }
    };
    aesm_fill_twiddles(a.f1, tw1.data()); aesm_fill_twiddles(a.f2, tw2.data()); fill(twlo, 1.0, (double)M); fill(twhi, 1024.0, (double)M);
    std::vector<int> rev1(n1), rev2(n2);
    for (int k = 0; k < n1; ++k) rev1[k] = aesm_rev(k, a.f1);
    for (int k = 0; k < n2; ++k) rev2[k] = aesm_rev(k, a.f2);
    a.rev1 = rev1.data(); a.rev2 = rev2.data();
    a.buf = buf.data(); a.tw1 = tw1.data(); a.tw2 = tw2.data(); a.twlo = twlo.data(); a.twhi = twhi.data();
    a.window = window; a.mask = mask; a.M = (int)M; a.n1 = n1; a.n2 = n2; a.nf = nf; a.np = (nf + 1) / 2; a.mode = mode;
    a.thr = thr; a.red = red; a.alpha = alpha;
    if (mode == 2) { a.clips = frames; a.yclips = y; } else { a.frames = frames; a.out = y; }
    const bool st = aesm_static_shape(a.f1, a.f2) == AESM_SHAPE_960x1000;       // the compile-time 960 x 1000 kernels
    emu::launch(st ? sm_k1<AESM_SHAPE_960x1000> : sm_k1<0>, &a, 3, AESM_NTC, (size_t)n1 * AESM_C * sizeof(cpx));
    if (st && !getenv("AES_SPECTRAL_ROWS_SMEM")) {
        // the register-resident row-pair kernel + the two self-paired rows on the shared-memory one, as smooth_process does
        std::vector<cpx> t10(AESR_TW_ENTRIES);
        aesm_fill_rows10(t10.data());
        a.tw10 = t10.data();
        // (one pair per CTA with alternating exchange buffers, or AES_EMU_ROWS10_ONE_BUFFER: two pairs per CTA, one buffer)
        if (getenv("AES_EMU_ROWS10_ONE_BUFFER")) emu::launch(sm_k2_rows10_one_buffer, &a, 2, AESR_NT_OF(2), AESR_SMEM_OF3(2, 1, 1));
        else emu::launch(sm_k2_rows10, &a, 3, AESR_NT_OF(1), AESR_SMEM_OF2(1, 2));
        a.rows_self_only = 1;
        emu::launch(sm_k2<AESM_SHAPE_960x1000>, &a, 2, AESM_NT, (size_t)2 * n2 * sizeof(cpx));
        a.rows_self_only = 0;
    } else {
        emu::launch(st ? sm_k2<AESM_SHAPE_960x1000> : sm_k2<0>, &a, 3, AESM_NT, (size_t)2 * n2 * sizeof(cpx));
    }
    emu::launch(st ? sm_k3<AESM_SHAPE_960x1000> : sm_k3<0>, &a, 3, AESM_NT, (size_t)n1 * AESM_C * sizeof(cpx));
    return 0;
}
