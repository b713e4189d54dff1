// aes_fast_build.h -- host side of the shape-specialised kernels: decides whether a
// compiled DevPlan fits a pre-instantiated shape and flattens its descriptors into the
// kernel-parameter block (FastArgs).  Pure host C++ (shared with tests/cpu_emu).
#pragma once
#include <string.h>
#include "aes_plan.h"
#include "aes_fast_kernel.cuh"

// X(C0, C1, C2, C3): the shapes instantiated for FR = 4 (T = 1024 frames).
#define AESF_DELAY_PF   AESF_CODE(AESK_DELAY, 0, 0, AES_MODE_REG, 0, 1)
#define AESF_DELAY_REG  AESF_CODE(AESK_DELAY, 0, 0, AES_MODE_REG, 0, 0)
#define AESF_DELAY_WALK AESF_CODE(AESK_DELAY, 0, 0, AES_MODE_WALK, 0, 0)
#define AESF_REVERB(pm) AESF_CODE(AESK_REVERB, 4, 2, 0, pm, 0)
#define AESF_BIQUAD     AESF_CODE(AESK_BIQUAD, 0, 0, 0, 0, 0)
#define AESF_GATE       AESF_CODE(AESK_GATE, 0, 0, 0, 0, 0)
#define AESF_OCTAVER    AESF_CODE(AESK_OCTAVER, 0, 0, 0, 0, 0)
#define AESF_DIST       AESF_CODE(AESK_DISTORTION, 0, 0, 0, 0, 0)

// X(C0, C1, C2, C3, TOPO): the shapes instantiated for FR = 4 (T = 1024 frames).  Reverb shapes come
// with the compile-time topologies of aes_fast_kernel.cuh first and the run-time variant last.
#define AESF_SHAPES(X)                                                                                  \
    X(AESF_DELAY_PF, 0, 0, 0, AESF_TOPO_NONE)                       /* Slapback Echo                  */ \
    X(AESF_DELAY_REG, 0, 0, 0, AESF_TOPO_NONE)                                                           \
    X(AESF_DELAY_WALK, 0, 0, 0, AESF_TOPO_NONE)                                                          \
    X(AESF_DELAY_PF, AESF_REVERB(0), 0, 0, AESF_TOPO_48K)          /* Rain Delay                     */ \
    X(AESF_DELAY_PF, AESF_REVERB(0), 0, 0, AESF_TOPO_44K)                                               \
    X(AESF_DELAY_PF, AESF_REVERB(0), 0, 0, AESF_TOPO_NONE)                                               \
    X(AESF_REVERB(0), 0, 0, 0, AESF_TOPO_48K)                      /* default reverb                 */ \
    X(AESF_REVERB(0), 0, 0, 0, AESF_TOPO_44K)                                                           \
    X(AESF_REVERB(0), 0, 0, 0, AESF_TOPO_NONE)                                                           \
    X(AESF_REVERB(1), 0, 0, 0, AESF_TOPO_NONE)                                                           \
    X(AESF_REVERB(3), 0, 0, 0, AESF_TOPO_48K)                      /* Cathedral (20 ms pre-delay)    */ \
    X(AESF_REVERB(3), 0, 0, 0, AESF_TOPO_44K)                                                           \
    X(AESF_REVERB(3), 0, 0, 0, AESF_TOPO_NONE)                                                           \
    X(AESF_BIQUAD, AESF_REVERB(0), 0, 0, AESF_TOPO_48K)            /* Guitar Filter                  */ \
    X(AESF_BIQUAD, AESF_REVERB(0), 0, 0, AESF_TOPO_44K)                                                 \
    X(AESF_BIQUAD, AESF_REVERB(0), 0, 0, AESF_TOPO_NONE)                                                 \
    X(AESF_DIST, AESF_OCTAVER, AESF_DELAY_PF, 0, AESF_TOPO_NONE)    /* BASELINE configs[2]            */ \
    X(AESF_BIQUAD, 0, 0, 0, AESF_TOPO_NONE)                                                              \
    X(AESF_BIQUAD, AESF_BIQUAD, AESF_BIQUAD, AESF_BIQUAD, AESF_TOPO_NONE)   /* BASELINE configs[1]    */ \
    X(AESF_DIST, 0, 0, 0, AESF_TOPO_NONE)

// Shapes built for THREE resident CTAs per SM (launch bounds 256 x 3: 80 registers, which they fit without a spill;
// their rings leave room for three).  Measured (r2at, 8192 clips): Robot Voice 204 -> 224 Gsamples/s, the octaver
// alone 327 -> 378 on 1184, the gate alone 3.48 -> 3.04 ms per 2048 clips (Clean Noise Removal's second stage);
// on exactly 4 x 296 clips Robot Voice loses 3 % (1184 clips are 2.67 waves of 444 CTAs).
// Not for distortion > octaver > delay: 231 -> 217 on 1184 clips.
#define AESF_SHAPES_3CTA(X)                                                                             \
    X(AESF_GATE, AESF_OCTAVER, AESF_DELAY_PF, 0, AESF_TOPO_NONE)    /* Robot Voice                    */ \
    X(AESF_GATE, 0, 0, 0, AESF_TOPO_NONE)                                                                \
    X(AESF_OCTAVER, 0, 0, 0, AESF_TOPO_NONE)

// compile-time topology (aes_fast_kernel.cuh) matching the plan's single reverb stage, or 0
static inline int aes_fast_topo(const DevPlan &p)
{
    int n = 0, topo = AESF_TOPO_NONE;
    for (int s = 0; s < p.n_stages; ++s) {
        const DevStage &d = p.stage[s];
        if (d.kind != AESK_REVERB) continue;
        if (++n > 1 || d.nc != 4 || d.na != 2) return AESF_TOPO_NONE;
        for (int t = AESF_TOPO_48K; t <= AESF_TOPO_44K && topo == AESF_TOPO_NONE; ++t) {
            bool same = true;
            for (int ch = 0; ch < 2; ++ch) {
                for (int c = 0; c < 4; ++c) {
                    const DevRing &r = p.ring[d.ring[ch][c]];
                    same = same && r.lag == aesf_topo_comb(t, ch, c) && r.off == aesf_topo_comb_off(t, ch, c);
                }
                for (int k = 0; k < 2; ++k) {
                    const DevRing &r = p.ring[d.apring[ch][k]];
                    same = same && r.lag == aesf_topo_ap(t, ch, k) && r.off == aesf_topo_ap_off(t, ch, k);
                }
            }
            if (same) topo = t;
        }
    }
    return topo;
}

// dynamic shared memory of a specialised kernel: the generic layout (superset of what the
// fast kernel carves) + the TMA staging area: 2 input tiles, 2 x [2][T+8] line samples, 2 mbarriers
static inline size_t aes_fast_smem_bytes(const DevPlan &p)
{
    const size_t T = (size_t)p.T;
    // + TMA staging: 2 input tiles, 2 x [2 ch][T+8] line samples when a feedback delay is prefetched, 2 mbarriers
    return aes_plan_smem_bytes(p) + 16 + (4 * T + (p.pf_stage >= 0 ? 4 * (T + 8) : 0)) * sizeof(float) + 2 * sizeof(unsigned long long) + 16;
}

static inline FRing aesf_ring(const DevRing &r)
{
    FRing f;
    f.off = (int)r.off; f.len = r.len; f.tinc = r.tinc; f.lag = r.lag;
    return f;
}

// Returns true and fills `fa`, `codes`, `lane_tab` (AESF_MAX_STAGES*32*FAST_LANE_STRIDE floats)
// when the plan can run on a specialised kernel (the caller still has to find the shape in
// its instantiation table).  Pointers / sizes of the launch are filled by the caller.
static inline bool aes_fast_build(const DevPlan &p, FastArgs *fa, int codes[AESF_MAX_STAGES], float *lane_tab)
{
    if (p.FR != 4 || p.n_stages < 1 || p.n_stages > AESF_MAX_STAGES) return false;
    memset(fa, 0, sizeof *fa);
    memset(lane_tab, 0, sizeof(float) * AESF_MAX_STAGES * 32 * FAST_LANE_STRIDE);
    for (int s = 0; s < AESF_MAX_STAGES; ++s) codes[s] = 0;
    int nwalk = 0;
    auto add_walk = [&](int rid, int *slot) -> bool {
        if (nwalk >= AESF_MAX_WALK) return false;
        fa->walk[nwalk] = aesf_ring(p.ring[rid]);
        *slot = nwalk++;
        return true;
    };
    for (int s = 0; s < p.n_stages; ++s) {
        const DevStage &d = p.stage[s];
        FastStage &f = fa->st[s];
        f.dry = d.dry; f.wet = d.wet; f.h = d.h; f.a = d.a; f.fb = d.fb; f.mix = d.mix; f.drive = d.drive;
        for (int i = 0; i < 6; ++i) f.hp[i] = d.hp[i];
        f.nscan = d.nscan; f.nxw = d.nxw;
        float *lt = lane_tab + (size_t)s * 32 * FAST_LANE_STRIDE;
        switch (d.kind) {
        case AESK_DELAY: {
            const DevRing &r0 = p.ring[d.ring[0][0]], &r1 = p.ring[d.ring[1][0]];
            if (r0.space != r1.space) return false;
            f.glob = r0.space == AES_SPACE_GLOBAL;
            if (d.mode == AES_MODE_REG) {
                f.ring[0][0] = aesf_ring(r0); f.ring[1][0] = aesf_ring(r1);
            } else {
                if (!add_walk(d.ring[0][0], &f.walk_pre[0]) || !add_walk(d.ring[1][0], &f.walk_pre[1])) return false;
            }
            codes[s] = AESF_CODE(AESK_DELAY, 0, 0, d.mode, 0, (p.pf_stage == s) ? 1 : 0);
            break;
        }
        case AESK_REVERB: {
            if (d.nc != 4 || d.na != 2) return false;
            int pm = 0;
            if (d.pre_ring[0] >= 0) {
                const DevRing &r0 = p.ring[d.pre_ring[0]], &r1 = p.ring[d.pre_ring[1]];
                if (r0.space != r1.space) return false;
                f.glob = r0.space == AES_SPACE_GLOBAL;
                pm = d.mode == AES_MODE_REG ? 1 : 3;
                f.pre[0] = aesf_ring(r0); f.pre[1] = aesf_ring(r1);
            }
            for (int ch = 0; ch < 2; ++ch) {
                for (int c = 0; c < 4; ++c) {
                    f.ring[ch][c] = aesf_ring(p.ring[d.ring[ch][c]]);
                    f.gs[ch][c] = (float)((double)d.g[ch][c] * (1.0 - (double)d.h));
                }
                for (int k = 0; k < 2; ++k) {
                    if (p.ring[d.apring[ch][k]].space != AES_SPACE_SMEM) return false;
                    if (!add_walk(d.apring[ch][k], &f.walk_ap[ch][k])) return false;
                }
            }
            for (int l = 0; l < 32; ++l) lt[l * FAST_LANE_STRIDE] = d.hlane[l];
            codes[s] = AESF_CODE(AESK_REVERB, 4, 2, 0, pm, 0);
            break;
        }
        case AESK_BIQUAD:
            for (int i = 0; i < 5; ++i) f.bq[i] = d.bq[i];
            memcpy(f.bq_pow, d.bq_pow, sizeof f.bq_pow);
            {   // How far back the state still matters: a filter whose A^(4*2^s) has decayed below
                // 2^-40 needs no scan step over 2^s lanes or beyond, one whose A^128 has decayed needs
                // only the previous warp's total (the damped combs truncate their scan the same way).
                // A step is dropped only if every longer span has decayed too (A^k of a stable
                // biquad may grow before it decays).
                const double eps = ldexp(1.0, -40);
                auto small = [&](const double *m) { return fabs(m[0]) < eps && fabs(m[1]) < eps && fabs(m[2]) < eps && fabs(m[3]) < eps; };
                double wp[4], acc[4];
                memcpy(wp, d.bq_pow[5], sizeof wp);                     // A^128
                memcpy(acc, wp, sizeof acc);
                f.nxw = 1;
                while (f.nxw < 8 && !small(acc)) { aes_mat2_mul(acc, wp, acc); ++f.nxw; }
                f.nscan = 5;
                if (small(wp)) while (f.nscan > 0 && small(d.bq_pow[f.nscan - 1])) --f.nscan;
            }
            {   // row 0 of (A^T)^j, A^T = [[-a1, 1], [-a2, 0]]
                double r0 = 1.0, r1 = 0.0;
                for (int j = 0; j < 4; ++j) {
                    f.bq_row[j][0] = r0; f.bq_row[j][1] = r1;
                    const double n0 = -r0 * d.bq[3] - r1 * d.bq[4], n1 = r0;
                    r0 = n0; r1 = n1;
                }
            }
            for (int l = 0; l < 32; ++l) memcpy(lt + l * FAST_LANE_STRIDE + 4, d.bq_lane[l], 4 * sizeof(double));
            for (int i = 0; i < 8; ++i) fa->init[s][i] = d.init[i];
            codes[s] = AESF_BIQUAD;
            break;
        case AESK_GATE:
            f.thr = d.thr; f.att = d.att; f.rel = d.rel;
            fa->init[s][0] = d.init[0];
            codes[s] = AESF_GATE;
            break;
        case AESK_OCTAVER:
            f.ring[0][0] = aesf_ring(p.ring[d.ring[0][0]]);
            f.oct_size = d.oct_size; f.oct_mask = d.oct_mask;
            f.ph0 = d.ph0; f.step = d.step; f.fsize = d.fsize;
            codes[s] = AESF_OCTAVER;
            break;
        case AESK_DISTORTION:
            codes[s] = AESF_DIST;
            break;
        default:
            return false;
        }
    }
    fa->n_walk = nwalk;
    fa->n_stages = p.n_stages;
    fa->smem_floats = p.smem_floats;
    fa->scratch_floats = p.scratch_floats;
    return true;
}
