// aes_plan_build.h -- host-side "chain compiler": turns the resolved stage
// descriptors of include/aesim.h into the DevPlan the fused kernel executes
// (tile size, ring placement, scan tables).  Pure host C++ (no CUDA calls) so the
// GPU-less kernel-logic tests (tests/cpu_emu) build the very same plans.
#pragma once
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../../include/aesim.h"
#include "aes_plan.h"

#define AES_SMEM_LIMIT (227 * 1024)
#define AES_SMEM_RING_MAX_LEN 4096     // longer lines go to the per-CTA global scratch

static inline void aes_mat2_mul(const double a[4], const double b[4], double o[4])
{
    double r[4] = { a[0] * b[0] + a[1] * b[2], a[0] * b[1] + a[1] * b[3],
                    a[2] * b[0] + a[3] * b[2], a[2] * b[1] + a[3] * b[3] };
    memcpy(o, r, sizeof r);
}

static inline void aes_mat2_pow(const double a[4], long long n, double o[4])
{
    double r[4] = { 1, 0, 0, 1 }, b[4];
    memcpy(b, a, sizeof b);
    while (n > 0) {
        if (n & 1) aes_mat2_mul(r, b, r);
        aes_mat2_mul(b, b, b);
        n >>= 1;
    }
    memcpy(o, r, sizeof r);
}

struct AesPlanBuilder {
    DevPlan *p;
    long long smem_off = 0, glob_off = 0;
    char *err;
    size_t errlen;
    int prio[AES_MAX_RINGS] = {};      // shared-memory layout order: 0 reverb combs, 1 all-passes, 2 everything else
    int next_prio = 2;

    int fail(const char *msg) { snprintf(err, errlen, "%s", msg); return AES_ERR_INVALID; }

    // WALK ring: period == lag.  REG ring: period = roundup4(lag) (+ extra), see aes_plan.h.
    int add_ring(long long lag, long long period, bool allow_global, int *id)
    {
        if (lag < 1) return fail("delay line length must be >= 1");
        if (period > 0x0fffffff) return fail("delay line too long");
        if (p->n_rings >= AES_MAX_RINGS) return fail("too many delay lines in one chain");
        DevRing &r = p->ring[p->n_rings];
        r.len = (int)period;
        r.lag = (int)lag;
        r.tinc = (int)(p->T % period);
        if (allow_global && period > AES_SMEM_RING_MAX_LEN) {
            r.space = AES_SPACE_GLOBAL;
            r.off = glob_off;
            glob_off += (period + 31) & ~31LL;        // 128-byte aligned lines
        } else {
            r.space = AES_SPACE_SMEM;
            r.off = smem_off;
            smem_off += (period + 3) & ~3LL;
        }
        prio[p->n_rings] = next_prio;
        *id = p->n_rings++;
        return 0;
    }
    // Final shared-memory offsets: the reverb's comb rings first, then its all-pass rings, then the
    // rest, each group in creation order -- so that a chain with the default reverb topology has
    // those rings at the fixed offsets the compile-time topologies assume (aes_fast_kernel.cuh).
    void layout_smem()
    {
        long long off = 0;
        for (int pass = 0; pass < 3; ++pass)
            for (int i = 0; i < p->n_rings; ++i) {
                DevRing &r = p->ring[i];
                if (r.space != AES_SPACE_SMEM || prio[i] != pass) continue;
                r.off = off;
                off += ((r.len + 3) & ~3LL) + (pass == 0 ? AES_COMB_RING_PAD : 0);
            }
        smem_off = off;
    }
    int add_walk(long long lag, bool allow_global, int *id) { return add_ring(lag, lag, allow_global, id); }
    // register-path line without a barrier between its reads and writes: the period
    // must cover lag + T so a tile's writes never land on samples the tile still reads
    int add_reg_line(long long lag, int *id) { return add_ring(lag, ((lag + 3) & ~3LL) + p->T, true, id); }
};

static inline long long aes_roundup4(long long v) { return (v + 3) & ~3LL; }

// Returns 0 or a negative aes_status; on failure `err` holds the reason.
static inline int aes_build_devplan(const aes_stage_desc *stages, int n, int fs, DevPlan *p,
                                    char *err, size_t errlen)
{
    (void)fs;
    memset(p, 0, sizeof *p);
    AesPlanBuilder B;
    B.p = p; B.err = err; B.errlen = errlen;
    if (n < 0 || n > AES_MAX_STAGES) return B.fail("a chain holds 0..16 stages");

    // Tile size: every damped comb needs its lag >= T.
    long long min_comb = 1LL << 40;
    for (int s = 0; s < n; ++s) {
        if (stages[s].kind != AES_STAGE_REVERB) continue;
        const aes_stage_desc &d = stages[s];
        if (d.q[0] < 0 || d.q[0] > AES_MAX_COMB || d.q[1] < 0 || d.q[1] > AES_MAX_AP)
            return B.fail("reverb: at most 8 combs and 4 all-passes per side");
        for (int side = 0; side < 2; ++side)
            for (int c = 0; c < d.q[0]; ++c)
                if (d.q[4 + 8 * side + c] < min_comb) min_comb = d.q[4 + 8 * side + c];
    }
    int FR = 4;
    while (FR > 1 && (long long)AES_NT * FR > min_comb) FR >>= 1;
    if ((long long)AES_NT * FR > min_comb) {
        snprintf(err, errlen, "reverb comb line of %lld samples is shorter than the smallest tile (256)",
                 min_comb);
        return AES_ERR_UNSUPPORTED;
    }
    p->FR = FR;
    p->T = AES_NT * FR;
    p->n_stages = n;
    p->n_state = 16 * n;
    p->pf_stage = -1;
    const int T = p->T;

    for (int s = 0; s < n; ++s) {
        const aes_stage_desc &d = stages[s];
        DevStage &st = p->stage[s];
        st.kind = d.kind;
        st.pre_ring[0] = st.pre_ring[1] = -1;
        int rc;
        switch (d.kind) {
        case AES_STAGE_DELAY: {
            if (d.q[0] < 1 || d.q[1] < 1) return B.fail("delay: lag must be >= 1 sample");
            st.mode = (d.q[0] >= T && d.q[1] >= T) ? AES_MODE_REG : AES_MODE_WALK;
            for (int c = 0; c < 2; ++c) {
                rc = st.mode == AES_MODE_REG ? B.add_reg_line(d.q[c], &st.ring[c][0])
                                             : B.add_walk(d.q[c], true, &st.ring[c][0]);
                if (rc) return rc;
            }
            st.fb = (float)d.p[0]; st.dry = (float)d.p[1]; st.wet = (float)d.p[2];
            const bool glob = p->ring[st.ring[0][0]].space == AES_SPACE_GLOBAL &&
                              p->ring[st.ring[1][0]].space == AES_SPACE_GLOBAL;
            if (st.mode == AES_MODE_REG && glob && d.q[0] >= 2 * T && d.q[1] >= 2 * T && p->pf_stage < 0 && !getenv("AES_NO_PF")) {
                st.pf = 1;
                p->pf_stage = s;
            }
            break;
        }
        case AES_STAGE_REVERB: {
            st.nc = (int)d.q[0]; st.na = (int)d.q[1];
            st.dry = (float)d.p[0]; st.wet = (float)d.p[1];
            const double h = d.p[2];
            if (!(h >= 0.0 && h < 1.0)) return B.fail("reverb: damp must be in [0,1)");
            st.h = (float)h; st.omh = (float)(1.0 - h); st.a = (float)d.p[3];
            st.nscan = 0;
            const double eps = ldexp(1.0, -32);
            for (int i = 0; i < 5; ++i) {
                const double v = pow(h, (double)FR * (double)(1 << i));
                st.hp[i] = (float)v;
                if (v >= eps) st.nscan = i + 1;
            }
            const double hw = pow(h, 32.0 * FR);
            st.hp[5] = (float)hw;
            st.nxw = 1;
            for (double v = hw; v >= eps && st.nxw < 8; v *= hw) ++st.nxw;
            for (int l = 0; l < 32; ++l) st.hlane[l] = (float)pow(h, (double)FR * l);
            if (d.q[2] < 0) return B.fail("reverb: negative pre-delay");
            // the pre-delay has no feedback: a line shorter than the tile needs no phase walk, only
            // its writes ahead of its reads (period lag + T either way, see add_reg_line)
            st.mode = d.q[2] >= T ? AES_MODE_REG : AES_MODE_REGB;
            for (int side = 0; side < 2; ++side) {
                if (d.q[2] > 0) {
                    if ((rc = B.add_reg_line(d.q[2], &st.pre_ring[side]))) return rc;
                }
                for (int c = 0; c < st.nc; ++c) {
                    const long long L = d.q[4 + 8 * side + c];
                    // reads and writes of a comb are separated by the scan barrier: period roundup4(L) suffices
                    B.next_prio = 0;
                    rc = B.add_ring(L, aes_roundup4(L), false, &st.ring[side][c]);
                    B.next_prio = 2;
                    if (rc) return rc;
                    st.g[side][c] = (float)d.p[4 + 8 * side + c];
                }
                for (int k = 0; k < st.na; ++k) {
                    B.next_prio = 1;
                    rc = B.add_walk(d.q[20 + 4 * side + k], true, &st.apring[side][k]);
                    B.next_prio = 2;
                    if (rc) return rc;
                }
            }
            break;
        }
        case AES_STAGE_BIQUAD: {
            for (int i = 0; i < 5; ++i) st.bq[i] = d.p[i];
            const double A[4] = { -d.p[3], -d.p[4], 1.0, 0.0 };
            for (int i = 0; i < 5; ++i) aes_mat2_pow(A, (long long)FR << i, st.bq_pow[i]);
            aes_mat2_pow(A, 32LL * FR, st.bq_pow[5]);
            for (int l = 0; l < 32; ++l) aes_mat2_pow(A, (long long)FR * l, st.bq_lane[l]);
            for (int i = 0; i < 8; ++i) st.init[i] = d.p[8 + i];
            break;
        }
        case AES_STAGE_GATE:
            st.thr = d.p[0]; st.att = d.p[1]; st.rel = d.p[2];
            st.init[0] = d.p[3];
            break;
        case AES_STAGE_OCTAVER: {
            if (d.q[0] < 4 || d.q[0] > (1 << 20)) return B.fail("octaver: ring size out of range");
            st.oct_size = (int)d.q[0];
            int R = 4;
            while (R < st.oct_size + T + 4) R <<= 1;
            st.oct_mask = R - 1;
            if ((rc = B.add_ring(R, R, false, &st.ring[0][0]))) return rc;
            st.ph0 = d.p[0]; st.step = d.p[1]; st.mix = (float)d.p[2];
            st.fsize = (double)st.oct_size;
            break;
        }
        case AES_STAGE_DISTORTION:
            st.drive = (float)d.p[0]; st.mix = (float)d.p[1];
            break;
        default:
            return B.fail("unknown stage kind");
        }
    }
    B.layout_smem();
    p->smem_floats = (int)B.smem_off;
    p->scratch_floats = B.glob_off > 0 ? B.glob_off : 32;
    if (aes_plan_smem_bytes(*p) > AES_SMEM_LIMIT) {
        snprintf(err, errlen, "chain needs %zu bytes of shared memory per clip (limit %d)",
                 aes_plan_smem_bytes(*p), AES_SMEM_LIMIT);
        return AES_ERR_UNSUPPORTED;
    }
    return 0;
}
