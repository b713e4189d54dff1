// aes_plan.h -- device-side description of a fused effect chain (shared by the host
// plan compiler in aes_plan_build.h and the kernel in aes_chain_kernel.cuh).
#pragma once
#include <stdint.h>

#define AES_NT 256            // threads per CTA of the chain kernel
#define AES_MAX_STAGES 16
#define AES_MAX_RINGS 64
#define AES_MAX_COMB 8
#define AES_MAX_AP 4
// Every reverb comb ring is followed by 4 spare floats in shared memory: the pipelined kernel
// (aes_rv_kernel.cuh) lays misaligned comb rings out in planes with one guard element per plane.
#define AES_COMB_RING_PAD 4

// Where a ring lives.  Small, hot rings (combs, all-passes, octaver) sit in shared
// memory; long ones (feedback delay, pre-delay) in a per-CTA global scratch that is
// meant to stay L2-resident (one CTA touches only its own few hundred KB).
enum { AES_SPACE_SMEM = 0, AES_SPACE_GLOBAL = 1 };

// How a delay line is walked.
//  WALK: ring length == lag L, slot(n) = n mod L; thread j walks samples j, j+L, ... of
//        the tile with the line value in a register (any lag >= 1; tile in smem).
//  REG : "aligned" ring for the register-resident tile (lag >= T): ring period `len` is a
//        multiple of 4, slot(n) = n mod len, so the 4 consecutive frames of a thread are
//        one aligned float4 on the write side; the read side at (n - lag) mod len is
//        misaligned by a per-ring constant and is served by two aligned float4 loads.
enum { AES_MODE_WALK = 0, AES_MODE_REG = 1, AES_MODE_REGB = 2 };   // REGB: REG ring for a PURE delay shorter than the tile: written first, read behind a barrier

struct DevRing {
    int len;        // ring period in floats
    int lag;        // delay served by this ring (== len for WALK rings)
    int tinc;       // T mod len: per-tile advance of the slot of the tile's first frame
    int space;
    long long off;  // float offset in the smem ring area or in the CTA's global scratch
};

struct DevStage {
    int kind;
    int nc, na;                 // reverb: combs / all-passes per side
    int mode;                   // delay: AES_MODE_*; reverb: mode of the pre-delay line
    int pre_ring[2];            // reverb: pre-delay ring id per side, -1 when pre_dS == 0
    int ring[2][AES_MAX_COMB];  // delay: ring[c][0]; reverb: comb ring ids; octaver: ring[0][0]
    int apring[2][AES_MAX_AP];
    int nscan;                  // comb one-pole: warp-scan steps that still matter (h^(FR*2^s) >= 2^-32)
    int nxw;                    // comb one-pole: preceding warps whose carry still matters
    int pf;                     // delay REG/global: 1 = ring reads are prefetched one tile ahead (lag >= 2T)
    float g[2][AES_MAX_COMB];   // comb feedback gains
    float dry, wet, h, omh, a, fb;
    float hp[6];                // h^(FR*2^s) for s=0..4, hp[5] = h^(32*FR)
    float hlane[32];            // h^(FR*lane)
    float drive, mix;
    // biquad (f64): coefficients, A^(FR*2^s) (s=0..4), A^(32*FR), A^(FR*lane); A = [[-a1,-a2],[1,0]]
    double bq[5];
    double bq_pow[6][4];
    double bq_lane[32][4];
    double init[16];            // initial carried scalars (biquad x1,x2,y1,y2 per channel; gate gain)
    // gate (f64)
    double thr, att, rel;
    // octaver
    double ph0, step, fsize;
    int oct_size, oct_mask;
};

struct DevPlan {
    int n_stages;
    int n_rings;
    int FR;                     // frames per thread; tile T = 256*FR frames
    int T;
    int smem_floats;            // ring area size in floats (after the tile buffer)
    int n_state;                // doubles of carried scalar state (16 per stage)
    int pf_stage;               // stage whose global delay ring is prefetched a tile ahead, or -1
    int pad;
    long long scratch_floats;   // per-CTA global scratch
    DevRing ring[AES_MAX_RINGS];
    DevStage stage[AES_MAX_STAGES];
};

// Dynamic shared memory layout of the chain kernel:
//   tile[2][T] f32 | rings[smem_floats] f32 | wtot[64] f64 | state[2][n_state] f64 | rpos[2][n_rings] i32
static inline size_t aes_plan_smem_bytes(const DevPlan &p)
{
    size_t f = (size_t)2 * p.T + (size_t)p.smem_floats;
    f = (f + 3) & ~(size_t)3;
    size_t bytes = f * 4 + 64 * 8 + (size_t)2 * p.n_state * 8 + (size_t)2 * p.n_rings * 4;
    return (bytes + 15) & ~(size_t)15;
}
