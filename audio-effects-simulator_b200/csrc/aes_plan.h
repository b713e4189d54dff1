// aes_plan.h -- device-side description of a fused effect chain (shared by the host
// plan compiler in aes_chain.cu and the kernel in aes_chain_kernel.cuh).
#pragma once
#include <stdint.h>

#define AES_NT 256            // threads per CTA of the chain kernel
#define AES_MAX_STAGES 16
#define AES_MAX_RINGS 64
#define AES_MAX_COMB 8
#define AES_MAX_AP 4

// Where a ring lives.  Small, hot rings (combs, all-passes, octaver) sit in shared
// memory; long ones (feedback delay, pre-delay) in a per-CTA global scratch that
// stays L2-resident (one CTA touches only its own few hundred KB).
enum { AES_SPACE_SMEM = 0, AES_SPACE_GLOBAL = 1 };

struct DevRing {
    int len;        // ring length == the lag it serves (slot of sample n is n mod len)
    int tinc;       // T mod len: per-tile advance of the slot of the tile's first sample
    int space;
    int pad;
    long long off;  // float offset in dynamic smem (after the tile buffers) or in the CTA's global scratch
};

struct DevStage {
    int kind;
    int nc, na;                 // reverb: combs / all-passes per side
    int pre_ring[2];            // reverb: pre-delay ring id per side, -1 when pre_dS == 0
    int ring[2][AES_MAX_COMB];  // delay: ring[c][0]; reverb: comb ring ids; octaver: ring[0][0]
    int apring[2][AES_MAX_AP];
    int nscan;                  // comb one-pole: warp-scan steps that still matter (h^(K*2^s) >= 2^-32)
    int state_off;              // offset (in doubles) of this stage's carried scalars in the smem state area
    float g[2][AES_MAX_COMB];   // comb feedback gains
    float dry, wet, h, omh, a, fb;
    float hp[6];                // h^(K*2^s) for s=0..4, hp[5] = h^(32K)
    float hlane[32];            // h^(K*lane)
    float drive, mix;
    // biquad (f64): coefficients, A^(K*2^s) (s=0..4), A^(32K), A^(K*lane); A = [[-a1,-a2],[1,0]]
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
    int K;                      // samples per thread chunk in the scan stages; tile T = 128*K frames
    int T;
    int smem_floats;            // ring area size in floats (after cur/aux)
    int n_state;                // doubles of carried scalar state
    long long scratch_floats;   // per-CTA global scratch
    DevRing ring[AES_MAX_RINGS];
    DevStage stage[AES_MAX_STAGES];
};

// Dynamic shared memory layout of the chain kernel (floats unless noted):
//   cur[2][T] | aux[2][T] | rings[smem_floats] | wtot[64 doubles] | state[n_state doubles x2 (ping-pong)] | rpos[n_rings ints]
static inline size_t aes_plan_smem_bytes(const DevPlan &p)
{
    size_t f = (size_t)4 * p.T + (size_t)p.smem_floats;
    f = (f + 1) & ~(size_t)1;                       // 8-byte align the double area
    size_t bytes = f * 4 + 64 * 8 + (size_t)2 * p.n_state * 8 + (size_t)p.n_rings * 4;
    return (bytes + 15) & ~(size_t)15;
}
