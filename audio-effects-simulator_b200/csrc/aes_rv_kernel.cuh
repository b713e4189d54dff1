// aes_rv_kernel.cuh -- software-pipelined kernel for chains that END in the reference's default
// reverb topology (reverb.py:74-81,158-177 at 48 kHz / 44.1 kHz): Rain Delay, Cathedral, Guitar
// Filter, the bare reverb.
//
// Why a second kernel: ncu on aes_fast_kernel (profiles/r1o) showed the reverb chains issue-bound
// with 47 % of the executed instructions integer / control / uniform bookkeeping, and five CTA
// barriers per tile that put the latency-bound all-pass walks strictly behind the comb bank.
// Here
//   * the tile loop is software-pipelined:  the all-pass walks, the dry/wet mix and the store of
//     tile i-1 run in the same barrier phases as the delay / comb work of tile i (two tile buffers
//     alternate), so a tile costs TWO CTA barriers instead of five and warps that have no
//     all-pass column to walk go straight on to comb work;
//
//         phase 1:  all-pass 1 of tile i-1   |  input, feedback delay / biquad pass 1, comb ring
//                                            |  reads, local one-pole ends, warp scan        -- barrier
//         phase 2:  all-pass 2 of tile i-1   |  carries, comb update + ring writes, comb sum -- barrier
//         phase 3:  mix + store of tile i-1, slots advance, loads for tile i+1 issued
//
//   * the next tile's input frames and feedback-delay line samples are fetched with plain 128-bit
//     loads issued in phase 3 (registers are slack there) instead of a TMA staging round trip
//     through shared memory: the shared-memory data pipe was 56 % busy and is the second limiter;
//   * everything about the reverb is a compile-time constant (TOPO), per-lane constants are loaded
//     once per kernel, carried values move as float4, all-pass ring phases live in registers.
// Arithmetic per sample is exactly that of aes_fast_kernel (same parity tests).  Any other chain
// shape / sample rate / comb set keeps using aes_fast_kernel or the interpreter.
#pragma once
#include <type_traits>
#include "aes_fast_kernel.cuh"

// compile-time loop: f(std::integral_constant<int, I>) for I = 0 .. N-1 (ring lengths are template arguments)
template <int I, int N, class F>
__device__ __forceinline__ void aes_static_for(F &&f)
{
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        aes_static_for<I + 1, N>(f);
    }
}

#define AESRV_PRE_NONE 0
#define AESRV_PRE_DELAY 1           // prefetched feedback delay ahead of the reverb (lag >= 2T + 4, global lines)
#define AESRV_PRE_BIQUAD 2          // one biquad ahead of the reverb

// CTA: 8 "comb" warps (threads 0..255: frames in registers, everything but the all-passes) and
// 4 "walker" warps (threads 256..383: the two all-pass stages of tile i-1 -- two warps per channel --
// and the TMA copies of tile i+2).  The register file is split with setmaxnreg: 104 + 32 registers
// (384 threads x 80 at launch = 256 x 104 + 128 x 32).  Measured alternative (ncu r2j): eight walker
// warps with 24 registers take the walkers off the critical path (comb warps wait 3.6 % instead of 10 %)
// but the kernel as a whole gets slower, 268 against 293 Gsamples/s -- issue slots and the shared-memory
// pipe are the budget, and twice the walker warps spend more of both.  Also measured: walks vectorised
// over adjacent columns (64- / 128-bit accesses for even ring lengths, r2k): fewer instructions, but the
// work concentrates in a quarter of the threads and the walkers' critical path grows, 271;  a 96 + 48
// register split with four interleaved columns per walker (r2l): the walkers get 30 % faster and the bare
// reverb gains 1 % (307), but the feedback-delay shape spills in the comb warps, 265;  four columns in 32
// registers (r2m): 286.  The L1 data pipe is the binding resource (ncu r2i: l1tex data-pipe wavefronts 88 % of peak;
// 1781 wavefronts per tile against ~1024 for rings + I/O alone, the rest are the tile hand-offs between the
// warp roles and the TMA's own shared-memory writes, ~229 per tile for 16.4 KB).  Two placements of the staged line
// samples were measured against the plain 16-byte aligned one (r2t, r2u): at the 128-byte phase of their global
// source, and 128-byte aligned destinations; both RAISED the copy's shared-memory wavefronts (288 per tile) and cost 3 %.
// Also measured (r2aa): the staging as per-thread cp.async (LDGSTS, 16 bytes per copy) issued by the comb warps two
// tiles ahead, each thread waiting for its own copies ahead of the comb-wide barrier -- no mbarrier, no issuing
// thread.  268 against 289 Gsamples/s: the copies run through the LSU (925 M against 871 M LSU wavefronts, bank
// conflicts 144 M against 101 M: 16-byte pieces at a 32-byte lane stride) and add 4 % to the comb warps' instructions,
// while the TMA's writes bypass the LSU queue altogether.
#define AESRV_NT 384
#define AESRV_NW 128
#define AESRV_NWC 64                // walkers per channel
#define AESRV_REGS_COMB 104
#define AESRV_REGS_WALK 32
#define AESRV_NXS 4                 // tile buffers (8 KB each): TMA input -> comb sum -> all-pass output
// named barriers (0 is __syncthreads over all 384 threads)
#define AESRV_BAR_COMB 1            // the comb warps between their two phases (256 threads)
#define AESRV_BAR_FULL 2            // +parity: comb sum of a tile is in its buffer (comb warps arrive, walkers wait)
#define AESRV_BAR_DONE 4            // +parity: all-pass output of a tile is in its buffer (walkers arrive, comb warps wait)
#define AESRV_BAR_CHAN 6            // +channel: the walkers of one channel between the two all-passes

// shared memory (floats): XS[4][2][T] | rings[smem_floats] | wt[8][8] | cst[2][8] | f64: wtot[32] | bst[2][8] |
//                         LN[2][2][T+8] (feedback-delay shapes only) | 4 mbarriers
__host__ __device__ inline size_t aes_rv_smem_bytes(int smem_floats, int pre)
{
    const size_t T = AES_NT * 4;
    size_t f = (2 * AESRV_NXS * T + (size_t)smem_floats + 3) & ~(size_t)3;
    return (f + 64 + 16) * 4 + (32 + 16) * 8 + (pre == AESRV_PRE_DELAY ? 4 * (T + 8) * 4 : 0) + AESRV_NXS * 8 + 16;
}

// ---- comb rings: bank-conflict-free layouts ------------------------------------------------------
// A thread reads the four samples L behind its frames and writes four new ones; both as one float4
// only if L is a multiple of 4.  In a linear ring the misaligned reads cost twice the wavefronts
// (ptxas narrows the second vector to scalar / 64-bit loads that conflict 4- / 2-way; ncu r2e: 273 of
// 1765 shared-memory wavefronts per tile were conflicts, the data pipe 80 % busy).  So the layout of a
// ring follows its misalignment m = roundup4(L) - L, all at compile time:
//   m == 0  LIN : linear; float4 slot u at byte 16u;                        LDS.128 / STS.128
//   m == 2  P2  : two planes of 8-byte units, unit k -> plane k&1, index k>>1; LDS.64 x2 / STS.64 x2
//   m odd   P4  : four planes of floats, element q -> plane q&3, index q>>2;   LDS.32 x4 / STS.32 x4
// Every plane ends in one guard entry that mirrors its entry 0 (written with it), so a read that runs
// one slot past the thread's own never needs a wrap.  The slot register holds u in the layout's own
// bytes (16u / 8u / 4u) and advances by one add-min per tile.
template <int L> struct AesrvComb {
    static constexpr int P = (L + 3) & ~3, M = P - L, NSLOT = P / 4;
    static constexpr int KIND = M == 0 ? 0 : (M == 2 ? 2 : 4);
    static constexpr int SB = KIND == 0 ? 16 : (KIND == 2 ? 8 : 4);        // slot bytes
    static constexpr int WRAP = NSLOT * SB;
    static constexpr int TINC = ((AES_NT * 4 / 4) % NSLOT) * SB;
    static constexpr int PS = (NSLOT + 1) * SB;                            // plane stride (P2 / P4), guard included
};

// y[j] = ring sample L behind frame j of the thread (slot register s, ring base rb in bytes)
template <int L>
__device__ __forceinline__ void aesrv_comb_read(const char *rb, int s, float (&y)[4])
{
    using C = AesrvComb<L>;
    if constexpr (C::KIND == 0) {
        const float4 A = aes_lds_v4(reinterpret_cast<const float *>(rb + s));
        y[0] = A.x; y[1] = A.y; y[2] = A.z; y[3] = A.w;
    } else if constexpr (C::KIND == 2) {
        const float2 a = *reinterpret_cast<const float2 *>(rb + C::PS + s);        // unit 2u+1: plane 1, index u
        const float2 b = *reinterpret_cast<const float2 *>(rb + s + 8);            // unit 2u+2: plane 0, index u+1
        y[0] = a.x; y[1] = a.y; y[2] = b.x; y[3] = b.y;
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            constexpr int M = C::M;
            const int e = M + j;                            // element 4u + e: plane e&3, index u + (e>>2)
            y[j] = *reinterpret_cast<const float *>(rb + (e & 3) * C::PS + s + 4 * (e >> 2));
        }
    }
}

template <int L>
__device__ __forceinline__ void aesrv_comb_write(char *rb, int s, const float (&nb)[4])
{
    using C = AesrvComb<L>;
    if constexpr (C::KIND == 0) {
        *reinterpret_cast<float4 *>(rb + s) = make_float4(nb[0], nb[1], nb[2], nb[3]);
    } else if constexpr (C::KIND == 2) {
        *reinterpret_cast<float2 *>(rb + s) = make_float2(nb[0], nb[1]);
        *reinterpret_cast<float2 *>(rb + C::PS + s) = make_float2(nb[2], nb[3]);
        if (s == 0) *reinterpret_cast<float2 *>(rb + C::NSLOT * 8) = make_float2(nb[0], nb[1]);    // guard of plane 0
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) *reinterpret_cast<float *>(rb + j * C::PS + s) = nb[j];
        if (s == 0) {
#pragma unroll
            for (int j = 0; j < C::M; ++j) *reinterpret_cast<float *>(rb + j * C::PS + C::NSLOT * 4) = nb[j];
        }
    }
}

// all-pass walk over one channel's tile `s` (T samples) by NTH threads: column j of the tile, seen as
// rows of L samples, is one serial chain  y = line - g*x,  line' = x + g*y  (reverb.py:48-67) down the
// rows; the chain's state enters from / leaves to the ring (length L, phase `pos` = tile start mod L).
// A thread walks NC columns (j0 + NTH*(r0 + c)) at once: the chains are independent, and two FFMA
// chains in flight is what the walkers' 32 registers allow (a single chain left them latency-bound at
// 0.08 instructions per cycle, ncu r2d).
template <int L, int NTH, int R0, int NCOL>
__device__ __forceinline__ void aesrv_walk_cols(float *s, float *rb, int pos, int j0, float g)
{
    constexpr int T = AES_NT * 4;
    constexpr int KMAX = (T + L - 1) / L;                   // rows that can hold a column
    static_assert(L >= NTH, "an idle lane shadows column j0 < NTH, which has to exist");
    int slot[NCOL];
    float line[NCOL];
    float *p[NCOL];
    bool act[NCOL];
#pragma unroll
    for (int c = 0; c < NCOL; ++c) {
        const int j = j0 + NTH * (R0 + c);
        act[c] = (NTH * (R0 + c + 1) <= L) || j < L;        // compile-time true when the whole trip fits
        // idle lanes shadow their own trip-0 column without storing: its bank, (j0 + k*L) mod 32, is the one the
        // lane would have hit anyway.  (Shadowing column 0 put every idle lane on lane 0's bank at another address:
        // ncu r2ad showed 2.0 wavefronts per load in the partial trips, 34 M excess wavefronts per launch.)
        const int jj = act[c] ? j : j0;
        slot[c] = pos + jj;
        if (slot[c] >= L) slot[c] -= L;
        line[c] = rb[slot[c]];
        p[c] = s + jj;
    }
#pragma unroll
    for (int k0 = 0; k0 < KMAX; k0 += 4) {                  // loads of four rows ahead of their serial chains
        float xs[NCOL][4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int c = 0; c < NCOL; ++c) {
                const int k = k0 + u;
                if (k < KMAX) {
                    if ((L - 1) + k * L < T) xs[c][u] = p[c][k * L];                 // row k is complete
                    else xs[c][u] = (p[c] - s) + k * L < T ? p[c][k * L] : 0.0f;
                }
            }
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int c = 0; c < NCOL; ++c) {
                const int k = k0 + u;
                if (k < KMAX) {
                    // (carrying line' = g*line + (1-g*g)*x -- one dependent FFMA per row instead of two -- changed
                    // nothing: the walkers wait on shared-memory loads, not on the FFMA chain; ncu r2i)
                    const float yo = fmaf(-g, xs[c][u], line[c]);
                    if ((L - 1) + k * L < T) {              // complete row: only idle lanes must not store
                        if (act[c]) p[c][k * L] = yo;
                        line[c] = fmaf(g, yo, xs[c][u]);
                    } else if ((p[c] - s) + k * L < T) {    // last row: the column may end above it
                        if (act[c]) p[c][k * L] = yo;
                        line[c] = fmaf(g, yo, xs[c][u]);
                    }
                }
            }
    }
#pragma unroll
    for (int c = 0; c < NCOL; ++c)
        if (act[c]) rb[slot[c]] = line[c];
}

// columns j0, j0 + NTH, ... of a ring of L samples, two at a time; a warp whose lanes all lie beyond the
// last, partial trip walks its remaining column alone
template <int L, int NTH>
__device__ __forceinline__ void aesrv_walk(float *s, float *rb, int pos, int j0, float g)
{
    constexpr int NTRIP = (L + NTH - 1) / NTH;
    aes_static_for<0, NTRIP / 2>([&](auto ir) {
        constexpr int r0 = 2 * decltype(ir)::value;
        constexpr bool second_partial = NTH * (r0 + 2) > L;
        if (!second_partial || (j0 & ~31) + NTH * (r0 + 1) < L) aesrv_walk_cols<L, NTH, r0, 2>(s, rb, pos, j0, g);
        else aesrv_walk_cols<L, NTH, r0, 1>(s, rb, pos, j0, g);
    });
    if constexpr ((NTRIP & 1) != 0) {
        if ((j0 & ~31) + NTH * (NTRIP - 1) < L) aesrv_walk_cols<L, NTH, NTRIP - 1, 1>(s, rb, pos, j0, g);
    }
}

template <int TOPO, int PRE, int PM>
__device__ void aes_rv_body(const FastArgs &a)
{
    constexpr int FR = 4, T = AES_NT * FR, NC = 4;
    constexpr int SR = PRE == AESRV_PRE_NONE ? 0 : 1;      // index of the reverb stage
    constexpr int LNS = T + 8;                              // staged line samples per channel (T + 4 used)
    static_assert(TOPO == AESF_TOPO_48K || TOPO == AESF_TOPO_44K, "compile-time reverb topology only");
    static_assert(!(PRE == AESRV_PRE_BIQUAD && PM != 0), "biquad + pre-delay is not instantiated (the biquad output only exists in phase 2)");
    AES_DYN_SMEM(float, smem);
    const int tid = threadIdx.x;
    float *const rings = smem + 2 * AESRV_NXS * T;
    const int foff = (2 * AESRV_NXS * T + a.smem_floats + 3) & ~3;
    float *const wt = smem + foff;                          // [8 warps][2 ch][4 combs] warp totals of the comb one-poles
    float *const cst = wt + 64;                             // [2 parity][8] one-pole values carried across tiles
    double *const wtot = reinterpret_cast<double *>(cst + 16);  // biquad: [8 warps][2 ch][2]
    double *const bst = wtot + 32;                          // biquad: [2 parity][8] DF-I state carried across tiles
    float *const LN = reinterpret_cast<float *>(bst + 16);  // [2 parity][2 ch][LNS] staged feedback-delay line samples
    unsigned long long *const mbar = reinterpret_cast<unsigned long long *>(LN + (PRE == AESRV_PRE_DELAY ? 4 * LNS : 0));
    float *const gscr = a.scratch + (long long)blockIdx.x * a.scratch_floats;
    const FastStage &rs = a.st[SR];
    const int N = (int)a.N;                                 // the host routes clips of 2^31 frames or more elsewhere
    const int ntiles = (N + T - 1) / T;

    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < AESRV_NXS; ++i) aes_mbar_init(mbar + i, 1);
        aes_mbar_init_fence();
    }
    // (the first __syncthreads of the clip loop orders the init before any use)

    if (tid >= AES_NT) {
        // =========================== walker warps ===========================
        aes_setmaxnreg_dec<AESRV_REGS_WALK>();
        const int wt_id = tid - AES_NT;                     // 0..127
        const int ch = wt_id / AESRV_NWC, j0 = wt_id % AESRV_NWC;
        const bool issuer = wt_id == 32;                    // a channel's second warp walks one all-pass-2 column, its first
                                                            // two: the TMA bookkeeping goes to a warp with time to spare
        const float apg = rs.a;
        constexpr int L1_0 = aesf_topo_ap(TOPO, 0, 0), L1_1 = aesf_topo_ap(TOPO, 1, 0);
        constexpr int L2_0 = aesf_topo_ap(TOPO, 0, 1), L2_1 = aesf_topo_ap(TOPO, 1, 1);
        const int apL1 = ch ? L1_1 : L1_0, apI1 = ch ? T % L1_1 : T % L1_0;
        const int apL2 = ch ? L2_1 : L2_0, apI2 = ch ? T % L2_1 : T % L2_0;
        float *const ring1 = rings + (ch ? aesf_topo_ap_off(TOPO, 1, 0) : aesf_topo_ap_off(TOPO, 0, 0));
        float *const ring2 = rings + (ch ? aesf_topo_ap_off(TOPO, 1, 1) : aesf_topo_ap_off(TOPO, 0, 1));
        int q = 0;                                          // buffer of the tile being walked; never reset
        for (long long b = blockIdx.x; b < a.B; b += gridDim.x) {
            {   // fresh all-pass lines (the comb warps clear theirs): the all-pass rings follow the comb rings
                constexpr int ap0 = aesf_topo_ap_off(TOPO, 0, 0);
                for (int i = ap0 + wt_id; i < a.smem_floats; i += AESRV_NW) rings[i] = 0.0f;
            }
            __syncthreads();
            const bool clip_staged = a.in_fmt == AESK_F32_STEREO && ((b * a.N) & 1) == 0;
            const float *const xin = reinterpret_cast<const float *>(a.x) + 2 * (b * a.N);
            int lnbase[2] = { 0, 0 };                       // thread 0's aligned line read base of the tile staged next
            if constexpr (PRE == AESRV_PRE_DELAY) {
#pragma unroll
                for (int c2 = 0; c2 < 2; ++c2) {
                    int w0;
                    aesf_line_init<FR>(a.st[0].ring[c2][0], 0, w0, lnbase[c2]);
                }
            }
            // hand the tile starting at frame f0 to the TMA: input frames into buffer qq, line samples into LN[pp]
            auto issue_tile = [&](int f0, int qq, int pp) {
                unsigned long long *bar = mbar + qq;
                // generic-proxy accesses (line stores; loads / stores of the tile buffer) before the copies
                if (PRE == AESRV_PRE_DELAY) aes_fence_proxy_async();
                aes_fence_proxy_async_smem();
                aes_mbar_expect(bar, (unsigned)(T * 8 + (PRE == AESRV_PRE_DELAY ? 2 * (T + 4) * 4 : 0)));
                aes_bulk_g2s(smem + qq * 2 * T, xin + 2 * (long long)f0, T * 8, bar, aes_policy_evict_first());
                if constexpr (PRE == AESRV_PRE_DELAY) {
                    const unsigned long long keep = aes_policy_evict_last();
#pragma unroll
                    for (int c2 = 0; c2 < 2; ++c2) {
                        const FRing rg = a.st[0].ring[c2][0];
                        const int a0 = lnbase[c2];
                        const float *rb = gscr + rg.off;
                        float *dst = LN + (pp * 2 + c2) * LNS;
                        const int n1 = (rg.len - a0) < (T + 4) ? (rg.len - a0) : (T + 4);
                        aes_bulk_g2s(dst, rb + a0, (unsigned)n1 * 4, bar, keep);
                        if (n1 < T + 4) aes_bulk_g2s(dst + n1, rb, (unsigned)(T + 4 - n1) * 4, bar, keep);   // the ring wraps
                        lnbase[c2] = aesf_adv(a0, rg.tinc, rg.len);
                    }
                }
#ifdef AES_CPU_EMU
                aes_mbar_complete_emu(bar);
#endif
            };
            // tiles 0 and 1 right away; tile i+2 once the comb warps are through with tile i (its buffer, tile
            // i-2's, was read last before that, and every line sample the copy needs has been stored)
            if (issuer && clip_staged) {
                if (N >= T) issue_tile(0, q, 0);
                if (N >= 2 * T) issue_tile(T, (q + 1) & (AESRV_NXS - 1), 1);
            }
            int pos1 = 0, pos2 = 0;                         // all-pass ring phases of the tile being walked
            for (int it = 0; it < ntiles; ++it) {
                const int par = it & 1;
                float *const S = smem + q * 2 * T + ch * T;
                aes_bar_sync(AESRV_BAR_FULL + par, AESRV_NT);
                if (issuer && clip_staged && N - it * T >= 3 * T)
                    issue_tile((it + 2) * T, (q + 2) & (AESRV_NXS - 1), par);
                if (ch == 0) aesrv_walk<L1_0, AESRV_NWC>(S, ring1, pos1, j0, apg);
                else         aesrv_walk<L1_1, AESRV_NWC>(S, ring1, pos1, j0, apg);
                aes_bar_sync(AESRV_BAR_CHAN + ch, AESRV_NWC);
                if (ch == 0) aesrv_walk<L2_0, AESRV_NWC>(S, ring2, pos2, j0, apg);
                else         aesrv_walk<L2_1, AESRV_NWC>(S, ring2, pos2, j0, apg);
                aes_bar_arrive(AESRV_BAR_DONE + par, AESRV_NT);
                pos1 = aesf_adv(pos1, apI1, apL1);
                pos2 = aesf_adv(pos2, apI2, apL2);
                q = (q + 1) & (AESRV_NXS - 1);
            }
            __syncthreads();                                // (matches the comb warps' clip-end barrier)
        }
        return;
    }

    // =========================== comb warps ===========================
    aes_setmaxnreg_inc<AESRV_REGS_COMB>();
    const int lane = tid & 31, warp = tid >> 5, i0 = FR * tid;
    // feedback-delay lines of this CTA.  Kept in registers on purpose: left alone, ptxas re-derives
    // blockIdx * scratch_floats + offset in 64 bits in front of every line store (ten instructions each).
    float *gl0 = gscr + (PRE == AESRV_PRE_DELAY ? a.st[0].ring[0][0].off : 0);
    float *gl1 = gscr + (PRE == AESRV_PRE_DELAY ? a.st[0].ring[1][0].off : 0);
#ifndef AES_CPU_EMU
    asm volatile("" : "+l"(gl0), "+l"(gl1));
#endif
    const float h = rs.h, hw = rs.hp[5], rdry = rs.dry, rwet = rs.wet;
    const int nscan = rs.nscan, nxw = rs.nxw;
    const float hl = a.lane_tab[(SR * 32 + lane) * FAST_LANE_STRIDE];       // h^(FR*lane)
    // where this warp finds the one-pole values entering its first frame: the previous warp's totals,
    // or (warp 0) the values carried from the previous tile, by tile parity
    const float *const cp0 = warp == 0 ? cst : wt + (warp - 1) * 8;
    const float *const cp1 = warp == 0 ? cst + 8 : wt + (warp - 1) * 8;

    ChainArgs io;                                           // ragged / odd-format tile I/O is shared with the generic kernel
    io.x = a.x; io.y = a.y; io.N = a.N; io.in_fmt = a.in_fmt; io.out_fmt = a.out_fmt;
    int q = 0;                                              // tile buffer of the current tile; never reset
    unsigned phbits = 0;                                    // parity of the next wait on each buffer's mbarrier

    for (long long b = blockIdx.x; b < a.B; b += gridDim.x) {
        {   // fresh lines and carried values at every clip start (core.py:123-129 re-prepares)
            constexpr int ap0 = aesf_topo_ap_off(TOPO, 0, 0);           // comb rings come first (aes_plan_build.h)
            float4 *r4 = reinterpret_cast<float4 *>(rings);
            const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int i = tid; i < ap0 / 4; i += AES_NT) r4[i] = z4;
            if (tid < 16) cst[tid] = 0.0f;
            if (PRE == AESRV_PRE_BIQUAD && tid < 8) { bst[tid] = a.init[0][tid]; bst[8 + tid] = a.init[0][tid]; }
        }
        int wb[2][NC];                                      // comb slots: this thread's float4 slot, in its ring's slot bytes
        aes_static_for<0, 2 * NC>([&](auto ic) {
            constexpr int ch = decltype(ic)::value / NC, cc = decltype(ic)::value % NC;
            wb[ch][cc] = tid * AesrvComb<aesf_topo_comb(TOPO, ch, cc)>::SB;
        });
        int dw[2] = { 0, 0 };                               // feedback-delay line: write slot of this thread's frames
        if constexpr (PRE == AESRV_PRE_DELAY) {
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
                int a0;
                aesf_line_init<FR>(a.st[0].ring[ch][0], i0, dw[ch], a0);
            }
        }
        // reverb pre-delay line: write slot of this thread's frames.  One slot serves both channels (their lines have
        // one lag and one period) and the read slot is derived at its use: the Cathedral shape spilled 32 bytes into
        // the tile loop with the four slots and both channels' delayed frames live at once (ncu-less evidence: the
        // same reverb without a pre-delay ran at 324 against 248 Gsamples/s, profiles/tools/time_reverb_predelay.py).
        int pw = 0;
        if constexpr (PM != 0) {
            int a0;
            aesf_line_init<FR>(rs.pre[0], i0, pw, a0);
        }
        __syncthreads();

        // a tile goes through TMA staging when it is full, 16-byte aligned and plain f32 stereo
        const bool clip_staged = a.in_fmt == AESK_F32_STEREO && ((b * a.N) & 1) == 0;
        const bool fast_out = a.out_fmt == AESK_F32_STEREO && ((b * a.N) & 1) == 0;
        float *yp = reinterpret_cast<float *>(a.y) + 2 * (b * a.N) + 2 * i0 - 2 * T;      // this thread's frames of tile i-1
        float vprev[2][FR];                                 // dry signal at the reverb input, tile i-1
#pragma unroll
        for (int ch = 0; ch < 2; ++ch)
#pragma unroll
            for (int j = 0; j < FR; ++j) vprev[ch][j] = 0.0f;

        int par = 0;
        // One iteration: phases 1 and 2 of tile `it`, phase 3 of tile it-1.  FAST is the steady state -- a full,
        // TMA-staged tile with a full tile behind it, past the delay's first lap, f32 stereo out, the usual
        // one-warp carry reach -- where every flag below is a compile-time constant; the general
        // instantiation serves the first tiles, the ragged tail, the drain iteration and odd formats.
        auto tile_iter = [&](auto fast_tag, const int it) {
            constexpr bool FAST = decltype(fast_tag)::value;
            const bool has_cur = FAST || it < ntiles, has_prev = FAST || it > 0;
            const int n0 = it * T;
            const int rem = N - n0;
            const int len = FAST ? T : (rem < T ? rem : T);
            float *const Sc = smem + q * 2 * T;             // this tile's buffer: staged input, then the comb sum
            float *const Sp = smem + ((q + AESRV_NXS - 1) & (AESRV_NXS - 1)) * 2 * T;    // tile i-1's: its wet signal
            const bool staged = FAST || (clip_staged && len == T);
            float v[2][FR];

            if (has_cur) {
                float y[2][NC][FR], e[2][NC];
                [[maybe_unused]] double bq_e1[2], bq_e2[2];
                // ---------------- phase 1 ----------------
                if (staged) {
                    aes_mbar_wait_parity(mbar + q, (phbits >> q) & 1u);
                    phbits ^= 1u << q;
                    // interleaved L R L R ..., 32 bytes per thread: two 128-bit loads at a 32-byte lane stride
                    // conflict 2-way, so lanes 4..7 of every eight take their halves in the other order
                    const float *sx = Sc + 2 * i0;
                    const int sw = (tid >> 2) & 1;
                    const float4 ta = aes_lds_v4(sx + 4 * sw), tb = aes_lds_v4(sx + 4 * (sw ^ 1));
                    const float4 t0 = sw ? tb : ta, t1 = sw ? ta : tb;
                    v[0][0] = t0.x; v[1][0] = t0.y; v[0][1] = t0.z; v[1][1] = t0.w;
                    v[0][2] = t1.x; v[1][2] = t1.y; v[0][3] = t1.z; v[1][3] = t1.w;
                } else {
                    aes_load_frames<FR>(io, b, n0, len, tid, v);            // ragged / unaligned / non-f32 tiles
                }
                if constexpr (PRE == AESRV_PRE_DELAY) {
                    // feedback delay (delay.py:7-22,94-96); every delayed sample predates the tile
                    const FastStage &ds = a.st[0];
                    const float fb = ds.fb, dry = ds.dry, wet = ds.wet;
#pragma unroll
                    for (int ch = 0; ch < 2; ++ch) {
                        const FRing rg = ds.ring[ch][0];
                        const int m = ((rg.lag + 3) & ~3) - rg.lag;
                        float line[FR];
                        if (staged) {
                            // line samples staged by TMA: [i0 + m, i0 + m + FR) of the staged span
                            // (unaligned: the two aligned vectors around them; scalar loads at a 16-byte lane stride
                            // would conflict 4-way)
                            const float *sp = LN + (par * 2 + ch) * LNS + i0;
                            const float4 A = aes_lds_v4(sp);
                            if (m == 0) {
                                line[0] = A.x; line[1] = A.y; line[2] = A.z; line[3] = A.w;
                            } else {
                                aesf_select4<FR>(A, aes_lds_v4(sp + 4), m, line);
                            }
                        } else {
                            int a0 = dw[ch] - ((rg.lag + 3) & ~3);          // aligned read base: lag4 behind the write slot
                            if (a0 < 0) a0 += rg.len;
                            aesf_read<FR, 1>(ch ? gl1 : gl0, a0, m, rg.len, line);
                        }
                        if (!FAST && n0 < rg.lag) {                         // only the first tiles of a clip: zero history
#pragma unroll
                            for (int j = 0; j < FR; ++j)
                                if (n0 + i0 + j < rg.lag) line[j] = 0.0f;
                        }
                        float nb[FR];
#pragma unroll
                        for (int j = 0; j < FR; j += 2) {                   // two frames per packed instruction
                            const float2 x = make_float2(v[ch][j], v[ch][j + 1]), ln = make_float2(line[j], line[j + 1]);
                            const float2 n2 = aes_fma2(ln, make_float2(fb, fb), x);
                            nb[j] = n2.x; nb[j + 1] = n2.y;
                            aes_mix_clip2(dry, x, wet, ln, v[ch][j], v[ch][j + 1]);
                        }
                        aes_stg_v4((ch ? gl1 : gl0) + dw[ch], nb);
                    }
                }
                if constexpr (PRE == AESRV_PRE_BIQUAD) {
                    // transposed DF-II from zero state: only the chunk's end state is kept (filter.py:8-40; see aes_fast_kernel.cuh)
                    const FastStage &bs = a.st[0];
                    const double b0 = bs.bq[0], b1 = bs.bq[1], b2 = bs.bq[2], a1 = bs.bq[3], a2 = bs.bq[4];
#pragma unroll
                    for (int ch = 0; ch < 2; ++ch) {
                        double s1 = 0.0, s2 = 0.0;
#pragma unroll
                        for (int j = 0; j < FR; ++j) {
                            const double xj = (double)v[ch][j];
                            const double yy = fma(b0, xj, s1);
                            s1 = fma(b1, xj, fma(-a1, yy, s2));
                            s2 = fma(b2, xj, -a2 * yy);
                        }
                        bq_e1[ch] = s1; bq_e2[ch] = s2;             // (outputs are recomputed in phase 2: registers)
                    }
                    const int bq_nscan = bs.nscan;
                    for (int s = 0; s < bq_nscan; ++s) {
                        const double m0 = bs.bq_pow[s][0], m1 = bs.bq_pow[s][1], m2 = bs.bq_pow[s][2], m3 = bs.bq_pow[s][3];
#pragma unroll
                        for (int ch = 0; ch < 2; ++ch) {
                            const double u1 = __shfl_up_sync(0xffffffffu, bq_e1[ch], 1 << s);
                            const double u2 = __shfl_up_sync(0xffffffffu, bq_e2[ch], 1 << s);
                            if (lane >= (1 << s)) {
                                bq_e1[ch] = fma(m0, u1, fma(m2, u2, bq_e1[ch]));
                                bq_e2[ch] = fma(m1, u1, fma(m3, u2, bq_e2[ch]));
                            }
                        }
                    }
                    if (lane == 31) {
#pragma unroll
                        for (int ch = 0; ch < 2; ++ch) {
                            wtot[(warp * 2 + ch) * 2] = bq_e1[ch];
                            wtot[(warp * 2 + ch) * 2 + 1] = bq_e2[ch];
                        }
                    }
                }
                // damped combs (reverb.py:33-46) on u = lp/(1-h): delayed ring reads, one-pole over the
                // thread's 4 samples from zero, Kogge-Stone over the warp, lane 31 publishes the warp total
                aes_static_for<0, 2 * NC>([&](auto ic) {
                    constexpr int ch = decltype(ic)::value / NC, cc = decltype(ic)::value % NC;
                    aesrv_comb_read<aesf_topo_comb(TOPO, ch, cc)>(
                        reinterpret_cast<const char *>(rings + aesf_topo_comb_off(TOPO, ch, cc)), wb[ch][cc], y[ch][cc]);
                    float u = y[ch][cc][0];
#pragma unroll
                    for (int j = 1; j < FR; ++j) u = fmaf(h, u, y[ch][cc][j]);
                    e[ch][cc] = u;
                });
#pragma unroll 1
                for (int s = 0; s < nscan; ++s) {
                    const float m = rs.hp[s];
#pragma unroll
                    for (int ch = 0; ch < 2; ++ch)
#pragma unroll
                        for (int cc = 0; cc < NC; ++cc) {
                            const float t = __shfl_up_sync(0xffffffffu, e[ch][cc], 1 << s);
                            if (lane >= (1 << s)) e[ch][cc] = fmaf(m, t, e[ch][cc]);
                        }
                }
                if (lane == 31) {
                    aes_stv<4>(wt + warp * 8, e[0]);
                    aes_stv<4>(wt + warp * 8 + 4, e[1]);
                }
                if constexpr (PM != 0) {
                    // pre-delay line (reverb.py:11-31): written here, read behind the barrier
#pragma unroll
                    for (int ch = 0; ch < 2; ++ch) aes_stv<FR>(rings + rs.pre[ch].off + pw, v[ch]);
                }
                aes_bar_sync(AESRV_BAR_COMB, AES_NT);

                // ---------------- phase 2 ----------------
                if constexpr (PRE == AESRV_PRE_BIQUAD) {
                    const FastStage &bs = a.st[0];
                    const double b1 = bs.bq[1], b2 = bs.bq[2], a1 = bs.bq[3], a2 = bs.bq[4];
                    const double w0 = bs.bq_pow[5][0], w1 = bs.bq_pow[5][1], w2 = bs.bq_pow[5][2], w3 = bs.bq_pow[5][3];
                    const double *lt = reinterpret_cast<const double *>(a.lane_tab + lane * FAST_LANE_STRIDE + 4);
                    const double l0 = lt[0], l1 = lt[1], l2 = lt[2], l3 = lt[3];
                    const double *sin = bst + par * 8;
                    double *sout = bst + (par ^ 1) * 8;
#pragma unroll
                    for (int ch = 0; ch < 2; ++ch) {
                        const double cx1 = sin[4 * ch + 0], cx2 = sin[4 * ch + 1], cy1 = sin[4 * ch + 2], cy2 = sin[4 * ch + 3];
                        double C1 = b1 * cx1 + b2 * cx2 - a1 * cy1 - a2 * cy2;
                        double C2 = b2 * cx1 - a2 * cy1;
                        const int u0 = warp > bs.nxw ? warp - bs.nxw : 0;
                        if (u0 > 0) { C1 = 0.0; C2 = 0.0; }
                        for (int t = u0; t < warp; ++t) {
                            const double t1 = fma(w0, C1, fma(w2, C2, wtot[(t * 2 + ch) * 2]));
                            const double t2 = fma(w1, C1, fma(w3, C2, wtot[(t * 2 + ch) * 2 + 1]));
                            C1 = t1; C2 = t2;
                        }
                        double x1 = __shfl_up_sync(0xffffffffu, bq_e1[ch], 1), x2 = __shfl_up_sync(0xffffffffu, bq_e2[ch], 1);
                        if (lane == 0) { x1 = 0.0; x2 = 0.0; }
                        double s1 = fma(l0, C1, fma(l2, C2, x1));           // state entering this thread's chunk
                        double s2 = fma(l1, C1, fma(l3, C2, x2));
                        const double b0 = bs.bq[0];
#pragma unroll
                        for (int j = 0; j < FR; ++j) {
                            // the recurrence again, from the true state (keeping the zero-state outputs of phase 1
                            // across the barrier cost 16 registers the comb warps do not have)
                            const double xj = (double)v[ch][j];
                            const double yy = fma(b0, xj, s1);
                            s1 = fma(b1, xj, fma(-a1, yy, s2));
                            s2 = fma(b2, xj, -a2 * yy);
                            if (len == T) {
                                if (j >= FR - 2 && tid == AES_NT - 1) {
                                    sout[4 * ch + (FR - 1 - j)] = xj; sout[4 * ch + 2 + (FR - 1 - j)] = yy;
                                }
                            } else {
                                if (i0 + j == len - 1) { sout[4 * ch + 0] = xj; sout[4 * ch + 2] = yy; }
                                if (i0 + j == len - 2) { sout[4 * ch + 1] = xj; sout[4 * ch + 3] = yy; }
                            }
                            v[ch][j] = (float)yy;
                        }
                        if (len == 1 && tid == 0) { sout[4 * ch + 1] = cx1; sout[4 * ch + 3] = cy1; }
                    }
                }
                [[maybe_unused]] int pa = 0;                        // pre-delay line: aligned read base, lag4 behind the write slot
                if constexpr (PM != 0) {
                    pa = pw - ((rs.pre[0].lag + 3) & ~3);
                    if (pa < 0) pa += rs.pre[0].len;
                }
                const float *const cp = par ? cp1 : cp0;
                const float4 Ca = aes_lds_v4(cp), Cb = aes_lds_v4(cp + 4);
                const float Cin[2][NC] = { { Ca.x, Ca.y, Ca.z, Ca.w }, { Cb.x, Cb.y, Cb.z, Cb.w } };
                float *const cnext = cst + (par ^ 1) * 8;
                const int u0 = warp > nxw ? warp - nxw : 0;
                float sum[2][FR];
                aes_static_for<0, 2>([&](auto ich) {
                    constexpr int ch = decltype(ich)::value;
                    float uend[NC];
                    float pre[FR];                                  // the reverb's input: the dry frames, or their pre-delayed copies
                    if constexpr (PM == 0) {
#pragma unroll
                        for (int j = 0; j < FR; ++j) pre[j] = v[ch][j];
                    } else {
                        const FRing rg = rs.pre[ch];
                        // (the line lives in the zero-initialised ring area: no first-lap check)
                        aesf_read<FR, 0>(rings + rg.off, pa, ((rg.lag + 3) & ~3) - rg.lag, rg.len, pre);
                    }
                    aes_static_for<0, NC>([&](auto icc) {
                        constexpr int cc = decltype(icc)::value;
                        const float ex = __shfl_up_sync(0xffffffffu, e[ch][cc], 1);
                        float C;
                        if (FAST || nxw == 1) {                          // usual case: h^(32*FR) < 2^-32
                            C = Cin[ch][cc];
                        } else {
                            C = u0 == 0 ? cst[par * 8 + ch * 4 + cc] : 0.0f;
                            for (int t = u0; t < warp; ++t) C = fmaf(hw, C, wt[t * 8 + ch * 4 + cc]);
                        }
                        float u = fmaf(hl, C, lane == 0 ? 0.0f : ex);
                        const float gs = rs.gs[ch][cc];
                        float nb[FR];
#pragma unroll
                        for (int j = 0; j < FR; j += 2) {                   // the one-pole is serial, the rest packs in pairs
                            const float ua = fmaf(h, u, y[ch][cc][j]);
                            u = fmaf(h, ua, y[ch][cc][j + 1]);
                            const float2 n2 = aes_fma2(make_float2(gs, gs), make_float2(ua, u),
                                                       make_float2(pre[j], pre[j + 1]));   // buf[n] = x + g*(1-h)*u
                            nb[j] = n2.x; nb[j + 1] = n2.y;
                            if (cc == 0) { sum[ch][j] = y[ch][0][j]; sum[ch][j + 1] = y[ch][0][j + 1]; }   // reverb.py:235-241: sum starts at 0
                            else {
                                const float2 s2 = aes_add2(make_float2(sum[ch][j], sum[ch][j + 1]),
                                                           make_float2(y[ch][cc][j], y[ch][cc][j + 1]));
                                sum[ch][j] = s2.x; sum[ch][j + 1] = s2.y;
                            }
                        }
                        aesrv_comb_write<aesf_topo_comb(TOPO, ch, cc)>(
                            reinterpret_cast<char *>(rings + aesf_topo_comb_off(TOPO, ch, cc)), wb[ch][cc], nb);
                        uend[cc] = u;
                    });
                    if (tid == AES_NT - 1) aes_stv<4>(cnext + ch * 4, uend);
                });
                aes_stv<FR>(Sc + i0, sum[0]);
                aes_stv<FR>(Sc + T + i0, sum[1]);
                aes_bar_arrive(AESRV_BAR_FULL + par, AESRV_NT);             // the walkers take the tile from here
            }

            // ---------------- phase 3: tile i-1 comes back from the walkers ----------------
            if (has_prev) {
                // (also the barrier between this tile's ring writes and the next tile's ring reads)
                aes_bar_sync(AESRV_BAR_DONE + (par ^ 1), AESRV_NT);
                float o[2][FR];
                aes_ldv_sp<FR, 0>(Sp + i0, o[0]);
                aes_ldv_sp<FR, 0>(Sp + T + i0, o[1]);
#pragma unroll
                for (int ch = 0; ch < 2; ++ch)
#pragma unroll
                    for (int j = 0; j < FR; j += 2)
                        aes_mix_clip2(rdry, make_float2(vprev[ch][j], vprev[ch][j + 1]), rwet,
                                      make_float2(o[ch][j], o[ch][j + 1]), o[ch][j], o[ch][j + 1]);
                if (FAST || (fast_out && rem >= 0)) {       // tile i-1 was a full one
                    __stcs(reinterpret_cast<float4 *>(yp), make_float4(o[0][0], o[1][0], o[0][1], o[1][1]));
                    __stcs(reinterpret_cast<float4 *>(yp) + 1, make_float4(o[0][2], o[1][2], o[0][3], o[1][3]));
                } else {
                    aes_store_frames<FR>(io, b, n0 - T, rem + T < T ? rem + T : T, tid, o);
                }
            } else {
                aes_bar_sync(AESRV_BAR_COMB, AES_NT);       // first tile: nothing to wait for, but the ring writes still
            }                                               // have to be ordered before the next tile's reads
            yp += 2 * T;
            if (has_cur) {
                if constexpr (PM != 0) {
                    // the dry frames come back from the pre-delay line, where phase 1 put them (this thread's own
                    // stores, the slot has not moved yet): they need no registers across phases 2 and 3
#pragma unroll
                    for (int ch = 0; ch < 2; ++ch) aes_ldv_sp<FR, 0>(rings + rs.pre[ch].off + pw, vprev[ch]);
                } else {
#pragma unroll
                    for (int ch = 0; ch < 2; ++ch)
#pragma unroll
                        for (int j = 0; j < FR; ++j) vprev[ch][j] = v[ch][j];
                }
                aes_static_for<0, 2 * NC>([&](auto ic) {
                    constexpr int ch = decltype(ic)::value / NC, cc = decltype(ic)::value % NC;
                    using CR = AesrvComb<aesf_topo_comb(TOPO, ch, cc)>;
                    wb[ch][cc] = aesf_adv(wb[ch][cc], CR::TINC, CR::WRAP);
                });
                if constexpr (PM != 0) pw = aesf_adv(pw, rs.pre[0].tinc, rs.pre[0].len);
                if constexpr (PRE == AESRV_PRE_DELAY) {
#pragma unroll
                    for (int ch = 0; ch < 2; ++ch) {
                        const FRing rg = a.st[0].ring[ch][0];
                        dw[ch] = aesf_adv(dw[ch], rg.tinc, rg.len);
                    }
                }
                q = (q + 1) & (AESRV_NXS - 1);
            }
        };
        // first tile the fast instantiation may take: behind tile 0 and behind the feedback delay's first lap
        int it_fast = 1;
        if constexpr (PRE == AESRV_PRE_DELAY) {
            const int lmax = a.st[0].ring[0][0].lag > a.st[0].ring[1][0].lag ? a.st[0].ring[0][0].lag : a.st[0].ring[1][0].lag;
            it_fast = (lmax + T - 1) / T;
        }
        const bool fast_clip = clip_staged && fast_out && nxw == 1;
        const int nfull = N / T;                            // tiles 0 .. nfull-1 are full
        for (int it = 0; it <= ntiles; ++it, par ^= 1) {
            if (fast_clip && it >= it_fast && it < nfull) tile_iter(std::true_type{}, it);
            else tile_iter(std::false_type{}, it);
        }
        if (PRE == AESRV_PRE_BIQUAD && a.state_out != nullptr) {
            // carried scalars of the biquad (stage 0) at the end of the clip: written by the last tile into the
            // parity opposite to its own (every comb warp is past the DONE barrier behind that write)
            if (tid < 8) a.state_out[b * 32 + tid] = bst[(ntiles & 1) * 8 + tid];
        }
        __syncthreads();                                    // the next clip re-initialises rings and carried values
    }
}

#ifndef AES_CPU_EMU
template <int TOPO, int PRE, int PM>
__global__ void __launch_bounds__(AESRV_NT, 2) aes_rv_kernel(const __grid_constant__ FastArgs a)
{
    aes_rv_body<TOPO, PRE, PM>(a);
}
#endif
