// aes_rv_build.h -- host side of the pipelined reverb-chain kernels (aes_rv_kernel.cuh): which
// instantiation, if any, serves a plan.  Pure host C++ (shared with tests/cpu_emu).
#pragma once
#include <stdlib.h>
#include "aes_fast_build.h"
#include "aes_rv_kernel.cuh"

// X(TOPO, PRE, PM): the instantiations (each for the 48 kHz and the 44.1 kHz reverb topology)
#define AESRV_SHAPES(X)                                                                         \
    X(AESF_TOPO_48K, AESRV_PRE_DELAY, 0)      /* Rain Delay                                  */ \
    X(AESF_TOPO_44K, AESRV_PRE_DELAY, 0)                                                        \
    X(AESF_TOPO_48K, AESRV_PRE_NONE, 0)       /* bare reverb                                 */ \
    X(AESF_TOPO_44K, AESRV_PRE_NONE, 0)                                                         \
    X(AESF_TOPO_48K, AESRV_PRE_NONE, 1)       /* Cathedral: reverb with a pre-delay line     */ \
    X(AESF_TOPO_44K, AESRV_PRE_NONE, 1)                                                         \
    X(AESF_TOPO_48K, AESRV_PRE_BIQUAD, 0)     /* Guitar Filter                               */ \
    X(AESF_TOPO_44K, AESRV_PRE_BIQUAD, 0)

// Shape of a plan in rv terms: true when (pre, pm) name an instantiated kernel for a chain whose
// FastArgs / codes were produced by aes_fast_build and whose reverb has compile-time topology `topo`.
static inline bool aes_rv_shape(const FastArgs &fa, const int codes[AESF_MAX_STAGES], int topo, int *pre, int *pm)
{
    if (topo == AESF_TOPO_NONE) return false;
    int sr;
    if (codes[1] == 0 && codes[2] == 0 && codes[3] == 0) { *pre = AESRV_PRE_NONE; sr = 0; }
    else if (codes[2] == 0 && codes[3] == 0 && codes[0] == AESF_DELAY_PF) { *pre = AESRV_PRE_DELAY; sr = 1; }
    else if (codes[2] == 0 && codes[3] == 0 && codes[0] == AESF_BIQUAD) {
        // The f64 biquad in front of the reverb does not fit the comb warps' 104 registers without spilling
        // (141 against 163 Gsamples/s for aes_fast_kernel on Guitar Filter, B200): off unless asked for.
        if (!getenv("AES_RV_BIQUAD")) return false;
        *pre = AESRV_PRE_BIQUAD; sr = 1;
    }
    else return false;
    if (*pre == AESRV_PRE_DELAY) {
        // the walkers stage a tile's line samples two tiles ahead: everything it reads must have been stored by then
        for (int ch = 0; ch < 2; ++ch)
            if (fa.st[0].ring[ch][0].lag < 2 * AES_NT * 4 + 4) return false;
    }
    const int rc = codes[sr];
    if (AESF_KIND(rc) != AESK_REVERB || AESF_NC(rc) != 4 || AESF_NA(rc) != 2) return false;
    const int premode = AESF_PREMODE(rc);
    if (premode == 0) { *pm = 0; return true; }
    // a pre-delay line is written in phase 1 and read behind the barrier, whatever its length; it has
    // to live in shared memory, and only the bare reverb is instantiated with one
    if (*pre != AESRV_PRE_NONE || fa.st[sr].glob) return false;
    *pm = 1;
    return true;
}
