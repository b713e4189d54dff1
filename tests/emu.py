"""Test helper: run the product's chain kernel SOURCE on the CPU emulator
(tests/cpu_emu) with plans resolved by the product's own host logic.
Test infrastructure only -- the product never loads libaes_emu.so."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

import audioblocks as ab
from audioblocks import _native
from audioblocks.engine import make_effect

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cpu_emu")
_lib = None


def lib():
    global _lib
    if _lib is None:
        import fcntl
        with open(os.path.join(_DIR, ".build.lock"), "w") as lk:      # xdist workers build once, one at a time
            fcntl.flock(lk, fcntl.LOCK_EX)
            subprocess.check_call(["make", "-s", "-C", _DIR], stdout=subprocess.DEVNULL)
        L = C.CDLL(os.path.join(_DIR, "libaes_emu.so"))
        L.emu_last_error.restype = C.c_char_p
        L.emu_chain_run.argtypes = [C.POINTER(_native.StageDesc), C.c_int, C.c_int, C.c_void_p, C.c_int,
                                    C.c_void_p, C.c_int, C.c_longlong, C.c_longlong, C.c_int, C.c_void_p]
        _lib = L
    return _lib


def resolved_descs(config, fs, frames, channels_in=1, blocksize=1024):
    """The descriptors the product would hand to aes_chain_plan_create for a file
    of `frames` frames: build at `blocksize`, two silent warm-up blocks (host-side
    state advance only), re-prepare at `frames` (engine.py:86-102)."""
    chain = ab.EffectsChain(fs, channels_in, 2, blocksize)
    for cfg in config:
        fx = make_effect(cfg)
        if fx is not None:
            chain.add(fx)
    for _ in range(2):
        for fx in chain.effects:
            fx._stages(blocksize)
            fx._advance(blocksize, True, None)
    chain._ensure_blocksize(frames)
    return chain.stage_descs(frames)


def run(descs, fs, x, out_dtype=np.float32, grid=1, state_out=None):
    """x: (B, N, 1|2) f32 or (B, N, 2) int16.  state_out: optional (B, 16*len(descs)) f64."""
    B, N, ch = x.shape
    if x.dtype == np.int16:
        fmt_in = _native.FMT_I16_DOWNMIX
    else:
        fmt_in = _native.FMT_F32_MONO if ch == 1 else _native.FMT_F32_STEREO
    x = np.ascontiguousarray(x)
    y = np.full((B, N, 2), 77, out_dtype)
    fmt_out = _native.FMT_I16_STEREO if out_dtype == np.int16 else _native.FMT_F32_STEREO
    arr = _native.desc_array(descs)
    sp = state_out.ctypes.data if state_out is not None else None
    rc = lib().emu_chain_run(arr, len(descs), fs, x.ctypes.data, fmt_in, y.ctypes.data, fmt_out, B, N, grid, sp)
    if rc != 0:
        raise RuntimeError(lib().emu_last_error().decode())
    return y
