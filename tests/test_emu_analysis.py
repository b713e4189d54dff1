"""Spectrum / chromagram kernel (csrc/aes_analysis.cuh) on the CPU emulator against the oracle's float64
restatement of assets/02_custom.js:65-154 (the JS itself cannot run here: no node; parity unpinned)."""
import ctypes as C

import numpy as np
import pytest

import analysis_case
import emu


def run(a, b, fs, n_fft):
    L = emu.lib()
    L.emu_spectrum_chroma.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong, C.c_longlong, C.c_int, C.c_double,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    n_pairs, n_samples = a.shape
    nb = n_fft // 2 + 1
    db = np.zeros((n_pairs, 2, nb), np.float32)
    lin = np.zeros_like(db)
    chroma = np.zeros((n_pairs, 2, 12), np.float32)
    peak = np.zeros((n_pairs, 2), np.float32)
    assert L.emu_spectrum_chroma(a.ctypes.data, b.ctypes.data, n_pairs, n_samples, n_fft, fs, db.ctypes.data,
                                 lin.ctypes.data, chroma.ctypes.data, peak.ctypes.data) == 0
    return db, lin, chroma, peak


@pytest.mark.parametrize("n_fft,n_samples,fs,n_pairs", [(256, 256, 8000.0, 3), (2048, 3000, 16000.0, 3),
                                                         (16384, 20000, 48000.0, 2), (16384, 16384, 44100.0, 1)])
def test_spectrum_and_chroma_match_the_float64_restatement(n_fft, n_samples, fs, n_pairs):
    a, b = analysis_case.signals(n_fft, n_pairs, n_samples, fs)
    db, lin, chroma, peak = run(a, b, fs, n_fft)
    analysis_case.check(a, b, fs, n_fft, db, chroma, peak, lin)


def test_silence_gives_the_floor_and_no_chroma():
    a = np.zeros((1, 2048), np.float32)
    db, lin, chroma, peak = run(a, a, 48000.0, 2048)
    assert np.allclose(db, -180.0, atol=1e-3) and np.all(chroma == 0)
