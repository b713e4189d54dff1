"""SpectralFilter transforms (csrc/aes_spectral.cuh: Bluestein over power-of-two FFTs) on the CPU
emulator against numpy's rfft/irfft, for power-of-two and awkward frame lengths."""
import ctypes as C

import numpy as np
import pytest

import emu


@pytest.mark.parametrize("M", [512, 600, 1500, 2 * 1029, 5000, 12000])   # P = 2^10 .. 2^15: 0-5 grid-wide stages in passes of up to 3
def test_gated_rfft_irfft_roundtrip(M):
    check_frames(M, 2)


@pytest.mark.parametrize("nb", [1, 3])          # an unpaired last frame rides alone in its transform
def test_odd_frame_counts(nb):
    check_frames(600, nb)


def check_frames(M, nb):
    L = emu.lib()
    L.emu_spectral_frames.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_float, C.c_float, C.c_float]
    rng = np.random.default_rng(M)
    fr = (0.3 * rng.standard_normal((nb, M))).astype(np.float32) * np.hanning(M).astype(np.float32)
    fr[nb // 2] *= 0.001                                   # a quiet frame: most bins fall under the threshold
    mask0 = (0.5 + 0.5 * rng.random((nb, M // 2 + 1))).astype(np.float32)
    mask = mask0.copy()
    y = np.zeros((nb, M), np.float32)
    thr, red, alpha = 0.05, 0.1, 0.8
    assert L.emu_spectral_frames(fr.ctypes.data, mask.ctypes.data, y.ctypes.data, M, nb, thr, red, alpha) == 0
    for b in range(nb):
        X = np.fft.rfft(fr[b].astype(np.float64))
        m = alpha * mask0[b] + (1 - alpha) * np.where(np.abs(X) > thr, 1.0, red)      # spectral.py:68-71
        want = np.fft.irfft(X * m, M)                                                 # spectral.py:74-77
        assert np.max(np.abs(mask[b] - m)) < 1e-6
        assert np.max(np.abs(y[b] - want)) < 2e-6, (M, b)
