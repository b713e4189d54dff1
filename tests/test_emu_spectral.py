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


# ---- frame lengths with factors 2, 3, 5 only: the four-step path (csrc/aes_spectral_smooth.cuh) ----
SMOOTH_ARGS = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_int,
               C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_void_p]


@pytest.mark.parametrize("M", [16, 60, 96, 600, 1500, 2048, 4050, 9600, 20000])   # all radix mixes, n2 not a multiple of the tile
@pytest.mark.parametrize("nb", [1, 2, 3])
def test_smooth_lengths_gated_roundtrip(M, nb):
    L = emu.lib()
    L.emu_spectral_smooth.argtypes = SMOOTH_ARGS
    rng = np.random.default_rng(M + nb)
    raw = (0.3 * rng.standard_normal((nb, M))).astype(np.float32)
    raw[nb // 2] *= 0.001
    win = np.hanning(M).astype(np.float32)
    mask0 = (0.5 + 0.5 * rng.random((nb, M // 2 + 1))).astype(np.float32)
    mask = mask0.copy()
    y = np.zeros((nb, M), np.float32)
    thr, red, alpha = 0.05, 0.1, 0.8
    n1, n2 = C.c_int(0), C.c_int(0)
    assert L.emu_spectral_smooth(1, raw.ctypes.data, win.ctypes.data, mask.ctypes.data, y.ctypes.data, M, nb, thr, red, alpha,
                                 C.addressof(n1), C.addressof(n2)) == 0
    assert n1.value * n2.value == M and n1.value % 2 == 0
    for b in range(nb):
        X = np.fft.rfft((raw[b] * win).astype(np.float64))
        m = alpha * mask0[b] + (1 - alpha) * np.where(np.abs(X) > thr, 1.0, red)
        want = np.fft.irfft(X * m, M)
        assert np.max(np.abs(mask[b] - m)) < 1e-6, (M, b)
        assert np.max(np.abs(y[b] - want)) < 2e-6, (M, b, n1.value, n2.value)


# the last: BASELINE's frame on the compile-time 960 x 1000 kernels, threshold in the middle of the magnitude
# distribution (a handful of the 480 001 bins sit within rounding of it and flip: 1.5e-5 each at most)
@pytest.mark.parametrize("N,nb,thr,tol", [(48, 1, 0.5, 2e-6), (300, 2, 0.5, 2e-6), (4800, 3, 0.5, 2e-6), (480000, 2, 80.0, 1e-4)])
def test_smooth_whole_clip_mode_reads_clips_and_emits_first_half(N, nb, thr, tol):
    """mode 2 = the file route's single whole-clip block: frame = [zeros(N), mean(x) * hanning(2N)[N:]],
    fresh mask of ones (not kept), output = first N samples on both channels (spectral.py:30-42, 80-100)"""
    L = emu.lib()
    L.emu_spectral_smooth.argtypes = SMOOTH_ARGS
    M = 2 * N
    rng = np.random.default_rng(N)
    x = (0.3 * rng.standard_normal((nb, N, 2))).astype(np.float32)
    win = np.hanning(M).astype(np.float32)
    y = np.full((nb, N, 2), 7.0, np.float32)
    red, alpha = 0.1, 0.8
    assert L.emu_spectral_smooth(2, x.ctypes.data, win.ctypes.data, None, y.ctypes.data, M, nb, thr, red, alpha, None, None) == 0
    for b in range(nb):
        mono = ((x[b, :, 0] + x[b, :, 1]) * np.float32(0.5)) * win[N:]
        fr = np.concatenate([np.zeros(N, np.float32), mono]).astype(np.float64)
        X = np.fft.rfft(fr)
        m = alpha * 1.0 + (1 - alpha) * np.where(np.abs(X) > thr, 1.0, red)
        want = np.fft.irfft(X * m, M)[:N]
        assert 0.2 < np.mean(np.abs(X) > thr) < 0.8 or N < 100000
        assert np.max(np.abs(y[b, :, 0] - want)) < tol and np.array_equal(y[b, :, 0], y[b, :, 1]), (N, b)
        assert np.sqrt(np.mean((y[b, :, 0] - want) ** 2)) < 2e-6


@pytest.mark.parametrize("one_buffer", [False, True])
def test_register_resident_row_pairs_with_a_kept_mask_and_an_odd_frame_count(one_buffer, monkeypatch):
    """BASELINE's frame length in frame mode (mask read and written, three frames: the last transform pair has one):
    aesm_rows10_body for row pairs 1..479 plus aesm_rows_body for the two self-paired rows.  A zero threshold keeps
    every bin on the `mag > thr` side (no bins flipping on rounding), so the bar is the transform's own accuracy."""
    if one_buffer:
        monkeypatch.setenv("AES_EMU_ROWS10_ONE_BUFFER", "1")      # two row pairs per CTA, one exchange buffer per row
    else:
        monkeypatch.delenv("AES_EMU_ROWS10_ONE_BUFFER", raising=False)
    L = emu.lib()
    L.emu_spectral_smooth.argtypes = SMOOTH_ARGS
    M, nb = 960000, 3
    rng = np.random.default_rng(5)
    raw = (0.3 * rng.standard_normal((nb, M))).astype(np.float32)
    win = np.hanning(M).astype(np.float32)
    mask0 = rng.uniform(0.1, 1.0, (nb, M // 2 + 1)).astype(np.float32)
    mask = mask0.copy()
    y = np.zeros((nb, M), np.float32)
    thr, red, alpha = 0.0, 0.1, 0.8
    assert L.emu_spectral_smooth(1, raw.ctypes.data, win.ctypes.data, mask.ctypes.data, y.ctypes.data, M, nb, thr, red, alpha,
                                 None, None) == 0
    for b in range(nb):
        X = np.fft.rfft((raw[b] * win).astype(np.float64))
        m = alpha * mask0[b] + (1 - alpha) * np.where(np.abs(X) > thr, 1.0, red)
        want = np.fft.irfft(X * m, M)
        assert np.max(np.abs(mask[b] - m)) < 1e-6, b
        assert np.max(np.abs(y[b] - want)) < 4e-6, b
        assert np.sqrt(np.mean((y[b] - want) ** 2)) < 5e-7, b


def test_non_smooth_length_has_no_split():
    L = emu.lib()
    L.emu_spectral_smooth.argtypes = SMOOTH_ARGS
    z = np.zeros(2 * 1029, np.float32)
    assert L.emu_spectral_smooth(1, z.ctypes.data, None, None, z.ctypes.data, 2 * 1029, 1, 0.1, 0.1, 0.8, None, None) == 1
