"""IR-convolution reverb kernels (csrc/aes_convreverb.cuh) on the CPU emulator (FFT sizes 2^8, 2^11, 2^14)
against the oracle's float64 scipy.signal.fftconvolve restatement."""
import ctypes as C

import numpy as np
import pytest

import emu
import synth
from oracle import oracle as orc


def conv(ir, x, dry, wet, log2n=8, pc=4, mac_grid=3):
    L = emu.lib()
    L.emu_convreverb.argtypes = [C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_longlong, C.c_longlong,
                                 C.c_float, C.c_float, C.c_int, C.c_int, C.c_int]
    y = np.full_like(x, 9.0)
    assert L.emu_convreverb(ir.ctypes.data, ir.shape[0], x.ctypes.data, y.ctypes.data, x.shape[0], x.shape[1], dry, wet,
                            log2n, pc, mac_grid) == 0
    return y


@pytest.mark.parametrize("n_taps,n", [(1, 300), (128, 128), (129, 1000), (700, 2049)])
def test_partitioned_fft_convolution_matches_float64_oracle(n_taps, n):
    ir = orc.synthetic_ir(n_taps, rt60=0.004)
    x = synth.batch(0, 3, n)
    y = conv(ir, x, 0.7, 0.5)
    for b in range(3):
        want = np.zeros_like(x[b])
        orc.OConvReverb(ir, 0.7, 0.5).process_into(x[b], want)
        mx, snr = synth.err_stats(y[b], want)
        assert mx <= 1e-5 and snr >= 100.0, (n_taps, n, b, mx, snr)


def test_unit_impulse_ir_is_a_pure_delay():
    ir = np.zeros((200, 2), np.float32)
    ir[37, 0] = 1.0
    ir[150, 1] = 1.0          # different delays per channel: exercises the L/R un-mixing of the packed spectrum
    x = synth.batch(4, 1, 1500) * np.float32(0.5)
    y = conv(ir, x, 0.0, 1.0)[0]
    want = np.zeros_like(x[0])
    want[37:, 0] = x[0, :-37, 0]
    want[150:, 1] = x[0, :-150, 1]
    assert np.max(np.abs(y - want)) <= 1e-6


@pytest.mark.parametrize("log2n,pc,n_taps,n,grid", [
    (8, 4, 700, 2049, 1),        # 6 partitions in two MAC passes of 4 (the second accumulates), one persistent CTA
    (8, 9, 1100, 3000, 2),       # 9 partitions in one pass
    (8, 18, 2500, 4000, 5),      # 20 partitions: 18 + a second pass with zero padding
    (11, 4, 3000, 5000, 7),      # 2048-point transforms (radix 16 / 16 / 8), 8 bin tiles
    (14, 4, 9000, 20000, 64),    # the production size: 16384 points (radix 32 / 32 / 16), 64 bin tiles
])
def test_fft_sizes_and_mac_passes(log2n, pc, n_taps, n, grid):
    ir = orc.synthetic_ir(n_taps, rt60=n_taps / 48000.0)
    x = synth.batch(1, 3, n)
    y = conv(ir, x, 0.6, 0.4, log2n, pc, grid)
    for b in range(3):
        want = np.zeros_like(x[b])
        orc.OConvReverb(ir, 0.6, 0.4).process_into(x[b], want)
        mx, snr = synth.err_stats(y[b], want)
        assert mx <= 1e-5 and snr >= 100.0, (log2n, pc, b, mx, snr)
