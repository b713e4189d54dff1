"""Spectrum / chromagram analysis (assets/02_custom.js:65-154) on the GPU through the C ABI
(`audioblocks.analysis.spectrum_and_chroma` -> aes_spectrum_chroma_host) against the oracle's float64
restatement of the page's JavaScript (node is not in this image: parity unpinned by the reference)."""
import numpy as np
import pytest

import analysis_case

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ab():
    import audioblocks
    from audioblocks import _native
    _native.lib()                                  # fail loudly if the CUDA library is missing
    return audioblocks


@pytest.mark.parametrize("n_fft,n_samples,fs,n_pairs", [(16384, 131072, 48000.0, 6), (16384, 16384, 44100.0, 3),
                                                         (2048, 5000, 16000.0, 5), (256, 256, 8000.0, 4)])
def test_spectrum_and_chroma_match_the_float64_restatement(ab, n_fft, n_samples, fs, n_pairs):
    """The page's own case first: the last FFT_SIZE = 16384 samples of a PLOT_WINDOW_SIZE = 131072 buffer."""
    a, b = analysis_case.signals(n_fft + 1, n_pairs, n_samples, fs)
    r = ab.analysis.spectrum_and_chroma(a, b, fs, n_fft)
    assert r["freqs"].shape == (n_fft // 2 + 1,) and r["freqs"][1] == fs / n_fft
    analysis_case.check(a, b, fs, n_fft, r["magnitudesDB"], r["chroma"], r["peakFreq"])


def test_a_processed_file_against_its_original(ab):
    """What renderPlots does after a file request (02_custom.js:179-184): original vs the Rain Delay output."""
    import synth
    from audioblocks.engine import file_chain
    fs = 48000
    x = synth.batch(3, 1, 60000)
    y = file_chain(synth.PRESETS["Rain Delay"], fs, channels_in=2).process_batch(x)
    a = np.ascontiguousarray(x[0].mean(axis=1))
    b = np.ascontiguousarray(y[0].mean(axis=1))
    r = ab.analysis.spectrum_and_chroma(a, b, fs)
    assert r["magnitudesDB"].shape == (2, 8193) and r["chroma"].shape == (2, 12)
    analysis_case.check(a[None], b[None], fs, 16384, r["magnitudesDB"][None], r["chroma"][None], r["peakFreq"][None])


def test_many_pairs_come_out_identical(ab):
    """600 copies of two pairs over several waves of CTAs: bit-identical results (the race check of the
    other kernels, applied here)."""
    a, b = analysis_case.signals(5, 2, 16384, 48000.0)
    r = ab.analysis.spectrum_and_chroma(np.tile(a, (300, 1)), np.tile(b, (300, 1)), 48000.0)
    for k in ("magnitudesDB", "chroma", "peakFreq"):
        v = r[k]
        assert np.array_equal(v[0::2], np.broadcast_to(v[0], v[0::2].shape)), k
        assert np.array_equal(v[1::2], np.broadcast_to(v[1], v[1::2].shape)), k


def test_short_signals_are_refused(ab):
    from audioblocks._native import AesimError
    with pytest.raises(AesimError):
        ab.analysis.spectrum_and_chroma(np.zeros(1000, np.float32), np.zeros(1000, np.float32), 48000.0)
