"""Shared by the emulator and the GPU tests of the spectrum / chromagram analysis: signals and the
comparison against the oracle's float64 restatement of assets/02_custom.js:65-154."""
import numpy as np

from oracle import oracle as orc


def signals(seed, n_pairs, n_samples, fs):
    """Original / processed pairs with a few notes, harmonics above 800 and 1500 Hz and a noise floor."""
    rng = np.random.default_rng(seed)
    t = np.arange(n_samples) / fs
    a = np.zeros((n_pairs, n_samples), np.float32)
    b = np.zeros((n_pairs, n_samples), np.float32)
    for p in range(n_pairs):
        for sig in (a, b):
            x = 0.01 * rng.standard_normal(n_samples)
            for _ in range(int(rng.integers(1, 5))):
                midi = rng.integers(40, 100) + rng.uniform(-0.2, 0.2)
                f = 440.0 * 2 ** ((midi - 69) / 12)
                x += rng.uniform(0.05, 0.4) * np.sin(2 * np.pi * f * t + rng.uniform(0, 6.28))
            sig[p] = np.clip(x, -1, 1).astype(np.float32)
    return a, b


def check(a, b, fs, n_fft, db, chroma, peak, lin=None):
    """db (n_pairs, 2, n_fft/2+1), chroma (n_pairs, 2, 12), peak (n_pairs, 2) against the oracle."""
    for p in range(a.shape[0]):
        for s, sig in enumerate((a, b)):
            x = sig[p, -n_fft:]
            freqs, want_db, want_ch, want_pk, want_lin = orc.spectrum_and_chroma(x, fs)
            # dB: the page adds 1e-9 to |X|/n_fft, so the scale is absolute; a float32 transform is good to
            # ~1e-6 of the window's peak bin: compare as linear amplitudes with that bar, and in dB where a
            # bin stands 60 dB above it
            amp_got = 10.0 ** (db[p, s].astype(np.float64) / 20)
            amp_want = 10.0 ** (want_db / 20)
            bar = 2e-6 * amp_want.max() + 2e-9
            assert np.max(np.abs(amp_got - amp_want)) <= bar, (p, s, np.max(np.abs(amp_got - amp_want)), bar)
            loud = amp_want > 1e3 * bar
            assert np.max(np.abs(db[p, s][loud] - want_db[loud])) <= 0.02, (p, s)
            if lin is not None:
                assert np.max(np.abs(lin[p, s] - want_lin)) <= 2e-6 * want_lin.max() + 1e-6
            assert abs(peak[p, s] - want_pk) <= 1e-3 * max(want_pk, 1.0), (p, s, peak[p, s], want_pk)
            # chroma: a bin within float32 noise of the 15 % gate may fall on either side: bracket it
            lo = np.minimum.reduce([orc.spectrum_and_chroma(x, fs, sc)[2] for sc in (1 - 1e-4, 1.0, 1 + 1e-4)])
            hi = np.maximum.reduce([orc.spectrum_and_chroma(x, fs, sc)[2] for sc in (1 - 1e-4, 1.0, 1 + 1e-4)])
            got = chroma[p, s].astype(np.float64)
            assert np.all(got >= lo - 2e-4) and np.all(got <= hi + 2e-4), (p, s, got, want_ch)
