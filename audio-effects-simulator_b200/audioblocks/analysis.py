"""Plot-side analysis of the reference on the GPU: `calculateSpectrumAndChroma` and
`calculateChroma` of assets/02_custom.js:65-154 -- what the page evaluates for the original and the
processed signal at every plot refresh (02_custom.js:179-184).  Both signals of a pair go through one
complex FFT in csrc/aes_analysis.cuh; batches of pairs take one CTA each."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native

FFT_SIZE = 16384            # assets/02_custom.js:7
NOTES = ["C", "C#", "D", "D#", "E", "F", "F#", "G", "G#", "A", "A#", "B"]     # assets/02_custom.js:9


def spectrum_and_chroma(original: np.ndarray, processed: np.ndarray, sample_rate: float, n_fft: int = FFT_SIZE) -> dict:
    """original, processed: (n_samples,) or (n_pairs, n_samples) mono float32 signals with n_samples >= n_fft;
    the last n_fft samples are analysed (02_custom.js:179-181).  Returns the fields of the page's
    `{freqs, magnitudesDB, chroma, peakFreq}` objects for both signals:
    freqs (n_fft/2+1,), magnitudesDB (n_pairs, 2, n_fft/2+1), chroma (n_pairs, 2, 12), peakFreq (n_pairs, 2);
    index [:, 0] is the original, [:, 1] the processed signal (the pair axis is dropped for 1-D input)."""
    a = np.ascontiguousarray(original, dtype=np.float32)
    b = np.ascontiguousarray(processed, dtype=np.float32)
    if a.shape != b.shape or a.ndim not in (1, 2):
        raise _native.AesimError("spectrum_and_chroma: two signals of the same (n_pairs, n_samples) shape")
    single = a.ndim == 1
    a2, b2 = np.atleast_2d(a), np.atleast_2d(b)
    n_pairs, n_samples = a2.shape
    nb = n_fft // 2 + 1
    db = np.empty((n_pairs, 2, nb), np.float32)
    chroma = np.empty((n_pairs, 2, 12), np.float32)
    peak = np.empty((n_pairs, 2), np.float32)
    _native.check(_native.lib().aes_spectrum_chroma_host(
        C.c_void_p(a2.ctypes.data), C.c_void_p(b2.ctypes.data), n_pairs, n_samples, n_fft, float(sample_rate),
        C.c_void_p(db.ctypes.data), None, C.c_void_p(chroma.ctypes.data), C.c_void_p(peak.ctypes.data)))
    out = {"freqs": np.arange(nb) * (float(sample_rate) / n_fft), "magnitudesDB": db, "chroma": chroma, "peakFreq": peak}
    if single:
        out.update(magnitudesDB=db[0], chroma=chroma[0], peakFreq=peak[0])
    return out
