"""Spectral noise filter (reference src/audioblocks/spectral.py:5-100).

The reference block is a per-block rfft -> magnitude mask -> irfft overlap-add with
one block of latency; on the whole-file path (hop == N) it degenerates to emitting
the zero-padded half of a single 2N-point frame.  It is the lowest
priority row of the scope table (8-a8) and has no CUDA implementation yet: the
class keeps the reference's constructor and setters so presets can be built, and
fails loudly when asked to process -- there is no CPU fallback."""
from __future__ import annotations

import numpy as np

from .core import Effect, SmoothParam


class SpectralFilter(Effect):
    def __init__(self, threshold_db=-40.0, reduction=0.5, smoothing=0.8):
        self.threshold_db = SmoothParam(threshold_db, -80.0, 0.0)
        self.reduction = SmoothParam(reduction, 0.0, 1.0)
        self.alpha_param = smoothing
        self.blocksize = self.hop = 256
        self.n_fft = 512

    def set_threshold_db(self, v): self.threshold_db.set_target(v)
    def set_reduction(self, v): self.reduction.set_target(v)

    def prepare(self, sample_rate: int, channels_in: int, channels_out: int, blocksize: int):
        if blocksize != self.hop:
            self.blocksize = self.hop = blocksize
            self.n_fft = 2 * blocksize

    def process_into(self, x_in: np.ndarray, out: np.ndarray) -> None:
        raise NotImplementedError(
            "SpectralFilter has no CUDA implementation yet (scope row 8-a8, lowest priority); "
            "audioblocks (B200) has no CPU fallback")
