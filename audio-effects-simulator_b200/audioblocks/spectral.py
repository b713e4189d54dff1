"""Spectral noise filter (reference src/audioblocks/spectral.py:5-100).

Per block: shift the analysis buffer, rfft of the Hann-windowed frame of 2*hop samples,
per-bin magnitude gate with a temporally smoothed mask, irfft, overlap-add with one block of
latency.  hop follows the chain's block size (spectral.py:30-42), so on the whole-file path the
frame is 2N samples long and the block emits the zero-padded half of a single frame -- near
silence; that quirk is reproduced, not fixed.

The transforms run on the GPU: frame lengths whose factors are 2, 3 and 5 (BASELINE's 960 000) as a
four-step mixed-radix FFT with the gate fused between the row transforms (csrc/aes_spectral_smooth.cuh),
any other length by Bluestein's chirp-z over power-of-two FFTs (csrc/aes_spectral.cuh); the buffer shifting and the overlap-add bookkeeping around them are
the same few numpy lines as in the reference.  There is no CPU fallback for the transforms."""
from __future__ import annotations

import numpy as np

from . import _native
from .core import Effect, SmoothParam


class SpectralFilter(Effect):
    def __init__(self, threshold_db=-40.0, reduction=0.5, smoothing=0.8):
        self.threshold_db = SmoothParam(threshold_db, -80.0, 0.0)
        self.reduction = SmoothParam(reduction, 0.0, 1.0)
        self.alpha_param = smoothing
        self._plans: dict[int, _native.SpectralPlan] = {}
        self._alloc(256)

    def _alloc(self, hop: int):
        self.blocksize = self.hop = hop
        self.n_fft = 2 * hop
        self.in_buffer = np.zeros(self.n_fft, dtype=np.float32)
        self.out_accum = np.zeros(self.n_fft, dtype=np.float32)
        self.mask_smooth = np.ones(self.n_fft // 2 + 1, dtype=np.float32)

    def set_threshold_db(self, v): self.threshold_db.set_target(v)
    def set_reduction(self, v): self.reduction.set_target(v)

    def prepare(self, sample_rate: int, channels_in: int, channels_out: int, blocksize: int):
        if blocksize != self.hop:                      # re-initialise when the block size changes
            self._alloc(blocksize)

    def _plan(self, frame_len: int) -> _native.SpectralPlan:
        pl = self._plans.get(frame_len)
        if pl is None:
            for old in self._plans.values():
                old.close()
            self._plans = {frame_len: _native.SpectralPlan(frame_len)}
            pl = self._plans[frame_len]
        return pl

    def _params(self):
        th_db = self.threshold_db.step_towards(1.0)
        red = self.reduction.step_towards(0.05)
        return 10.0 ** (th_db / 20.0), red

    def process_into(self, x_in: np.ndarray, out: np.ndarray) -> None:
        thr, red = self._params()
        if x_in.shape[0] != self.hop:
            self._alloc(x_in.shape[0])
        hop = self.hop
        self.in_buffer[:-hop] = self.in_buffer[hop:]
        if x_in.dtype == np.float32 and x_in.ndim == 2 and x_in.shape[1] == 2:
            # np.mean(x_in, axis=1) of float32 stereo is (L + R) rounded to float32, then halved;
            # on the two column views that is 5x faster than the strided reduction and bit-identical
            self.in_buffer[-hop:] = (x_in[:, 0] + x_in[:, 1]) * np.float32(0.5)
        else:
            self.in_buffer[-hop:] = np.mean(x_in, axis=1)
        if not self.in_buffer.any():
            # an all-zero frame (the warm-up blocks, silent stretches): every bin is under the threshold,
            # the mask relaxes towards `reduction` (spectral.py:68-71) and the frame adds nothing
            a = np.float32(self.alpha_param)
            self.mask_smooth = a * self.mask_smooth + (np.float32(1.0) - a) * np.float32(red)
        else:
            mask = self.mask_smooth[None, :].copy()
            y = self._plan(self.n_fft).frames_host(self.in_buffer[None, :].copy(), mask, thr, red, float(self.alpha_param))
            self.mask_smooth = mask[0]
            self.out_accum += y[0]
        for c in range(out.shape[1]):
            out[:, c] = self.out_accum[:hop]
        self.out_accum[:-hop] = self.out_accum[hop:]
        self.out_accum[-hop:] = 0.0

    def process_batch(self, x: np.ndarray, out: np.ndarray | None = None) -> np.ndarray:
        """Whole clips, every clip from the state a re-prepare at its frame count leaves:
        x (B, frames, 2) float32 -> (B, frames, 2) float32."""
        x = np.ascontiguousarray(x, np.float32)
        if out is None:
            out = np.empty_like(x)
        thr, red = self._params()
        if x.shape[0] and x.shape[1]:
            self._plan(2 * x.shape[1]).process_host(x, out, thr, red, float(self.alpha_param))
        return out
