"""Biquad filter (reference src/audioblocks/filter.py:42-113)."""
from __future__ import annotations

import math

import numpy as np

from . import _native
from .core import NativeEffect, SmoothParam


class FilterEffect(NativeEffect):
    """RBJ low-/high-/band-pass biquad, Direct Form I per channel.  `filter_type`
    0/1/2 as in the reference (filter.py:44,72-95); 3 selects an RBJ peaking EQ with
    `gain_db`, an extension the reference does not have."""

    def __init__(self, filter_type=0.0, cutoff_hz=1000.0, q=0.707, gain_db=0.0):
        self.filter_type = SmoothParam(filter_type, 0.0, 3.0 if int(round(filter_type)) == 3 else 2.0)
        self.cutoff_hz = SmoothParam(cutoff_hz, 20.0, 20000.0)
        self.q = SmoothParam(q, 0.1, 10.0)
        self.gain_db = SmoothParam(gain_db, -24.0, 24.0)
        self._state = np.zeros((1, 4), dtype=np.float32)      # [x1, x2, y1, y2] per channel
        self._fs = 48000.0

    def set_filter_type(self, v): self.filter_type.set_target(v)
    def set_cutoff_hz(self, v): self.cutoff_hz.set_target(v)
    def set_q(self, v): self.q.set_target(v)
    def set_gain_db(self, v): self.gain_db.set_target(v)

    def prepare(self, sample_rate: int, channels_in: int, channels_out: int, blocksize: int):
        self._sr = sample_rate
        self._fs = float(sample_rate)
        if self._state.shape[0] != channels_out:               # state survives otherwise (filter.py:59-60)
            self._state = np.zeros((channels_out, 4), dtype=np.float32)

    def _calc_coeffs(self, f_type_val, fc, q, gain_db=0.0):
        """Normalised (b0, b1, b2, a1, a2), RBJ cookbook (reference filter.py:62-98)."""
        w0 = 2.0 * math.pi * fc / self._fs
        cos_w0, sin_w0 = math.cos(w0), math.sin(w0)
        alpha = sin_w0 / (2.0 * q)
        kind = int(round(f_type_val))
        a1 = -2 * cos_w0
        if kind == 3:
            big_a = 10.0 ** (gain_db / 40.0)
            b0, b1, b2 = 1 + alpha * big_a, -2 * cos_w0, 1 - alpha * big_a
            a0, a2 = 1 + alpha / big_a, 1 - alpha / big_a
        else:
            a0, a2 = 1 + alpha, 1 - alpha
            if kind == 0:
                b1 = 1 - cos_w0
                b0 = b2 = (1 - cos_w0) / 2
            elif kind == 1:
                b1 = -(1 + cos_w0)
                b0 = b2 = (1 + cos_w0) / 2
            else:
                b0, b1, b2 = alpha, 0, -alpha
        return (b0 / a0, b1 / a0, b2 / a0, a1 / a0, a2 / a0)

    def _stages(self, frames):
        f_type = self.filter_type.step_towards(1.0)
        fc = self.cutoff_hz.step_towards(self.cutoff_hz.current * 0.1)
        q_val = self.q.step_towards(0.1)
        g_db = self.gain_db.step_towards(1.0)
        d = _native.StageDesc()
        d.kind = _native.BIQUAD
        for i, v in enumerate(self._calc_coeffs(f_type, fc, q_val, g_db)):
            d.p[i] = v
        st = self._state
        for c in range(2):
            row = st[min(c, st.shape[0] - 1)]
            for k in range(4):
                d.p[8 + 4 * c + k] = float(row[k])
        return [d]

    def _at_rest(self):
        return not self._state.any()        # zero DF-I memories: zero in, zero out, memories stay zero

    def _absorb(self, desc, frames, silent):
        self._n_total += frames
        for c in range(self._state.shape[0]):
            self._state[c, :] = [desc.p[8 + 4 * c + k] for k in range(4)]      # stored as f32, filter.py:35-40

    def _advance(self, frames, silent, final=None):
        self._n_total += frames
        # The filter keeps no delay line: its DF-I scalars come back from the device and
        # are stored as float32 between calls, like the reference's state array (filter.py:35-40).
        if final is not None:
            for c in range(self._state.shape[0]):
                self._state[c, :] = final[4 * c:4 * c + 4]
