"""Noise gate (reference src/audioblocks/gate.py:45-90)."""
from __future__ import annotations

import numpy as np

from . import _native
from .core import NativeEffect, SmoothParam


class NoiseGateEffect(NativeEffect):
    """Stereo-linked threshold detector driving an attack/release one-pole gain."""

    def __init__(self, threshold_db=-40.0, attack_ms=10.0, release_ms=100.0):
        self.threshold_db = SmoothParam(threshold_db, -80.0, 0.0)
        self.attack_ms = SmoothParam(attack_ms, 1.0, 500.0)
        self.release_ms = SmoothParam(release_ms, 10.0, 1000.0)
        self._gain_state = 0.0          # starts closed (gate.py:53); prepare() never resets it
        self._fs = 48000.0

    def set_threshold_db(self, v): self.threshold_db.set_target(v)
    def set_attack_ms(self, v): self.attack_ms.set_target(v)
    def set_release_ms(self, v): self.release_ms.set_target(v)

    def prepare(self, sample_rate: int, channels_in: int, channels_out: int, blocksize: int):
        self._sr = sample_rate
        self._fs = float(sample_rate)

    def _calc_coeff(self, time_ms):
        t = max(1e-3, time_ms * 1e-3)
        return float(1.0 - np.exp(-2.2 / (t * self._fs)))        # gate.py:63-69

    def _stages(self, frames):
        th_db = self.threshold_db.step_towards(1.0)
        att_ms = self.attack_ms.step_towards(5.0)
        rel_ms = self.release_ms.step_towards(10.0)
        d = _native.StageDesc()
        d.kind = _native.GATE
        d.p[0] = 10.0 ** (th_db / 20.0)
        d.p[1], d.p[2] = self._calc_coeff(att_ms), self._calc_coeff(rel_ms)
        d.p[3] = self._gain_state
        self._rel_now = d.p[2]
        return [d]

    def _absorb(self, desc, frames, silent):
        self._n_total += frames
        self._gain_state = float(desc.p[3])

    def _at_rest(self):
        return True                 # the only memory is the gain; silence times any gain is silence

    def _rest_block(self, frames):
        # level 0 never exceeds the threshold: the gain releases towards 0 (gate.py:33-40)
        self._n_total += frames
        g, k = self._gain_state, 1.0 - self._rel_now
        if g != 0.0:
            if frames <= 8192:
                for _ in range(frames):
                    g = k * g
            else:
                g *= k ** frames
        self._gain_state = g

    def _advance(self, frames, silent, final=None):
        self._n_total += frames
        if final is not None:
            self._gain_state = float(final[0])      # gain after the block's last frame (gate.py:42)
