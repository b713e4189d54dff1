"""Schroeder/Moorer reverb (reference src/audioblocks/reverb.py:72-277)."""
from __future__ import annotations

from . import _native
from .core import NativeEffect, SmoothParam


class ReverbEffect(NativeEffect):
    """Per side: pre-delay -> parallel damped feedback combs -> series all-passes ->
    dry/wet mix and clip; left/right line lengths decorrelated by +-jitter_ms."""

    def __init__(self, *, comb_times_ms=(29.7, 37.1, 41.1, 43.7), allpass_times_ms=(5.0, 1.7),
                 allpass_gain=0.6, jitter_ms=0.3, max_delay_ms=200.0, max_pre_delay_ms=100.0,
                 mix_dry=0.7, mix_wet=0.5, rt60_s=1.5, damp=0.3, pre_delay_ms=0.0,
                 step_samples=2.0, rt60_step=0.05, damp_step=0.02):
        self._comb_ms_base = tuple(float(v) for v in comb_times_ms)
        self._ap_ms_base = tuple(float(v) for v in allpass_times_ms)
        self._ap_gain = float(allpass_gain)
        self._jitter_ms = float(jitter_ms)
        self._max_delay_ms = float(max_delay_ms)
        self._max_pre_ms = float(max_pre_delay_ms)
        self.mix_dry = float(mix_dry)
        self.mix_wet = float(mix_wet)
        self.rt60_s = SmoothParam(rt60_s, 0.1, 10.0)
        self.damp = SmoothParam(damp, 0.0, 0.99)
        self.pre_delay_ms = SmoothParam(pre_delay_ms, 0.0, self._max_pre_ms)
        self._step_samples = float(step_samples)
        self._rt60_step = float(rt60_step)
        self._damp_step = float(damp_step)
        self._delay_step_ms = 0.1
        self._fs = 48000
        self._lines = None

    def set_rt60_s(self, seconds: float): self.rt60_s.set_target(seconds)
    def set_damp(self, value: float): self.damp.set_target(value)
    def set_pre_delay_ms(self, ms: float): self.pre_delay_ms.set_target(ms)

    def set_mix(self, dry: float | None = None, wet: float | None = None):
        if dry is not None:
            self.mix_dry = float(dry)
        if wet is not None:
            self.mix_wet = float(wet)

    def set_mix_wet(self, wet: float): self.mix_wet = wet
    def set_mix_dry(self, dry: float): self.mix_dry = dry

    def _side_lengths(self, jitter: float):
        """Integer line lengths of one side (reference reverb.py:158-177)."""
        cap = self._max_delay_ms - 1.0
        combs = [max(1, int(self._fs * min(ms + jitter, cap) / 1000.0)) for ms in self._comb_ms_base]
        aps = [max(1, int(self._fs * min(ms + jitter * 0.2, cap) / 1000.0)) for ms in self._ap_ms_base]
        return combs, aps

    def prepare(self, sample_rate: int, channels_in: int, channels_out: int, blocksize: int):
        # every prepare rebuilds all lines from zero (reverb.py:180-201)
        self._sr = self._fs = int(sample_rate)
        self._delay_step_ms = 1000.0 * (self._step_samples / float(self._fs))
        self._lines = (self._side_lengths(+self._jitter_ms), self._side_lengths(-self._jitter_ms))
        self._pre_size = max(1, int(self._fs * self._max_pre_ms / 1000.0) + 1)
        self._reset_lines()

    def _blob_floats(self):
        n = 32                                            # double lp[2][8]
        for combs, aps in self._lines:
            n += self._pre_size + sum(c + 1 for c in combs) + sum(a + 1 for a in aps)
        return n

    def _stream_fields(self, desc):
        desc.q[29] = self._n_total
        desc.q[30] = self._pre_size

    def _g_from_rt60(self, L_samples: int, fs: int, rt60_s: float) -> float:
        return 10.0 ** (-3.0 * (float(L_samples) / float(fs)) / max(1e-3, rt60_s))

    def _stages(self, frames):
        if self._lines is None:
            raise RuntimeError("ReverbEffect.prepare() has not been called")
        if len(self._comb_ms_base) > 8 or len(self._ap_ms_base) > 4:
            raise ValueError("ReverbEffect (B200): at most 8 combs and 4 all-passes per side")
        rt60_now = self.rt60_s.step_towards(self._rt60_step)
        damp_now = self.damp.step_towards(self._damp_step)
        pre_ms_now = self.pre_delay_ms.step_towards(self._delay_step_ms)
        pre_ds = int(self._fs * pre_ms_now / 1000.0)
        if pre_ds >= self._pre_size:
            pre_ds = self._pre_size - 1
        d = _native.StageDesc()
        d.kind = _native.REVERB
        d.q[0], d.q[1], d.q[2] = len(self._comb_ms_base), len(self._ap_ms_base), pre_ds
        d.p[0], d.p[1], d.p[2], d.p[3] = self.mix_dry, self.mix_wet, damp_now, self._ap_gain
        for s, (combs, aps) in enumerate(self._lines):
            for c, n in enumerate(combs):
                d.q[4 + 8 * s + c] = n
                d.p[4 + 8 * s + c] = self._g_from_rt60(n, self._fs, rt60_now)
            for k, n in enumerate(aps):
                d.q[20 + 4 * s + k] = n
        return [d]
