"""Stereo feedback delay (reference src/audioblocks/delay.py:43-96)."""
from __future__ import annotations

from . import _native
from .core import NativeEffect, SmoothParam


class StereoDelayEffect(NativeEffect):
    """Two independent feedback delay lines (R = L + offset_ms), dry/wet mix, hard
    clip.  Constructor keywords, setters and SmoothParam attributes are those of the
    reference (delay.py:48-71)."""

    def __init__(self, max_delay_ms=1500.0, mix_dry=0.8, mix_wet=0.8, offset_ms=30.0,
                 delay_ms=375.0, feedback=0.2, fb_step=0.02, step_samples=2.0):
        self.max_delay_ms = max_delay_ms
        self.mix_dry = mix_dry
        self.mix_wet = mix_wet
        self.offset_ms = offset_ms
        self.delay_ms = SmoothParam(delay_ms, 1.0, max_delay_ms - 1.0)
        self.feedback = SmoothParam(feedback, 0.0, 0.95)
        self._fb_step = fb_step
        self._step_samples = step_samples
        self._delay_step_ms = 0.1
        self._size = 1

    def set_delay_ms(self, v: float): self.delay_ms.set_target(v)
    def nudge_delay_ms(self, dv: float): self.delay_ms.nudge(dv)
    def set_feedback(self, v: float): self.feedback.set_target(v)
    def set_mix_dry(self, v: float): self.mix_dry = v
    def set_mix_wet(self, v: float): self.mix_wet = v
    def set_offset_ms(self, v: float): self.offset_ms = v

    def prepare(self, sample_rate: int, channels_in: int, channels_out: int, blocksize: int):
        # DelayLine.configure zeroes both rings on every prepare (delay.py:31-35,73-78)
        self._sr = sample_rate
        self._size = int(sample_rate * self.max_delay_ms / 1000.0) + 1
        self._delay_step_ms = 1000.0 * (self._step_samples / sample_rate)
        self._reset_lines()

    def _blob_floats(self):
        return 2 * self._size

    def _stream_fields(self, desc):
        desc.q[29] = self._n_total
        desc.q[30] = self._size

    def lags(self, d_left_ms: float):
        """Integer lags with the reference's own float expressions (delay.py:38-40,84);
        the ring index arithmetic downstream is exact, so these must be too."""
        d_right_ms = min(d_left_ms + self.offset_ms, self.max_delay_ms - 1.0)
        out = []
        for ms in (d_left_ms, d_right_ms):
            ds = int(self._sr * ms / 1000.0)
            if ds >= self._size:
                ds = self._size - 1
            # dS == 0 reads the slot about to be overwritten: an effective lag of `size`
            out.append(ds if ds > 0 else self._size)
        return out

    def _stages(self, frames):
        d_now = self.delay_ms.step_towards(self._delay_step_ms)
        fb_now = self.feedback.step_towards(self._fb_step)
        d = _native.StageDesc()
        d.kind = _native.DELAY
        d.q[0], d.q[1] = self.lags(d_now)
        d.p[0], d.p[1], d.p[2] = fb_now, float(self.mix_dry), float(self.mix_wet)
        return [d]
