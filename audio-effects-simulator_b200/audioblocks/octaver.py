"""Granular octaver / pitch shifter (reference src/audioblocks/octaver.py:84-150)."""
from __future__ import annotations

from . import _native
from .core import NativeEffect, SmoothParam


class OctaverEffect(NativeEffect):
    """Two-grain delay-line pitch shifter with cubic Hermite taps and Hann grain
    windows on the mono mix, blended dry/wet into both channels."""

    def __init__(self, semitones=-12.0, mix=0.5, window_ms=40.0):
        self.semitones = SmoothParam(semitones, -24.0, 24.0)
        self.mix = SmoothParam(mix, 0.0, 1.0)
        self.window_ms = float(window_ms)
        self._fs = 48000
        self.w = 0
        self.phasor = 0.0
        self.size = 1

    def set_semitones(self, v): self.semitones.set_target(v)
    def set_mix(self, v): self.mix.set_target(v)

    def prepare(self, sample_rate: int, channels_in: int, channels_out: int, blocksize: int):
        self._sr = self._fs = sample_rate
        req = max(int(self._fs * self.window_ms / 1000.0), 16)
        if req != self.size:
            # only a ring-size change resets the write pointer and the phasor
            # (octaver.py:108-114): the warm-up phase leaks into the file otherwise
            self.size, self.w, self.phasor = req, 0, 0.0
            self._reset_lines()

    def _blob_floats(self):
        return self.size

    def _step(self, semi):
        return (1.0 - 2.0 ** (semi / 12.0)) / self.size          # octaver.py:121-122

    def _stages(self, frames):
        semi = self.semitones.step_towards(0.5)
        mix_now = self.mix.step_towards(0.05)
        self._step_now = self._step(semi)
        d = _native.StageDesc()
        d.kind = _native.OCTAVER
        d.q[0], d.q[1] = self.size, self.w
        d.p[0], d.p[1], d.p[2] = self.phasor, self._step_now, mix_now
        return [d]

    def _absorb(self, desc, frames, silent):
        # the streaming kernel iterates the reference's own phasor recurrence (octaver.py:77-80)
        self.w, self.phasor = int(desc.q[1]), float(desc.p[0])
        self._n_total += frames
        if not silent:
            self._dirty = True

    def _rest_block(self, frames):
        self._advance(frames, True)         # write pointer and phasor move, the ring stays zero

    def _advance(self, frames, silent, final=None):
        self._n_total += frames
        self.w = (self.w + frames) % self.size
        step, ph = self._step_now, self.phasor
        if frames <= 8192:            # the reference's running sum, bit for bit (octaver.py:77-80)
            for _ in range(frames):
                ph += step
                if ph >= 1.0:
                    ph -= 1.0
                elif ph < 0.0:
                    ph += 1.0
        else:
            ph = (ph + frames * step) % 1.0
        self.phasor = ph
        if not silent:
            self._stale = True
