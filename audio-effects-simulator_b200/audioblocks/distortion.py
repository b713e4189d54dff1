"""Waveshaper distortion.  The reference has no such block (README.md:7 only names
it as a goal): this is our definition, in the style of the others:
out = clip((1-mix)*x + mix*tanh(drive*x), -1, 1), float32 elementwise."""
from __future__ import annotations

from . import _native
from .core import NativeEffect, SmoothParam


class DistortionEffect(NativeEffect):
    def __init__(self, drive=4.0, mix=1.0):
        self.drive = SmoothParam(drive, 0.0, 100.0)
        self.mix = SmoothParam(mix, 0.0, 1.0)

    def set_drive(self, v): self.drive.set_target(v)
    def set_mix(self, v): self.mix.set_target(v)

    def prepare(self, sample_rate: int, channels_in: int, channels_out: int, blocksize: int):
        self._sr = sample_rate

    def _stages(self, frames):
        d = _native.StageDesc()
        d.kind = _native.DISTORTION
        d.p[0] = self.drive.step_towards(0.5)
        d.p[1] = self.mix.step_towards(0.05)
        return [d]

    def _advance(self, frames, silent, final=None):
        pass                      # memoryless

    def _absorb(self, desc, frames, silent):
        pass

    def _at_rest(self):
        return True

    def _rest_block(self, frames):
        pass
