"""IR-convolution reverb.  The reference has no such block (its ReverbEffect,
reverb.py:72-277, is a Schroeder network): this is a new operator for
BASELINE configs[3] in the style of the other blocks,
    out[:, c] = clip(mix_dry * x[:, c] + mix_wet * (x[:, c] (*) ir[:, c]), -1, 1),
evaluated as uniformly partitioned FFT convolution on the GPU (csrc/aes_convreverb.cuh).
It needs whole-clip FFT blocks, so it runs as its own kernels between fused chain
segments rather than inside the tile pipeline.  Every call convolves from silence (no tail
is carried between calls)."""
from __future__ import annotations

import numpy as np

from . import _native
from .core import Effect


class ConvolutionReverbEffect(Effect):
    def __init__(self, ir, mix_dry=0.7, mix_wet=0.5, block_log2=0):
        ir = np.asarray(ir, np.float32)
        if ir.ndim == 1:
            ir = np.stack([ir, ir], axis=1)
        if ir.ndim != 2 or ir.shape[1] != 2 or ir.shape[0] < 1:
            raise ValueError("ir must be (n_taps,) or (n_taps, 2)")
        self.ir = np.ascontiguousarray(ir)
        self.mix_dry = float(mix_dry)
        self.mix_wet = float(mix_wet)
        self._block_log2 = int(block_log2)
        self._plan = None

    def set_mix_dry(self, v): self.mix_dry = float(v)
    def set_mix_wet(self, v): self.mix_wet = float(v)

    def prepare(self, sample_rate: int, channels_in: int, channels_out: int, blocksize: int):
        pass

    def plan(self) -> _native.ConvReverbPlan:
        if self._plan is None:
            self._plan = _native.ConvReverbPlan(self.ir, self._block_log2)
        return self._plan

    def process_into(self, x_in: np.ndarray, out: np.ndarray) -> None:
        y = self.process_batch(np.ascontiguousarray(x_in, np.float32)[None])
        out[:, :] = y[0]

    def process_batch(self, x: np.ndarray, out: np.ndarray | None = None) -> np.ndarray:
        """x: (B, frames, 2) float32 host array -> (B, frames, 2) float32."""
        x = np.ascontiguousarray(x, np.float32)
        if x.ndim != 3 or x.shape[2] != 2:
            raise ValueError("ConvolutionReverbEffect processes (B, frames, 2) float32 batches")
        if out is None:
            out = np.empty_like(x)
        if x.shape[0] and x.shape[1]:
            self.plan().run_host(x, out, x.shape[0], x.shape[1], self.mix_dry, self.mix_wet)
        return out
