"""Small all-kernels workload (run on a GPU box; compute-sanitizer is closed on this pool, so this is a plain stress run):
    python tests/all_kernels_case.py
Pushes every preset (fast + generic kernels), the biquad scan, the convolution reverb and the
streaming kernel through short clips."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE, os.path.join(ROOT, "audio-effects-simulator_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np

import audioblocks as ab
import synth
from audioblocks.engine import file_chain

n = 5000
x = synth.batch(0, 3, n)
for generic in (False, True):
    if generic:
        os.environ["AES_NO_FAST"] = "1"
    for name, cfg in synth.PRESETS.items():
        y = file_chain(cfg, 48000, channels_in=2).process_batch(x)
        assert np.isfinite(y).all(), name
os.environ.pop("AES_NO_FAST", None)
cfg = [{"type": "filter", "params": {"filter_type": 0, "cutoff_hz": 500}}, {"type": "filter", "params": {"filter_type": 1, "cutoff_hz": 80}}]
one = synth.clip(1, 40000, 2)
out = np.zeros_like(one)
file_chain(cfg, 48000, channels_in=2).process(one, out)                     # biquad scan kernel
ir = np.random.default_rng(0).standard_normal((3000, 2)).astype(np.float32) * 0.01
ab.ConvolutionReverbEffect(ir, block_log2=11).process_batch(x)
ch = ab.EffectsChain(48000, 1, 2, 256)
for c in synth.PRESETS["Rain Delay"] + synth.PRESETS["Robot Voice"]:
    ch.add(ab.engine.make_effect(c))
ch.warmup()
blk = synth.clip(2, 256, 1)
o = np.zeros((256, 2), np.float32)
for _ in range(3):
    ch.process(blk, o)                                                       # streaming kernel
print("sanitizer case done")
