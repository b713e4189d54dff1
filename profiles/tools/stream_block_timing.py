#!/usr/bin/env python
"""Latency of the live route's block call (engine.py:156-163): EffectsChain.process on 256-frame
mono blocks (host buffers in and out, state carried on the device), per preset."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "audio-effects-simulator_b200"), os.path.join(ROOT, "tests")]
import numpy as np
import audioblocks as ab
import synth
for bs in (256, 1024):
    for name, cfg in synth.PRESETS.items():
        chain = ab.EffectsChain(48000, 1, 2, bs)
        for c in cfg:
            chain.add(ab.engine.make_effect(c))
        chain.warmup()
        x = synth.clip(3, bs * 60, 1)
        out = np.zeros((bs, 2), np.float32)
        ts = []
        for k in range(60):
            blk = np.ascontiguousarray(x[bs * k:bs * (k + 1)])
            t0 = time.perf_counter(); chain.process(blk, out); ts.append(time.perf_counter() - t0)
        ts = np.array(ts[10:]) * 1e3
        print(f"block {bs:5d}  {name:22s} median {np.median(ts):6.3f} ms   max {ts.max():6.3f} ms   (period {bs / 48.0:.2f} ms)")
