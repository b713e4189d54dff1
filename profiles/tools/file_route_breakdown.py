#!/usr/bin/env python
"""Where one process_file request spends its chain time when the GPU has been idle (as it is
between user requests): the numeric core of engine.process_file_arrays, piece by piece, each
repetition after a 1 s pause.  usage: python profiles/tools/file_route_breakdown.py"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "audio-effects-simulator_b200"), os.path.join(ROOT, "tests")]
import numpy as np                                        # noqa: E402
from audioblocks import engine as eng                     # noqa: E402
import synth                                              # noqa: E402

n, fs = 892775, 48000
audio = synth.clip(77, n, 2, fs)
cfg = synth.PRESETS["Rain Delay"]
for rep in range(4):
    time.sleep(1.0)
    t = [time.perf_counter()]
    mono = eng.mono_downmix(audio); t.append(time.perf_counter())
    chain = eng.file_chain(cfg, fs, 1); t.append(time.perf_counter())
    processed = np.zeros((len(mono), 2), dtype=np.float32); t.append(time.perf_counter())
    chain.process(np.ascontiguousarray(mono, np.float32), processed); t.append(time.perf_counter())
    processed = np.clip(processed, -1.0, 1.0); t.append(time.perf_counter())
    pcm = (processed * 32767).astype(np.int16); t.append(time.perf_counter())
    names = ["downmix", "file_chain (build + warm-up)", "zeros", "chain.process (H2D + kernel + D2H)", "clip", "quantise"]
    print(rep, {k: round(float(v), 2) for k, v in zip(names, np.diff(t) * 1e3)})
import cProfile
import pstats
time.sleep(1.0)
pr = cProfile.Profile(); pr.enable()
chain = eng.file_chain(cfg, fs, 1); processed = np.zeros((len(mono), 2), dtype=np.float32); chain.process(np.ascontiguousarray(mono, np.float32), processed)
pr.disable(); pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
