import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "audio-effects-simulator_b200"), os.path.join(ROOT, "tests")]
import numpy as np
from audioblocks import engine as eng
import synth
n, fs = 892775, 48000
audio = synth.clip(77, n, 2, fs)
mono = np.ascontiguousarray(eng.mono_downmix(audio), np.float32)
for name in ("Clean Noise Removal", "Robot Voice", "Guitar Filter"):
    cfg = synth.PRESETS[name]
    for rep in range(4):
        t0 = time.perf_counter()
        chain = eng.file_chain(cfg, fs, 1); t1 = time.perf_counter()
        out = np.zeros((n, 2), np.float32)
        chain.process(mono, out); t2 = time.perf_counter()
        del chain; t3 = time.perf_counter()
        print(name, rep, "build+warmup %.1f ms, process %.1f ms, drop %.1f ms" % ((t1-t0)*1e3, (t2-t1)*1e3, (t3-t2)*1e3))
import cProfile, pstats
cfg = synth.PRESETS["Clean Noise Removal"]
pr = cProfile.Profile(); pr.enable()
chain = eng.file_chain(cfg, fs, 1); out = np.zeros((n, 2), np.float32); chain.process(mono, out)
pr.disable(); pstats.Stats(pr).sort_stats("tottime").print_stats(12)
