import sys, time, os, queue
ROOT = os.environ.get("GRAFT_REPO_ROOT", "/root/repo")
sys.path[:0] = [os.path.join(ROOT, "audio-effects-simulator_b200"), os.path.join(ROOT, "tests")]
import numpy as np
import audioblocks as ab
from audioblocks import engine as eng
import synth
n, fs = 892775, 48000
audio = synth.clip(77, n, 2, fs)
cfg = synth.PRESETS["Rain Delay"]
for rep in range(3):
    t = [time.perf_counter()]
    mono = audio.mean(axis=1, keepdims=True); t.append(time.perf_counter())
    chain = eng.file_chain(cfg, fs, 1); t.append(time.perf_counter())
    processed = np.zeros((len(mono), 2), dtype=np.float32); t.append(time.perf_counter())
    chain.process(np.ascontiguousarray(mono, np.float32), processed); t.append(time.perf_counter())
    processed = np.clip(processed, -1.0, 1.0); t.append(time.perf_counter())
    pcm = (processed * 32767).astype(np.int16); t.append(time.perf_counter())
    print(rep, dict(zip(["mean", "file_chain(build+warmup)", "zeros", "process", "clip", "quantise"], (np.diff(t) * 1e3).round(2))))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
chain = eng.file_chain(cfg, fs, 1); processed = np.zeros((len(mono), 2), dtype=np.float32); chain.process(np.ascontiguousarray(mono, np.float32), processed)
pr.disable(); pstats.Stats(pr).sort_stats("cumulative").print_stats(25)
