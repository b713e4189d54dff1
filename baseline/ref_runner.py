"""Runs the UNMODIFIED reference package (javierdrp/audio-effects-simulator, src/audioblocks) on host
cores for bench.py's reference arm / cpu_baseline.  TEST / BENCH INFRASTRUCTURE, never imported by the
product.

The reference is pure Python + numba and has no installer; `__graft_entry__.build()` copies
`/root/reference/src/audioblocks` verbatim to `baseline/_ref/src/audioblocks` (git-ignored, shipped to
the GPU box with the working tree).  This module puts that directory first on `sys.path`, shims the
absent `soundfile` import (engine.py:7 -- the WAV decoder is not on the timed path) and replays the
reference's own file-route protocol (engine.py:86-102): EffectsChain(fs, ci, 2, 1024) -> add(Class(**params))
-> warmup() -> ONE process() call over the whole clip.  The numba kernels are single-threaded
(no parallel=True anywhere), so throughput on a box is one clip per worker process.
"""
from __future__ import annotations

import os
import sys
import time
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.path.join(HERE, "_ref", "src")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_SRC, "audioblocks", "core.py"))


_ab = None


def _ref():
    """Import the reference package (and nothing of this repository's `audioblocks`)."""
    global _ab
    if _ab is None:
        os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache_ref")
        sys.modules.setdefault("soundfile", types.SimpleNamespace(read=None))
        if "audioblocks" in sys.modules and not os.path.realpath(sys.modules["audioblocks"].__file__).startswith(os.path.realpath(REF_SRC)):
            raise RuntimeError("the product's audioblocks is already imported in this process")
        sys.path.insert(0, REF_SRC)
        import audioblocks as ab
        assert os.path.realpath(ab.__file__).startswith(os.path.realpath(REF_SRC)), ab.__file__
        _ab = ab
    return _ab


def classes():
    ab = _ref()
    return {"delay": ab.StereoDelayEffect, "reverb": ab.ReverbEffect, "gate": ab.NoiseGateEffect,
            "spectral": ab.SpectralFilter, "octaver": ab.OctaverEffect, "filter": ab.FilterEffect}


def supports(config) -> bool:
    """Chains made only of blocks the reference has (no distortion / peaking / IR-convolution extension)."""
    for c in config:
        if c["type"] not in ("delay", "reverb", "gate", "spectral", "octaver", "filter"):
            return False
        if c["type"] == "filter" and (int(round(c.get("params", {}).get("filter_type", 0))) > 2 or "gain_db" in c.get("params", {})):
            return False
    return True


def run_file_path(config, x, fs):
    """engine.py:86-102 against the reference classes; x: (N, 1|2) float32 -> (N, 2) float32."""
    import numpy as np
    ab = _ref()
    cls = classes()
    chain = ab.EffectsChain(fs, x.shape[1], 2, 1024)
    for cfg in config:
        chain.add(cls[cfg["type"]](**cfg.get("params", {})))
    chain.warmup()
    out = np.zeros((x.shape[0], 2), np.float32)
    chain.process(x, out)
    return out


# ---- worker pool (plain subprocesses: one warmed reference chain per host thread) -------------------
class Pool:
    """`workers` processes, each holding the reference package, JIT-warmed, and two synthetic clips.
    run(k) makes every worker process k whole clips and returns (Msamples/s, wall seconds)."""

    def __init__(self, config, fs, n_frames, workers):
        import json
        import subprocess
        self.workers, self.n_frames = workers, n_frames
        arg = json.dumps({"config": config, "fs": fs, "n_frames": n_frames})
        self.procs = [subprocess.Popen([sys.executable, os.path.abspath(__file__), "--worker", arg, str(w)],
                                       stdin=subprocess.PIPE, stdout=subprocess.PIPE, text=True, bufsize=1)
                      for w in range(workers)]
        for p in self.procs:                                        # every worker is up and warm
            line = p.stdout.readline()
            if not line.startswith("ready"):
                raise RuntimeError("reference worker failed to start: " + line)

    def run(self, clips_per_worker):
        t0 = time.perf_counter()
        for p in self.procs:
            p.stdin.write(f"{clips_per_worker}\n")
            p.stdin.flush()
        for p in self.procs:
            line = p.stdout.readline()
            if not line.startswith("done"):
                raise RuntimeError("reference worker died: " + line)
        dt = time.perf_counter() - t0
        return self.workers * clips_per_worker * self.n_frames * 2 / dt / 1e6, dt

    def close(self):
        for p in self.procs:
            try:
                p.stdin.write("0\n")
                p.stdin.flush()
                p.stdin.close()
            except OSError:
                pass
        for p in self.procs:
            try:
                p.wait(timeout=10)
            except Exception:
                p.kill()


def _worker_main(arg, index):
    import json
    import numpy as np
    a = json.loads(arg)
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), "tests"))
    import synth
    x = [np.ascontiguousarray(synth.clip(int(index) * 2 + k, a["n_frames"], 2, a["fs"])) for k in range(2)]
    run_file_path(a["config"], x[0][: min(a["n_frames"], 4096)], a["fs"])        # JIT / numba cache warm-up
    print("ready", flush=True)
    for line in sys.stdin:
        n = int(line)
        if n <= 0:
            break
        for k in range(n):
            run_file_path(a["config"], x[k & 1], a["fs"])
        print("done", flush=True)


if __name__ == "__main__":
    if len(sys.argv) >= 4 and sys.argv[1] == "--worker":
        _worker_main(sys.argv[2], sys.argv[3])
