import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "audio-effects-simulator_b200"), os.path.join(ROOT, "tests")]
import numpy as np
from audioblocks import engine as eng, _native
import synth
n, fs = 892775, 48000
audio = synth.clip(77, n, 2, fs)
mono = np.ascontiguousarray(eng.mono_downmix(audio), np.float32)
cfg = synth.PRESETS["Rain Delay"]
chain = eng.file_chain(cfg, fs, 1)
descs = chain.stage_descs(n) if hasattr(chain, "stage_descs") else None
out = np.zeros((n, 2), np.float32)
import subprocess
def clocks():
    return subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,pstate", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
for rep in range(5):
    time.sleep(1.0)
    c0 = clocks()
    t0 = time.perf_counter()
    plan = _native.ChainPlan(descs, fs); t1 = time.perf_counter()
    plan.run_host(mono, _native.FMT_F32_MONO, out, _native.FMT_F32_STEREO, 1, n); t2 = time.perf_counter()
    plan.run_host(mono, _native.FMT_F32_MONO, out, _native.FMT_F32_STEREO, 1, n); t3 = time.perf_counter()
    plan.close(); t4 = time.perf_counter()
    print(rep, c0, "| create %.1f ms, run1 %.1f ms, run2 %.1f ms, close %.1f ms" % ((t1-t0)*1e3, (t2-t1)*1e3, (t3-t2)*1e3, (t4-t3)*1e3))
