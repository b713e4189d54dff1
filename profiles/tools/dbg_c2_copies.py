import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "audio-effects-simulator_b200"), os.path.join(ROOT, "tests"), ROOT]
import numpy as np, synth
from audioblocks.engine import file_chain
import importlib.util
spec = importlib.util.spec_from_file_location("tg", os.path.join(ROOT, "tests/test_gpu_parity.py")); tg = importlib.util.module_from_spec(spec); spec.loader.exec_module(tg)
cfg = tg.RACE_CHAINS["c2-biquads"]
n, B = 12000, 900
base = synth.batch(95, 2, n)
x = np.ascontiguousarray(base[np.arange(B) % 2])
for env in ({}, {"AES_NO_SCAN": "1"}):
    os.environ.pop("AES_NO_SCAN", None); os.environ.update(env)
    y = file_chain(cfg, 48000, channels_in=2).process_batch(x)
    for k in range(2):
        d = np.abs(y[k::2] - y[k]).max(axis=(1, 2))
        bad = np.nonzero(d > 0)[0]
        print(env, "clip", k, "copies differing:", bad.tolist()[:20], "max diff", d.max(), "first bad frame", [int(np.argmax(np.abs(y[k + 2 * b] - y[k]).max(axis=1) > 0)) for b in bad[:3]])
