"""Device time of the presets and a few single-purpose chains on B clips x 10 s resident in HBM, each checked against the
oracle on clip 0's first 96 000 frames:  python profiles/tools/time_presets.py [clips] [name-substring ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "audio-effects-simulator_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch
import synth
from audioblocks.engine import file_chain
from oracle import oracle as orc

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
want = sys.argv[2:]
n = 480000
x = (0.3 * torch.randn((B, n, 2), device="cuda")).clamp_(-1, 1)
y = torch.empty_like(x)
st = torch.cuda.current_stream()
D = lambda ms, off, fb: {"type": "delay", "params": {"delay_ms": ms, "feedback": fb, "offset_ms": off}}
chains = dict(synth.PRESETS)
chains.update({
    "c3 dist>octaver>delay": [{"type": "distortion", "params": {"drive": 4.0}},
                              {"type": "octaver", "params": {"semitones": -12, "mix": 0.5}}, D(120, 10, 0.3)],
    "delay 375 ms (aligned lags)": [D(375, 0, 0.2)],
    "delay 375.03 ms + 10.01 ms offset (lags 18001 / 18481)": [D(375.03, 10.01, 0.2)],
    "octaver -12": [{"type": "octaver", "params": {"semitones": -12, "mix": 0.5}}],
    "octaver +7": [{"type": "octaver", "params": {"semitones": 7, "mix": 0.5}}],
    "bare reverb": [{"type": "reverb", "params": {}}],
})
for name, cfg in chains.items():
    if want and not any(w.lower() in name.lower() for w in want):
        continue
    chain = file_chain(cfg, 48000, channels_in=2)
    pipe, plans = chain.device_pipeline(n)
    f = lambda: pipe(x.data_ptr(), y.data_ptr(), y.data_ptr(), B, st.cuda_stream)
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); f(); f(); f(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    nchk = 96000          # every block of the chains above is causal: a prefix of the clip has the clip's own output
    ref = orc.run_file_path(cfg, np.ascontiguousarray(x[0, :nchk].cpu().numpy()), 48000)
    err = float(np.max(np.abs(ref - y[0, :nchk].cpu().numpy())))
    print(f"{name:56s} {ms:8.2f} ms {B * n * 2 / ms / 1e3:9.0f} Msamples/s  frac {B * n * 2 * 8 / ms / 1e6 / 6547.8:.3f}  max-abs vs oracle {err:.2e}", flush=True)
    for p in plans:
        p.close()
