"""Device time of a few chains on B clips x 10 s resident in HBM:  python profiles/tools/time_chains.py [clips]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "audio-effects-simulator_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import synth
from audioblocks.engine import file_chain

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2368
n = 480000
x = (0.3 * torch.randn((B, n, 2), device="cuda")).clamp_(-1, 1)
y = torch.empty_like(x)
st = torch.cuda.current_stream()
F = lambda t, fc, q: {"type": "filter", "params": {"filter_type": t, "cutoff_hz": fc, "q": q}}
chains = {
    "filter (BP 800)": [F(2, 800, 0.8)],
    "filter x2": [F(0, 4000, 0.707), F(1, 120, 0.707)],
    "filter x3": [F(0, 4000, 0.707), F(1, 120, 0.707), F(2, 1000, 1.0)],
    "bare reverb (Guitar Filter's)": [synth.PRESETS["Guitar Filter"][1]],
    "Guitar Filter": synth.PRESETS["Guitar Filter"],
}
for name, cfg in chains.items():
    for env in ({}, {"AES_NO_BQSEQ": "1"}):
        if env and not any(c["type"] == "filter" for c in cfg):
            continue
        for k in ("AES_NO_BQSEQ",):
            os.environ.pop(k, None)
        os.environ.update(env)
        chain = file_chain(cfg, 48000, channels_in=2)
        pipe, plans = chain.device_pipeline(n)
        f = lambda: pipe(x.data_ptr(), y.data_ptr(), y.data_ptr(), B, st.cuda_stream)
        f(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); f(); f(); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print(f"{name:32s} {'old kernels' if env else 'default':12s} {ms:8.2f} ms {B * n * 2 / ms / 1e3:9.0f} Msamples/s", flush=True)
        for p in plans:
            p.close()
