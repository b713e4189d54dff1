"""Time chains that have no kernel of their own: cut into runs of specialised kernels (default) against
the generic interpreter (AES_NO_SPLIT=1).   python profiles/tools/time_unlisted.py [clips]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "audio-effects-simulator_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import bench
from audioblocks.engine import file_chain

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
n = 480000
x = (0.3 * torch.randn((B, n, 2), device="cuda")).clamp_(-1, 1)
y = torch.empty_like(x)
st = torch.cuda.current_stream()
chains = dict(bench.UNLISTED_CHAINS)
chains["filter>filter"] = [{"type": "filter", "params": {"filter_type": 0, "cutoff_hz": 4000, "q": 0.707}},
                           {"type": "filter", "params": {"filter_type": 1, "cutoff_hz": 120, "q": 0.707}}]
chains["octaver>reverb"] = [{"type": "octaver", "params": {"semitones": 7, "mix": 0.4}},
                            {"type": "reverb", "params": {"rt60_s": 1.2, "mix_wet": 0.25, "mix_dry": 0.9}}]
for name, cfg in chains.items():
    for mode in ("split3", "split8", "whole"):
        os.environ.pop("AES_NO_SPLIT", None); os.environ.pop("AES_SPLIT_MAX", None)
        if mode == "whole":
            os.environ["AES_NO_SPLIT"] = "1"
        if mode == "split8":
            os.environ["AES_SPLIT_MAX"] = "8"
        chain = file_chain(cfg, 48000, channels_in=2)
        pipe, plans = chain.device_pipeline(n)
        f = lambda: pipe(x.data_ptr(), y.data_ptr(), y.data_ptr(), B, st.cuda_stream)
        f(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); f(); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 2
        print(f"{name:36s} {mode:7s} {ms:8.2f} ms {B * n * 2 / ms / 1e3:9.0f} Msamples/s  {plans[0].info()['kernel']}", flush=True)
        for p in plans:
            p.close()
