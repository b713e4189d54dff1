import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "audio-effects-simulator_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch, synth
from audioblocks.engine import file_chain
B, n = 1184, 480000
x = (0.3 * torch.randn((B, n, 2), device="cuda")).clamp_(-1, 1)
y = torch.empty_like(x)
st = torch.cuda.current_stream()
R = lambda **kw: [{"type": "reverb", "params": kw}]
chains = {
    "Cathedral (rt60 4, damp .2, pre 20 ms)": R(rt60_s=4.0, mix_wet=0.6, mix_dry=0.6, damp=0.2, pre_delay_ms=20),
    "same, no pre-delay": R(rt60_s=4.0, mix_wet=0.6, mix_dry=0.6, damp=0.2, pre_delay_ms=0),
    "same, pre-delay 10 ms": R(rt60_s=4.0, mix_wet=0.6, mix_dry=0.6, damp=0.2, pre_delay_ms=10),
    "same, pre-delay 20.03 ms (odd lag)": R(rt60_s=4.0, mix_wet=0.6, mix_dry=0.6, damp=0.2, pre_delay_ms=20.03),
    "default reverb (damp .3)": R(),
    "damp 0.05": R(damp=0.05),
}
for name, cfg in chains.items():
    chain = file_chain(cfg, 48000, channels_in=2)
    pipe, plans = chain.device_pipeline(n)
    f = lambda: pipe(x.data_ptr(), y.data_ptr(), y.data_ptr(), B, st.cuda_stream)
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); f(); f(); f(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"{name:44s} {ms:8.2f} ms {B * n * 2 / ms / 1e3:9.0f} Msamples/s", flush=True)
    for p in plans: p.close()
