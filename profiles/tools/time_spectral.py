"""Time the whole-file SpectralFilter ('Clean Noise Removal') over a batch of clips for several
work-buffer chunk sizes (AES_SPECTRAL_CHUNK_MB) and for the Bluestein path (AES_SPECTRAL_BLUESTEIN=1).
    python profiles/tools/time_spectral.py [clips] [seconds]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "audio-effects-simulator_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import synth
from audioblocks.engine import file_chain

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
secs = float(sys.argv[2]) if len(sys.argv) > 2 else 10.0
n = int(48000 * secs)
x = (0.3 * torch.randn((B, n, 2), device="cuda")).clamp_(-1, 1)
y = torch.empty_like(x)
st = torch.cuda.current_stream()


def run_case(label):
    chain = file_chain(synth.PRESETS["Clean Noise Removal"], 48000, channels_in=2)
    pipe, plans = chain.device_pipeline(n)
    tmp = torch.empty_like(y) if pipe.n_segments > 1 else y
    f = lambda: pipe(x.data_ptr(), y.data_ptr(), tmp.data_ptr(), B, st.cuda_stream)
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); f(); f(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 2
    print(f"{label:28s} {ms:9.2f} ms  {B * n * 2 / ms / 1e3:9.0f} Msamples/s  checksum {float(y.double().abs().sum()):.6e}", flush=True)
    for p in plans:
        p.close()


for mb in (int(v) for v in os.environ.get("CHUNKS_MB", "64,8192").split(",")):
    os.environ["AES_SPECTRAL_CHUNK_MB"] = str(mb)
    run_case(f"smooth chunk {mb} MB")
os.environ.pop("AES_SPECTRAL_CHUNK_MB")
os.environ["AES_SPECTRAL_BLUESTEIN"] = "1"
run_case("bluestein")
