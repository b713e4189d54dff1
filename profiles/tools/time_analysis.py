#!/usr/bin/env python
"""Device time of the spectrum / chromagram kernel: N pairs of 16384-sample signals resident in HBM."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "audio-effects-simulator_b200"))
import torch
from audioblocks import _native

n_pairs, n_fft, fs = int(sys.argv[1]) if len(sys.argv) > 1 else 4096, 16384, 48000.0
L = _native.lib()
dev = torch.device("cuda:0")
a = torch.rand(n_pairs, n_fft, device=dev) - 0.5
b = torch.rand(n_pairs, n_fft, device=dev) - 0.5
nb = n_fft // 2 + 1
db = torch.empty(n_pairs, 2, nb, device=dev)
lin = torch.empty_like(db)
ch = torch.empty(n_pairs, 2, 12, device=dev)
pk = torch.empty(n_pairs, 2, device=dev)
st = torch.cuda.current_stream().cuda_stream


def run():
    _native.check(L.aes_spectrum_chroma(C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()), n_pairs, n_fft, n_fft, fs,
                                        C.c_void_p(db.data_ptr()), C.c_void_p(lin.data_ptr()), C.c_void_p(ch.data_ptr()),
                                        C.c_void_p(pk.data_ptr()), C.c_void_p(st)))


for _ in range(3):
    run()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
for _ in range(10):
    run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
bytes_alg = n_pairs * 2 * (n_fft * 4 + nb * 8 + 13 * 4)
print(json.dumps({"what": "spectrum + chromagram, 16384-point, pairs resident in HBM", "pairs": n_pairs, "ms": ms,
                  "signals_per_s": 2 * n_pairs / ms * 1e3, "Msamples_per_s": 2 * n_pairs * n_fft / ms / 1e3,
                  "algorithmic_GBps": bytes_alg / ms / 1e6}))
