"""Seeded random chains (random blocks, parameters inside the reference's clamps, random lengths and
sample rates) through the kernels against the CPU oracle: on the CUDA-on-CPU emulator here, on the
GPU through the drop-in API with `-m gpu`."""
import numpy as np
import pytest

import emu
import synth
from oracle import oracle as orc

FP_TOL = 1e-5


def case(seed):
    rng = np.random.default_rng(7000 + seed)
    cfg = synth.random_chain(rng)
    fs = int(rng.choice([48000, 48000, 44100]))
    return cfg, fs, rng


def error_gain(cfg):
    """The 1e-5 bar is per block; a waveshaper multiplies whatever error reaches it by up to its
    slope at zero, (1 - mix) + mix * drive (two of them in series: up to 60x)."""
    g = 1.0
    for c in cfg:
        if c["type"] == "distortion":
            p = c["params"]
            g *= max(1.0, (1.0 - p["mix"]) + p["mix"] * p["drive"])
    return g


def check(got, want, what, cfg):
    mx, snr = synth.err_stats(got, want)
    scale = max(1.0, float(np.max(np.abs(want)))) * error_gain(cfg)
    assert mx <= FP_TOL * scale and snr >= 100.0, (what, mx, snr)


@pytest.mark.parametrize("seed", range(80))
def test_random_chain_on_the_emulator(seed):
    cfg, fs, rng = case(seed)
    n = int(rng.choice([700, 1024, 2500, 5000]))
    x = synth.clip(300 + seed, n, 2, fs)
    y = emu.run(emu.resolved_descs(cfg, fs, n, 2), fs, x[None])[0]
    check(y, orc.run_file_path(cfg, x, fs), (seed, cfg, fs, n), cfg)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(120))
def test_random_chain_on_the_gpu(seed):
    from audioblocks import _native
    from audioblocks.engine import file_chain
    _native.lib()
    cfg, fs, rng = case(seed)
    n = int(rng.choice([3000, 20000, 70000, 150000]))
    B = int(rng.choice([1, 3]))
    x = synth.batch(400 + seed, B, n, 2, fs)
    y = file_chain(cfg, fs, channels_in=2).process_batch(x)
    for b in range(B):
        check(y[b], orc.run_file_path(cfg, x[b], fs), (seed, b, cfg, fs, n), cfg)
