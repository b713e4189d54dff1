"""Seeded synthetic clips shared by the tests, the golden generator and bench.py
(SURVEY.md section 8d): white noise + a tone + periodic full-scale bursts so the
clippers and the gate both switch.  float32, frame-major (N, C)."""
from __future__ import annotations

import numpy as np

FS = 48000


def clip_channel(seed: int, n: int, tone_idx: int, fs: int = FS) -> np.ndarray:
    rng = np.random.default_rng(seed)
    t = np.arange(n, dtype=np.float64) / fs
    f = 110.0 * 2.0 ** ((tone_idx % 48) / 12.0)
    x = 0.25 * rng.uniform(-1.0, 1.0, n) + 0.35 * np.sin(2.0 * np.pi * f * t)
    burst = (np.arange(n) % (2 * fs)) < int(0.05 * fs)          # 50 ms every 2 s
    x = np.where(burst, 0.95 * np.sign(np.sin(2.0 * np.pi * 997.0 * t) + 1e-12), x)
    return x.astype(np.float32)


def clip(i: int, n: int, channels: int = 2, fs: int = FS) -> np.ndarray:
    """Clip `i` of the synthetic set: (n, channels) f32; R uses seed + 500000."""
    cols = [clip_channel(1000 + i + 500000 * c, n, i, fs) for c in range(channels)]
    return np.ascontiguousarray(np.stack(cols, axis=1))


def batch(first: int, count: int, n: int, channels: int = 2, fs: int = FS) -> np.ndarray:
    return np.stack([clip(first + k, n, channels, fs) for k in range(count)], axis=0)


# app.py:41-71 DEFAULT_PRESETS, as data (parameters verbatim).
PRESETS = {
    "Robot Voice": [
        {"type": "gate", "params": {"threshold_db": -30, "attack_ms": 10, "release_ms": 100}},
        {"type": "octaver", "params": {"semitones": -12, "mix": 1.0}},
        {"type": "delay", "params": {"delay_ms": 120, "feedback": 0.3, "mix_wet": 0.3, "mix_dry": 1.0, "offset_ms": 10}},
    ],
    "Cathedral": [
        {"type": "reverb", "params": {"rt60_s": 4.0, "mix_wet": 0.6, "mix_dry": 0.6, "damp": 0.2, "pre_delay_ms": 20}},
    ],
    "Slapback Echo": [
        {"type": "delay", "params": {"delay_ms": 100, "feedback": 0.0, "mix_wet": 0.5, "mix_dry": 1.0, "offset_ms": 0}},
    ],
    "Clean Noise Removal": [
        {"type": "spectral", "params": {"threshold_db": -50, "reduction": 0.1}},
        {"type": "gate", "params": {"threshold_db": -40, "attack_ms": 5, "release_ms": 200}},
    ],
    "Guitar Filter": [
        {"type": "filter", "params": {"filter_type": 2, "cutoff_hz": 800, "q": 0.8}},
        {"type": "reverb", "params": {"mix_wet": 0.2, "rt60_s": 1.0}},
    ],
    "Rain Delay": [
        {"type": "delay", "params": {"feedback": 0.2, "delay_ms": 375, "mix_dry": 1, "mix_wet": 1, "offset_ms": 0}},
        {"type": "reverb", "params": {"rt60_s": 2.1, "mix_wet": 0.4, "mix_dry": 0.8, "damp": 0.05, "pre_delay_ms": 0}},
    ],
}


def err_stats(got: np.ndarray, want: np.ndarray):
    """(max-abs error, SNR in dB) of `got` against `want`."""
    d = got.astype(np.float64) - want.astype(np.float64)
    mx = float(np.max(np.abs(d))) if d.size else 0.0
    p_sig = float(np.sum(want.astype(np.float64) ** 2))
    p_err = float(np.sum(d ** 2))
    snr = float("inf") if p_err == 0.0 else 10.0 * np.log10(max(p_sig, 1e-300) / p_err)
    return mx, snr


# parameter corners shared by the emulator (CPU) and the GPU parity tests
REVERB_VARIANTS = {
    "heavy-damping": {"damp": 0.9, "rt60_s": 4.0},                          # slow one-pole: full scan depth, cross-warp carries
    "max-damping": {"damp": 0.99, "rt60_s": 10.0},
    "long-predelay": {"pre_delay_ms": 60.0},                                # 2880 samples >= tile: register line
    "short-predelay": {"pre_delay_ms": 3.0},                                # 144 samples: phase walk
    "other-combs": {"comb_times_ms": (31.3, 36.0, 40.7, 45.1), "allpass_times_ms": (6.1, 2.3), "jitter_ms": 0.7},
    "long-tail": {"rt60_s": 9.5, "mix_wet": 0.9, "mix_dry": 0.2},           # drives the output into the clipper
}

DELAY_VARIANTS = {
    "max-feedback": {"delay_ms": 333.3, "feedback": 0.95, "offset_ms": 17.7},       # odd lags, long tail
    "short": {"delay_ms": 1.0, "feedback": 0.5, "offset_ms": 0.0},                  # 48 samples: phase walk
    "mid": {"delay_ms": 30.0, "feedback": 0.6, "offset_ms": 5.5},                   # 1440 / 1704: register line in smem
    "longest": {"delay_ms": 1499.0, "feedback": 0.3, "offset_ms": 30.0},            # clamps at max_delay - 1
}

OCTAVER_VARIANTS = {
    "octave-down": {"semitones": -12.0, "mix": 0.5},
    "fifth-up": {"semitones": 7.0, "mix": 1.0},                  # negative phasor step
    "odd-window": {"semitones": -5.0, "mix": 0.7, "window_ms": 30.03},   # ring of 1441 samples: grains not an integer apart
    "two-octaves-up": {"semitones": 24.0, "mix": 0.3, "window_ms": 10.0},
}


def random_chain(rng, max_blocks: int = 4):
    """A random chain config (reference parameter names, values inside the SmoothParam clamps of
    SURVEY 8b) for the fuzz tests; no spectral block (whole-file FFT, tested separately)."""
    def delay():
        return {"type": "delay", "params": {"delay_ms": float(rng.choice([1.0, 7.3, 21.4, 120.0, 375.0, 900.0])),
                                            "feedback": float(rng.uniform(0.0, 0.95)),
                                            "offset_ms": float(rng.choice([0.0, 3.3, 30.0])),
                                            "mix_dry": float(rng.uniform(0.2, 1.0)), "mix_wet": float(rng.uniform(0.0, 1.0))}}

    def reverb():
        p = {"rt60_s": float(rng.uniform(0.2, 8.0)), "damp": float(rng.uniform(0.0, 0.95)),
             "pre_delay_ms": float(rng.choice([0.0, 0.0, 2.0, 20.0, 45.0, 100.0])),
             "mix_dry": float(rng.uniform(0.2, 1.0)), "mix_wet": float(rng.uniform(0.0, 0.8))}
        if rng.random() < 0.3:
            p["comb_times_ms"] = tuple(float(v) for v in np.sort(rng.uniform(30.0, 48.0, 4)))
            p["allpass_times_ms"] = (float(rng.uniform(3.0, 8.0)), float(rng.uniform(1.0, 2.9)))
        return {"type": "reverb", "params": p}

    def filt():
        return {"type": "filter", "params": {"filter_type": int(rng.integers(0, 3)),
                                             "cutoff_hz": float(np.exp(rng.uniform(np.log(30.0), np.log(15000.0)))),
                                             "q": float(rng.uniform(0.3, 6.0))}}

    def octaver():
        return {"type": "octaver", "params": {"semitones": float(rng.choice([-24, -12, -7, -5, 3, 7, 12])),
                                              "mix": float(rng.uniform(0.1, 1.0))}}

    def gate():
        return {"type": "gate", "params": {"threshold_db": float(rng.uniform(-60.0, -10.0)),
                                           "attack_ms": float(rng.uniform(1.0, 50.0)),
                                           "release_ms": float(rng.uniform(10.0, 400.0))}}

    def distortion():
        return {"type": "distortion", "params": {"drive": float(rng.uniform(0.5, 8.0)), "mix": float(rng.uniform(0.1, 1.0))}}

    makers = [delay, reverb, filt, octaver, gate, distortion]
    n = int(rng.integers(1, max_blocks + 1))
    chain, have_reverb = [], False
    for _ in range(n):
        mk = makers[int(rng.integers(0, len(makers)))]
        if mk is reverb and have_reverb:
            mk = filt                                   # one reverb per chain keeps shared memory within one CTA's budget
        have_reverb |= mk is reverb
        chain.append(mk())
    return chain
