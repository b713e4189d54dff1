"""Generate the golden vectors under tests/golden/ by RUNNING THE REFERENCE.

Runs only in the build container, where the unmodified Python reference is
mounted read-only at /root/reference (it cannot travel to the GPU box).  It
imports the reference's own `audioblocks` package (with a 3-line `soundfile`
shim, because audioblocks/__init__.py imports engine.py which imports
soundfile, absent here), drives

  (k) the seven numba kernels directly on seeded random inputs,
  (b) single blocks through the reference EffectsChain with the file-path
      protocol of engine.py:86-102 (build@1024 -> warmup -> one whole call),
  (p) every DEFAULT_PRESET (app.py:41-71) on an excerpt of music/rain-raw.wav
      (mono, as engine.py:81-84 feeds it) and on a synthetic stereo clip,

and stores inputs (or their seeds) and the reference outputs as .npz.
Nothing from the reference's sources is copied into this repository.

    python tests/golden/make_golden.py
"""
from __future__ import annotations

import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("AES_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(ROOT, "tests"))

os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache_ref")
sys.modules.setdefault("soundfile", types.SimpleNamespace(read=None))
sys.path.insert(0, os.path.join(REF, "src"))

import audioblocks as ab                      # noqa: E402  (the REFERENCE package)
from audioblocks import delay as r_delay, filter as r_filter, octaver as r_oct  # noqa: E402
from audioblocks import reverb as r_rev, gate as r_gate                         # noqa: E402
import synth                                   # noqa: E402

assert os.path.realpath(ab.__file__).startswith(os.path.realpath(REF)), ab.__file__

CLASSES = {"delay": ab.StereoDelayEffect, "reverb": ab.ReverbEffect, "gate": ab.NoiseGateEffect,
           "spectral": ab.SpectralFilter, "octaver": ab.OctaverEffect, "filter": ab.FilterEffect}


def ref_file_path(config, x, fs):
    """engine.py:86-102 replayed against the reference classes."""
    chain = ab.EffectsChain(fs, x.shape[1], 2, 1024)
    for cfg in config:
        chain.add(CLASSES[cfg["type"]](**cfg.get("params", {})))
    chain.warmup()
    out = np.zeros((x.shape[0], 2), np.float32)
    chain.process(x, out)
    return out


def kernel_cases():
    rng = np.random.default_rng(20261018)
    out = {}
    meta = {}
    N = 6000

    def sig(n=N, ch=1, amp=0.8):
        return (amp * rng.uniform(-1, 1, (n, ch))).astype(np.float32)

    # delay_kernel (delay.py:7-22): ring index arithmetic incl. wrap and dS clamp
    for i, (size, dS, fb, w0) in enumerate([(1001, 1000, 0.5, 0), (1001, 37, 0.95, 998),
                                            (257, 1, 0.2, 5), (4801, 4800, 0.0, 17), (64, 0, 0.3, 3)]):
        x = sig(); buf = (0.1 * rng.standard_normal(size)).astype(np.float32)
        b = buf.copy(); wet = np.zeros((N, 1), np.float32)
        w1 = r_delay.delay_kernel(b, w0, size, x, wet, dS, fb)
        out[f"delay{i}_x"], out[f"delay{i}_buf0"] = x, buf
        out[f"delay{i}_wet"], out[f"delay{i}_buf1"] = wet, b
        meta[f"delay{i}"] = dict(size=size, dS=dS, fb=fb, w0=w0, w1=int(w1))

    # biquad_kernel (filter.py:8-40) with a non-zero f32 state
    for i, (t, fc, q) in enumerate([(0, 1000.0, 0.707), (1, 80.0, 0.707), (2, 800.0, 0.8),
                                    (0, 20.0, 10.0), (1, 20.0, 0.707)]):
        fx = ab.FilterEffect(t, fc, q); fx.prepare(48000, 2, 2, N)
        co = fx._calc_coeffs(t, fc, q)
        x = sig(ch=2); st = (0.05 * rng.standard_normal((2, 4))).astype(np.float32)
        s1 = st.copy(); y = np.zeros_like(x)
        r_filter.biquad_kernel(x, y, *co, s1)
        out[f"biquad{i}_x"], out[f"biquad{i}_st0"], out[f"biquad{i}_y"], out[f"biquad{i}_st1"] = x, st, y, s1
        meta[f"biquad{i}"] = dict(coeffs=[float(c) for c in co], t=t, fc=fc, q=q)

    # pitch_shift_kernel_cubic (octaver.py:17-82): down, up, tiny ring
    for i, (size, semi, w0, ph0) in enumerate([(1920, -12.0, 128, 0.5333333333333231), (1920, 7.0, 0, 0.0),
                                               (16, -5.0, 3, 0.25), (2116, 12.0, 2000, 0.9)]):
        step = (1.0 - 2.0 ** (semi / 12.0)) / size
        x = sig(); buf = (0.1 * rng.standard_normal(size)).astype(np.float32)
        b = buf.copy(); y = np.zeros((N, 1), np.float32)
        w1, ph1 = r_oct.pitch_shift_kernel_cubic(b, w0, size, x, y, ph0, step)
        out[f"pitch{i}_x"], out[f"pitch{i}_buf0"], out[f"pitch{i}_y"], out[f"pitch{i}_buf1"] = x, buf, y, b
        meta[f"pitch{i}"] = dict(size=size, step=step, w0=w0, ph0=ph0, w1=int(w1), ph1=float(ph1))

    # reverb kernels (reverb.py:11-67)
    for i, (size, dS, w0) in enumerate([(4801, 960, 0), (4801, 0, 11), (100, 99, 50)]):
        x = sig(); buf = (0.1 * rng.standard_normal(size)).astype(np.float32)
        b = buf.copy(); y = np.zeros((N, 1), np.float32)
        w1 = r_rev.pure_delay_kernel(b, w0, size, x, y, dS)
        out[f"pure{i}_x"], out[f"pure{i}_buf0"], out[f"pure{i}_y"], out[f"pure{i}_buf1"] = x, buf, y, b
        meta[f"pure{i}"] = dict(size=size, dS=dS, w0=w0, w1=int(w1))
    for i, (L, g, h, w0, lp0) in enumerate([(1440, 0.95, 0.2, 0, 0.0), (1411, 0.9796, 0.05, 700, 0.01),
                                            (50, 0.7, 0.99, 3, -0.2), (1, 0.5, 0.3, 0, 0.0)]):
        x = sig(amp=0.3); buf = (0.1 * rng.standard_normal(L + 1)).astype(np.float32)
        b = buf.copy(); y = np.zeros((N, 1), np.float32)
        w1, lp1 = r_rev.comb_damped_kernel(b, w0, L + 1, x, y, L, g, h, lp0)
        out[f"comb{i}_x"], out[f"comb{i}_buf0"], out[f"comb{i}_y"], out[f"comb{i}_buf1"] = x, buf, y, b
        meta[f"comb{i}"] = dict(L=L, g=g, h=h, w0=w0, lp0=lp0, w1=int(w1), lp1=float(lp1))
    for i, (L, a, w0) in enumerate([(242, 0.6, 0), (78, 0.6, 40), (1, 0.7, 0)]):
        x = sig(); buf = (0.1 * rng.standard_normal(L + 1)).astype(np.float32)
        b = buf.copy(); y = np.zeros((N, 1), np.float32)
        w1 = r_rev.allpass_kernel(b, w0, L + 1, x, y, L, a)
        out[f"ap{i}_x"], out[f"ap{i}_buf0"], out[f"ap{i}_y"], out[f"ap{i}_buf1"] = x, buf, y, b
        meta[f"ap{i}"] = dict(L=L, a=a, w0=w0, w1=int(w1))

    # gate_kernel (gate.py:6-42): signal straddling the threshold
    for i, (thr_db, att_ms, rel_ms, g0) in enumerate([(-30.0, 10.0, 100.0, 0.0), (-12.0, 1.0, 1000.0, 0.7)]):
        fx = ab.NoiseGateEffect(thr_db, att_ms, rel_ms); fx.prepare(48000, 2, 2, N)
        thr, att, rel = 10.0 ** (thr_db / 20.0), fx._calc_coeff(att_ms), fx._calc_coeff(rel_ms)
        env = (np.sin(np.arange(N) * 2 * np.pi / 1500.0) > 0.3).astype(np.float32)[:, None]
        x = (sig(ch=2) * (0.02 + 0.9 * env)).astype(np.float32)
        y = np.zeros_like(x)
        g1 = r_gate.gate_kernel(x, y, g0, thr, att, rel)
        out[f"gate{i}_x"], out[f"gate{i}_y"] = x, y
        meta[f"gate{i}"] = dict(thr=float(thr), att=float(att), rel=float(rel), g0=g0, g1=float(g1))

    out["meta"] = np.frombuffer(json.dumps(meta).encode(), np.uint8)
    np.savez_compressed(os.path.join(HERE, "kernels.npz"), **out)
    print("kernels.npz:", len(meta), "cases")


BLOCK_CASES = {
    # name: (config, channels_in)
    "delay_default":   ([{"type": "delay", "params": {}}], 2),
    "delay_fb0":       ([{"type": "delay", "params": {"delay_ms": 100, "feedback": 0.0, "mix_wet": 0.5, "mix_dry": 1.0, "offset_ms": 0}}], 1),
    "delay_fb95":      ([{"type": "delay", "params": {"delay_ms": 33.3, "feedback": 0.95, "offset_ms": 7.7}}], 2),
    "delay_short":     ([{"type": "delay", "params": {"delay_ms": 1.0, "feedback": 0.5, "offset_ms": 0.5}}], 2),
    "delay_clamp":     ([{"type": "delay", "params": {"delay_ms": 5000.0, "feedback": 0.4, "max_delay_ms": 200.0}}], 2),
    "filter_lp":       ([{"type": "filter", "params": {"filter_type": 0, "cutoff_hz": 8000, "q": 0.707}}], 2),
    "filter_hp":       ([{"type": "filter", "params": {"filter_type": 1, "cutoff_hz": 80, "q": 0.707}}], 2),
    "filter_bp":       ([{"type": "filter", "params": {"filter_type": 2, "cutoff_hz": 1000, "q": 0.8}}], 1),
    "filter_lp40q5":   ([{"type": "filter", "params": {"filter_type": 0, "cutoff_hz": 40, "q": 5.0}}], 2),
    "filter_hp20":     ([{"type": "filter", "params": {"filter_type": 1, "cutoff_hz": 20, "q": 0.707}}], 2),
    "filter_cascade3": ([{"type": "filter", "params": {"filter_type": 0, "cutoff_hz": 8000, "q": 0.707}},
                         {"type": "filter", "params": {"filter_type": 1, "cutoff_hz": 80, "q": 0.707}},
                         {"type": "filter", "params": {"filter_type": 2, "cutoff_hz": 1000, "q": 0.8}}], 2),
    "octaver_down":    ([{"type": "octaver", "params": {"semitones": -12, "mix": 0.5}}], 2),
    "octaver_up7":     ([{"type": "octaver", "params": {"semitones": 7, "mix": 1.0}}], 1),
    "octaver_win10":   ([{"type": "octaver", "params": {"semitones": -24, "mix": 0.8, "window_ms": 10.0}}], 2),
    "reverb_default":  ([{"type": "reverb", "params": {}}], 2),
    "reverb_cathedral": ([{"type": "reverb", "params": {"rt60_s": 4.0, "mix_wet": 0.6, "mix_dry": 0.6, "damp": 0.2, "pre_delay_ms": 20}}], 1),
    "reverb_rt10":     ([{"type": "reverb", "params": {"rt60_s": 10.0, "damp": 0.0, "mix_wet": 0.3}}], 2),
    "reverb_damp99":   ([{"type": "reverb", "params": {"rt60_s": 2.0, "damp": 0.99, "pre_delay_ms": 100.0}}], 2),
    "gate_default":    ([{"type": "gate", "params": {}}], 2),
    "gate_fast":       ([{"type": "gate", "params": {"threshold_db": -10, "attack_ms": 1, "release_ms": 10}}], 2),
    "gate_slow":       ([{"type": "gate", "params": {"threshold_db": -20, "attack_ms": 500, "release_ms": 1000}}], 1),
    "spectral_default": ([{"type": "spectral", "params": {}}], 2),
}


def block_cases(n=16384, fs=48000):
    out, meta = {}, {}
    for k, (name, (config, ci)) in enumerate(BLOCK_CASES.items()):
        x = synth.clip(100 + k, n, ci, fs)
        # shift the burst into the excerpt so clipping and gating are exercised
        x = np.ascontiguousarray(np.roll(x, 5000, axis=0))
        y = ref_file_path(config, x, fs)
        out[name + "_y"] = y
        meta[name] = dict(config=config, ci=ci, clip=100 + k, n=n, fs=fs, roll=5000)
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), np.uint8)
    np.savez_compressed(os.path.join(HERE, "blocks.npz"), **out)
    print("blocks.npz:", len(meta), "cases")


def preset_cases(n_rain=32768, n_syn=24000):
    from scipy.io import wavfile
    fs, pcm = wavfile.read(os.path.join(REF, "music", "rain-raw.wav"))
    audio = (pcm.astype(np.float32) / np.float32(32768.0))          # libsndfile int16 -> f32 scale
    # loudest region of the file, so that Rain Delay's dry+wet sum reaches the clipper
    env = np.abs(audio).max(axis=1)
    centre = int(np.argmax(np.convolve(env, np.ones(4096), "same")))
    lo = max(0, min(len(audio) - n_rain, centre - n_rain // 2))
    mono = audio[lo:lo + n_rain].mean(axis=1, keepdims=True)         # engine.py:81-84
    out = {"rain_mono": mono}
    meta = {"rain": dict(fs=int(fs), start=lo, n=n_rain), "presets": {}}
    syn = synth.clip(7, n_syn, 2, 48000)
    syn = np.ascontiguousarray(np.roll(syn, 3000, axis=0))
    meta["syn"] = dict(clip=7, n=n_syn, fs=48000, roll=3000)
    for name, config in synth.PRESETS.items():
        key = name.replace(" ", "_")
        y = ref_file_path(config, mono, int(fs))
        out[f"rain_{key}"] = y
        out[f"rain_{key}_i16"] = (np.clip(y, -1.0, 1.0) * 32767).astype(np.int16)   # engine.py:104-105
        out[f"syn_{key}"] = ref_file_path(config, syn, 48000)
        meta["presets"][name] = key
    # the exact-N==1024 edge: no re-prepare, warm-up state of every block carries over
    x1024 = synth.clip(9, 1024, 1, 48000)
    out["n1024_Robot_Voice"] = ref_file_path(synth.PRESETS["Robot Voice"], x1024, 48000)
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), np.uint8)
    np.savez_compressed(os.path.join(HERE, "presets.npz"), **out)
    print("presets.npz:", len(meta["presets"]), "presets; rain excerpt at", lo)


if __name__ == "__main__":
    kernel_cases()
    block_cases()
    preset_cases()
