"""CPU-side tests of the product's host logic and of the C-ABI boundary (no compute)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import audioblocks as ab
from audioblocks import _native
from audioblocks.engine import make_effect
from oracle import oracle as orc
import emu
import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_package_exports_reference_names():
    # src/audioblocks/__init__.py:1-8 of the reference
    for name in ("SmoothParam", "EffectsChain", "pick_devices", "Effect", "PlotDataTap", "StereoDelayEffect",
                 "ReverbEffect", "AudioEngine", "SAMPLE_RATE", "NoiseGateEffect", "SpectralFilter",
                 "OctaverEffect", "FilterEffect"):
        assert hasattr(ab, name), name
    assert ab.SAMPLE_RATE == 48000


def test_smoothparam_semantics():
    p = ab.SmoothParam(5.0, 0.0, 1.0)          # ctor does not clamp (core.py:57-59)
    assert p.current == 5.0 and p.target == 5.0
    p.set_target(7.0)
    assert p.target == 1.0
    assert p.step_towards(0.5) == 4.5
    p.nudge(-3.0)
    assert p.target == 0.0
    with pytest.raises(ValueError):
        p.step_towards(-1.0)
    q = ab.SmoothParam(0.25)
    assert q.step_towards(0.1) == 0.25


def test_setters_and_smoothparam_attributes_exist():
    # engine.py:131-145 update_param relies on these names
    table = {
        ab.StereoDelayEffect: (["delay_ms", "feedback"], ["set_delay_ms", "nudge_delay_ms", "set_feedback",
                                                          "set_mix_dry", "set_mix_wet", "set_offset_ms"]),
        ab.ReverbEffect: (["rt60_s", "damp", "pre_delay_ms"], ["set_rt60_s", "set_damp", "set_pre_delay_ms",
                                                                "set_mix", "set_mix_wet", "set_mix_dry"]),
        ab.FilterEffect: (["filter_type", "cutoff_hz", "q"], ["set_filter_type", "set_cutoff_hz", "set_q"]),
        ab.OctaverEffect: (["semitones", "mix"], ["set_semitones", "set_mix"]),
        ab.NoiseGateEffect: (["threshold_db", "attack_ms", "release_ms"],
                             ["set_threshold_db", "set_attack_ms", "set_release_ms"]),
        ab.SpectralFilter: (["threshold_db", "reduction"], ["set_threshold_db", "set_reduction"]),
    }
    for cls, (params, setters) in table.items():
        fx = cls()
        for p in params:
            assert isinstance(getattr(fx, p), ab.SmoothParam), (cls, p)
        for s in setters:
            assert callable(getattr(fx, s)), (cls, s)
    fx = ab.FilterEffect()
    fx.set_filter_type(3)
    assert fx.filter_type.target == 2.0        # clamp range of the reference (filter.py:45)


@pytest.mark.parametrize("fs", [48000, 44100, 22050, 96000])
def test_integer_lags_match_oracle_expressions(fs):
    for params in ({}, {"delay_ms": 120, "offset_ms": 10}, {"delay_ms": 33.3, "offset_ms": 7.7},
                   {"delay_ms": 5000.0, "max_delay_ms": 200.0}, {"delay_ms": 0.0}):
        fx, o = ab.StereoDelayEffect(**params), orc.ODelay(**params)
        fx.prepare(fs, 2, 2, 1024); o.prepare(fs, 2, 2, 1024)
        size, dl, dr = o.lags()
        want = [d if d > 0 else size for d in (dl, dr)]
        d = fx._stages(1024)[0]
        assert [d.q[0], d.q[1]] == want
    for params in ({}, {"jitter_ms": 1.3, "pre_delay_ms": 20}, {"comb_times_ms": (10.0, 300.0), "allpass_times_ms": (0.01,)}):
        fx, o = ab.ReverbEffect(**params), orc.OReverb(**params)
        fx.prepare(fs, 2, 2, 1024); o.prepare(fs, 2, 2, 1024)
        d = fx._stages(1024)[0]
        assert d.q[2] == o.pre_dS()
        for s, side in enumerate(o.sides):
            assert [d.q[4 + 8 * s + c] for c in range(d.q[0])] == [c["L"] for c in side["comb"]]
            assert [d.q[20 + 4 * s + k] for k in range(d.q[1])] == [a["L"] for a in side["ap"]]
            assert [d.p[4 + 8 * s + c] for c in range(d.q[0])] == [o.gain(c["L"]) for c in side["comb"]]


def test_reference_line_lengths_at_48k():
    # SURVEY 3.1 [probe]: combs L [1440,1795,1987,2112] R [1411,1766,1958,2083]; AP L [242,84] R [237,78]
    fx = ab.ReverbEffect()
    fx.prepare(48000, 1, 2, 1024)
    d = fx._stages(1024)[0]
    assert [d.q[4 + c] for c in range(4)] == [1440, 1795, 1987, 2112]
    assert [d.q[12 + c] for c in range(4)] == [1411, 1766, 1958, 2083]
    assert [d.q[20], d.q[21], d.q[24], d.q[25]] == [242, 84, 237, 78]


def test_octaver_warmup_phase_leaks_into_the_file():
    # SURVEY 3.1 [probe]: fs=48k, 40 ms, -12 st -> size 1920, w=128, phasor=0.5333333333333231
    descs = emu.resolved_descs([{"type": "octaver", "params": {"semitones": -12, "mix": 1.0}}], 48000, 50000, 1)
    d = descs[0]
    assert d.q[0] == 1920 and d.q[1] == 128
    assert d.p[0] == 0.5333333333333231


def test_warmup_of_a_chain_at_rest_launches_nothing(monkeypatch):
    """engine.py:96-99 pushes two zero blocks through the fresh chain.  Zero in, zero out: the state they
    leave (write pointers, phasor, gate gain, smoothed parameters, spectral mask) is advanced on the host."""
    def boom(*a, **k):
        raise AssertionError("the warm-up of a chain at rest must not reach the library")
    monkeypatch.setattr(_native, "stream_process", boom)
    monkeypatch.setattr(_native, "ChainPlan", boom)
    monkeypatch.setattr(_native, "SpectralPlan", boom)
    cfg = [{"type": "filter", "params": {"filter_type": 1, "cutoff_hz": 80}},
           {"type": "octaver", "params": {"semitones": -12, "mix": 1.0}},
           {"type": "gate", "params": {"threshold_db": -30}},
           {"type": "delay", "params": {"delay_ms": 100}},
           {"type": "spectral", "params": {"reduction": 0.25, "smoothing": 0.8}},
           {"type": "reverb", "params": {}}]
    chain = ab.EffectsChain(48000, 1, 2, 1024)
    for c in cfg:
        fx = make_effect(c)
        assert fx is not None, c
        chain.add(fx)
    gate = chain.effects[2]
    gate._gain_state = 0.5
    chain.warmup()
    octv, spec = chain.effects[1], chain.effects[4]
    assert (octv.size, octv.w, octv.phasor) == (1920, 128, 0.5333333333333231)        # SURVEY 3.1 [probe]
    k = 1.0 - gate._calc_coeff(100.0)
    g = 0.5
    for _ in range(2048):
        g = k * g
    assert gate._gain_state == g
    assert all(fx._n_total == 2048 for fx in chain.effects if hasattr(fx, "_n_total"))
    m = np.float32(1.0)
    for _ in range(2):
        m = np.float32(0.8) * m + (np.float32(1.0) - np.float32(0.8)) * np.float32(0.25)
    assert spec.mask_smooth.shape == (1025,) and np.all(spec.mask_smooth == m)
    out = np.ones((1024, 2), np.float32)
    chain.process(np.zeros((1024, 1), np.float32), out)
    assert not out.any()


def test_filter_and_gate_constants_match_oracle():
    for t, fc, q in [(0, 1000.0, 0.707), (1, 80.0, 0.707), (2, 800.0, 0.8), (0, 20.0, 10.0)]:
        fx, o = ab.FilterEffect(t, fc, q), orc.OFilter(t, fc, q)
        fx.prepare(44100, 2, 2, 64); o.prepare(44100, 2, 2, 64)
        d = fx._stages(64)[0]
        assert tuple(d.p[i] for i in range(5)) == o.coeffs()
    fx, o = ab.NoiseGateEffect(-30, 10, 100), orc.OGate(-30, 10, 100)
    fx.prepare(48000, 2, 2, 64); o.prepare(48000, 2, 2, 64)
    d = fx._stages(64)[0]
    assert (d.p[0], d.p[1], d.p[2]) == o.consts()


def test_make_effect_skips_unknown_types():
    assert make_effect({"type": "nope"}) is None
    assert isinstance(make_effect({"type": "delay", "params": {"delay_ms": 100}}), ab.StereoDelayEffect)


def test_c_abi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "aesim.h")).read()
    names = sorted(set(re.findall(r"\b(aes_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 20
    if not os.path.exists(_native.LIB_PATH):
        pytest.skip("libaesim.so not built (run __graft_entry__.build())")
    L = C.CDLL(_native.LIB_PATH)
    for n in names:
        assert hasattr(L, n), f"libaesim.so does not export {n}"
    assert L.aes_abi_version() == 1


def test_no_cpu_fallback_without_a_device():
    """On a GPU-less machine the product must fail loudly, not compute on the CPU."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a CUDA device is present")
    chain = ab.EffectsChain(48000, 1, 2, 1024)
    chain.add(ab.StereoDelayEffect())
    # (signal, not silence: a silent block through a chain at rest is zeros by construction and launches nothing)
    with pytest.raises(_native.AesimError):
        chain.process(np.full((1024, 1), 0.25, np.float32), np.zeros((1024, 2), np.float32))
    with pytest.raises(_native.AesimError):
        ab.SpectralFilter().process_into(np.full((8, 2), 0.25, np.float32), np.zeros((8, 2), np.float32))
    with pytest.raises(_native.AesimError):
        ab.ConvolutionReverbEffect(np.ones((4, 2), np.float32)).process_into(np.zeros((8, 2), np.float32),
                                                                               np.zeros((8, 2), np.float32))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "audio-effects-simulator_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no CPU fallback", ""), (f, "mentions the oracle")
                assert "cuda_emu" not in src or f.endswith("aes_chain_kernel.cuh"), f


def test_cpulist_parsing_and_harmless_bind():
    from audioblocks.sharding import _parse_cpulist, bind_host_to_gpu
    assert _parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert _parse_cpulist("") == set()
    import os
    before = os.sched_getaffinity(0)
    assert bind_host_to_gpu("0000:ff:1f.0") is None          # no such device: nothing changes
    assert os.sched_getaffinity(0) == before
