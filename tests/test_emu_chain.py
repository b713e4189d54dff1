"""Kernel-logic tests without a GPU: the product's chain kernel SOURCE
(csrc/aes_chain_kernel.cuh) and plan compiler run on the CPU emulator in
tests/cpu_emu and are held to the reference's golden vectors and to the oracle.
The same cases run on the real sm_100a build in tests/test_gpu_parity.py."""
import numpy as np
import pytest

import emu
import goldens
import synth
from oracle import oracle as orc

FP_TOL = 1e-5      # north_star: max-abs 1e-5 of full scale, SNR > 100 dB
NATIVE = [n for n in synth.PRESETS if n != "Clean Noise Removal"]      # spectral: no CUDA block yet


def check(got, want, exact=False, what=""):
    mx, snr = synth.err_stats(got, want)
    if exact:
        assert np.array_equal(got, want), (what, mx)
    scale = max(1.0, float(np.max(np.abs(want))))
    assert mx <= FP_TOL * scale and snr >= 100.0, (what, mx, snr)


def test_golden_blocks_through_emulated_kernel():
    z, meta = goldens.load("blocks")
    for name, m in meta.items():
        if name.startswith("spectral"):
            continue
        x = goldens.block_input(m)
        d = emu.resolved_descs(m["config"], m["fs"], m["n"], m["ci"])
        y = emu.run(d, m["fs"], x[None])[0]
        check(y, z[name + "_y"], exact=(name == "delay_fb0"), what=name)


def test_golden_presets_through_emulated_kernel():
    z, meta = goldens.load("presets")
    mono = np.ascontiguousarray(z["rain_mono"], np.float32)
    syn = goldens.syn_input(meta["syn"])
    for name in NATIVE:
        key = meta["presets"][name]
        cfg = synth.PRESETS[name]
        y = emu.run(emu.resolved_descs(cfg, meta["rain"]["fs"], mono.shape[0], 1), meta["rain"]["fs"], mono[None])[0]
        check(y, z[f"rain_{key}"], exact=(name == "Slapback Echo"), what=("rain", name))
        y = emu.run(emu.resolved_descs(cfg, 48000, syn.shape[0], 2), 48000, syn[None])[0]
        check(y, z[f"syn_{key}"], exact=(name == "Slapback Echo"), what=("syn", name))


@pytest.mark.parametrize("n", [1, 2, 255, 1024, 1025, 2999])
def test_ragged_lengths(n):
    cfg = synth.PRESETS["Robot Voice"] + synth.PRESETS["Guitar Filter"]
    x = synth.clip(11, n, 2)
    y = emu.run(emu.resolved_descs(cfg, 48000, n, 2), 48000, x[None])[0]
    check(y, orc.run_file_path(cfg, x, 48000), what=n)


def test_batch_striding_and_fresh_state_per_clip():
    cfg = synth.PRESETS["Rain Delay"]
    n = 21000
    x = synth.batch(20, 3, n)
    y = emu.run(emu.resolved_descs(cfg, 48000, n, 2), 48000, x, grid=2)      # CTA 0 takes clips 0 and 2
    for b in range(3):
        check(y[b], orc.run_file_path(cfg, x[b], 48000), what=b)


@pytest.mark.parametrize("fs,tile", [(44100, 1024), (22050, 512), (11025, 256)])
def test_smaller_tiles_at_lower_sample_rates(fs, tile):
    cfg = synth.PRESETS["Cathedral"]
    n = 9000
    x = synth.clip(5, n, 2, fs)
    d = emu.resolved_descs(cfg, fs, n, 2)
    y = emu.run(d, fs, x[None])[0]
    check(y, orc.run_file_path(cfg, x, fs), what=fs)


def test_int16_file_path_formats():
    """engine.py:78-84 down-mix on the way in, engine.py:104-105 quantise on the way out."""
    cfg = synth.PRESETS["Slapback Echo"]
    n = 12000
    rng = np.random.default_rng(5)
    pcm = rng.integers(-32768, 32767, (1, n, 2), dtype=np.int16)
    d = emu.resolved_descs(cfg, 48000, n, 1)
    q = emu.run(d, 48000, pcm, out_dtype=np.int16)[0]
    audio = pcm[0].astype(np.float32) / np.float32(32768.0)
    mono = orc.mono_downmix(audio)
    want = orc.quantize_i16(np.clip(orc.run_file_path(cfg, mono, 48000), -1.0, 1.0))
    assert np.array_equal(q, want)                       # index-only chain: bit-exact end to end


def test_distortion_and_peaking_extensions():
    cfg = [{"type": "distortion", "params": {"drive": 4.0, "mix": 0.7}},
           {"type": "filter", "params": {"filter_type": 3, "cutoff_hz": 1000, "q": 1.0, "gain_db": 6.0}}]
    n = 6000
    x = synth.clip(2, n, 2)
    y = emu.run(emu.resolved_descs(cfg, 48000, n, 2), 48000, x[None])[0]
    check(y, orc.run_file_path(cfg, x, 48000))


def test_unsupported_short_comb_is_refused():
    cfg = [{"type": "reverb", "params": {"comb_times_ms": (1.0, 2.0)}}]
    with pytest.raises(RuntimeError, match="shorter than the smallest tile"):
        emu.run(emu.resolved_descs(cfg, 48000, 4096, 2), 48000, synth.clip(0, 4096, 2)[None])


def test_final_state_of_filter_and_gate_matches_oracle():
    """What the reference keeps in the effect objects between calls (filter.py:35-40,
    gate.py:42) is read back from the device, also for a ragged last tile."""
    cfg = [{"type": "gate", "params": {"threshold_db": -25}},
           {"type": "filter", "params": {"filter_type": 0, "cutoff_hz": 300, "q": 2.0}}]
    for n in (1024, 3333):
        x = synth.clip(4, n, 2)
        st = np.zeros((1, 32), np.float64)
        emu.run(emu.resolved_descs(cfg, 48000, n, 2), 48000, x[None], state_out=st)
        ch = orc.build_chain(cfg, 48000, ci=2)
        ch.warmup()
        ch.process(x, np.zeros((n, 2), np.float32))
        gate, filt = ch.fx
        assert abs(st[0, 0] - gate.gain) < 1e-12
        got = st[0, 16:24].reshape(2, 4)
        assert np.max(np.abs(got - filt.state.astype(np.float64))) < 1e-6 * max(1.0, np.abs(filt.state).max())


RV_PRESETS = ("Rain Delay", "Cathedral")     # chains ending in the default reverb that take aes_rv_kernel by default
RV_FORCED = ("Guitar Filter",)               # ... and with AES_RV_BIQUAD=1 (slower on the GPU: registers)


@pytest.mark.parametrize("name", NATIVE)
def test_specialised_and_generic_kernels_agree_with_oracle(name, monkeypatch):
    """Every preset has a shape-specialised kernel (aes_fast_kernel.cuh); chains that end in the default
    reverb run on the software-pipelined kernel (aes_rv_kernel.cuh) first.  All of them and the generic
    interpreter (aes_chain_kernel.cuh) must give the oracle's answer."""
    cfg = synth.PRESETS[name]
    n = 26000
    x = synth.batch(30, 2, n)
    want = [orc.run_file_path(cfg, x[b], 48000) for b in range(2)]
    d = emu.resolved_descs(cfg, 48000, n, 2)
    runs = {}
    runs["default"] = emu.run(d, 48000, x)
    assert emu.lib().emu_last_was_fast() == (2 if name in RV_PRESETS else 1), name
    monkeypatch.setenv("AES_NO_RV", "1")
    runs["fast"] = emu.run(d, 48000, x)
    assert emu.lib().emu_last_was_fast() == 1, name
    monkeypatch.setenv("AES_NO_FAST", "1")
    runs["generic"] = emu.run(d, 48000, x)
    assert emu.lib().emu_last_was_fast() == 0
    for which, y in runs.items():
        for b in range(2):
            check(y[b], want[b], exact=(name == "Slapback Echo"), what=(name, which, b))


@pytest.mark.parametrize("name", RV_PRESETS + RV_FORCED)
@pytest.mark.parametrize("fs", [48000, 44100])
@pytest.mark.parametrize("n", [1, 777, 1024, 2048, 3100, 4096, 5000, 19000])
def test_pipelined_reverb_kernel_tile_edges(name, fs, n, monkeypatch):
    """aes_rv_kernel.cuh works one tile behind itself (all-pass walks, mix and store of tile i-1 in the
    phases of tile i): clips of less than one tile, exact multiples and ragged tails, two clips per CTA."""
    monkeypatch.setenv("AES_RV_BIQUAD", "1")
    cfg = synth.PRESETS[name]
    x = synth.batch(77, 3, n, fs=fs)
    y = emu.run(emu.resolved_descs(cfg, fs, n, 2), fs, x, grid=2)
    assert emu.lib().emu_last_was_fast() == 2
    for b in range(3):
        check(y[b], orc.run_file_path(cfg, x[b], fs), what=(name, fs, n, b))


def test_pipelined_reverb_kernel_formats_and_final_biquad_state(monkeypatch):
    """Mono / int16 input and int16 output through the pipelined kernel, and the biquad's DF-I state at
    the end of the clip (what a later streamed block continues from)."""
    monkeypatch.setenv("AES_RV_BIQUAD", "1")
    cfg = synth.PRESETS["Guitar Filter"]
    n = 5000
    x = synth.clip(9, n, 1)
    d = emu.resolved_descs(cfg, 48000, n, 1)
    st = np.zeros((1, 32))
    y = emu.run(d, 48000, x[None], state_out=st)[0]
    assert emu.lib().emu_last_was_fast() == 2
    want = orc.run_file_path(cfg, x, 48000)
    check(y, want, what="mono in")
    ch = orc.build_chain(cfg, 48000, ci=1)
    ch.warmup()
    ch.process(x, np.zeros((n, 2), np.float32))
    filt = ch.fx[0]
    got = st[0, 0:8].reshape(2, 4)
    assert np.max(np.abs(got - filt.state.astype(np.float64))) < 1e-6 * max(1.0, np.abs(filt.state).max())
    q = emu.run(d, 48000, x[None], out_dtype=np.int16)[0]
    wq = (np.clip(want, -1.0, 1.0) * np.float32(32767.0)).astype(np.int16)
    assert np.max(np.abs(q.astype(np.int32) - wq.astype(np.int32))) <= 1


@pytest.mark.parametrize("fs,topo", [(48000, 1), (44100, 2), (40000, 0)])
def test_reverb_topology_variants(fs, topo):
    """Reverb shapes exist with the 48 kHz / 44.1 kHz default delay lengths baked in
    (aes_fast_kernel.cuh, TOPO) and with run-time lengths; all three must match the oracle."""
    cfg = synth.PRESETS["Rain Delay"]
    n = 5000
    x = synth.clip(41, n, 2, fs)
    y = emu.run(emu.resolved_descs(cfg, fs, n, 2), fs, x[None])[0]
    assert emu.lib().emu_last_was_fast() == (2 if topo else 1)       # compile-time topology: the pipelined kernel
    assert emu.lib().emu_last_topo() == topo
    check(y, orc.run_file_path(cfg, x, fs), what=(fs, topo))


@pytest.mark.parametrize("variant", sorted(synth.REVERB_VARIANTS))
def test_reverb_parameter_corners(variant):
    cfg = [{"type": "reverb", "params": dict(synth.REVERB_VARIANTS[variant])}]
    n = 7000
    x = synth.clip(61, n, 2, 48000)
    y = emu.run(emu.resolved_descs(cfg, 48000, n, 2), 48000, x[None])[0]
    check(y, orc.run_file_path(cfg, x, 48000), what=variant)


@pytest.mark.parametrize("variant", sorted(synth.DELAY_VARIANTS))
def test_delay_parameter_corners(variant):
    cfg = [{"type": "delay", "params": dict(synth.DELAY_VARIANTS[variant])}]
    n = 40000 if variant != "longest" else 80000
    x = synth.clip(62, n, 2, 48000)
    y = emu.run(emu.resolved_descs(cfg, 48000, n, 2), 48000, x[None])[0]
    check(y, orc.run_file_path(cfg, x, 48000), what=variant)


@pytest.mark.parametrize("variant", sorted(synth.OCTAVER_VARIANTS))
def test_octaver_parameter_corners(variant):
    cfg = [{"type": "octaver", "params": dict(synth.OCTAVER_VARIANTS[variant])}]
    n = 12000
    x = synth.clip(63, n, 2, 48000)
    y = emu.run(emu.resolved_descs(cfg, 48000, n, 2), 48000, x[None])[0]
    assert emu.lib().emu_last_was_fast() == 1
    check(y, orc.run_file_path(cfg, x, 48000), what=variant)


@pytest.mark.parametrize("fs", [48000, 44100, 22050, 11025])
@pytest.mark.parametrize("pre_ms", [0.05, 3.0, 21.3, 60.0])
def test_predelay_lengths_both_kernels(pre_ms, fs, monkeypatch):
    """Pre-delays below the tile use a register line written ahead of its reads (barrier in
    between), longer ones the plain register line; 1-sample to 60 ms, aligned and not, on the
    specialised kernel where one exists and on the generic interpreter."""
    cfg = [{"type": "reverb", "params": {"pre_delay_ms": pre_ms}}]
    n = 6000
    x = synth.clip(5, n, 2, fs)
    want = orc.run_file_path(cfg, x, fs)
    d = emu.resolved_descs(cfg, fs, n, 2)
    check(emu.run(d, fs, x[None])[0], want, what=(pre_ms, fs, "default"))
    monkeypatch.setenv("AES_NO_FAST", "1")
    check(emu.run(d, fs, x[None])[0], want, what=(pre_ms, fs, "generic"))


def test_several_reverbs_in_one_chain_and_the_shared_memory_limit():
    """Ring state of every reverb lives in the clip's CTA: three default reverbs still fit one CTA per
    SM (generic kernel), a fourth is refused with the byte count instead of running wrong."""
    R = {"type": "reverb", "params": {}}
    n = 4000
    x = synth.clip(3, n, 2, 48000)
    for k in (2, 3):
        cfg = [R] * k
        y = emu.run(emu.resolved_descs(cfg, 48000, n, 2), 48000, x[None])[0]
        check(y, orc.run_file_path(cfg, x, 48000), what=k)
    with pytest.raises(RuntimeError, match="shared memory"):
        emu.run(emu.resolved_descs([R] * 4, 48000, n, 2), 48000, x[None])
