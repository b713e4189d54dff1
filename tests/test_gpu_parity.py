"""Parity tests proper: the sm_100a build behind the C ABI (through the drop-in
`audioblocks` API) against the reference's golden vectors and the CPU oracle."""
import ctypes as C

import numpy as np
import pytest

import goldens
import synth

pytestmark = pytest.mark.gpu

FP_TOL = 1e-5      # north_star: max-abs 1e-5 of full scale, SNR > 100 dB
NATIVE = [n for n in synth.PRESETS if n != "Clean Noise Removal"]


@pytest.fixture(scope="module")
def ab():
    import audioblocks
    from audioblocks import _native
    _native.lib()                                  # fail loudly if the CUDA library is missing
    return audioblocks


@pytest.fixture(scope="module")
def orc():
    from oracle import oracle
    return oracle


def check(got, want, exact=False, what=""):
    mx, snr = synth.err_stats(got, want)
    if exact:
        assert np.array_equal(got, want), (what, mx)
    scale = max(1.0, float(np.max(np.abs(want))))
    assert mx <= FP_TOL * scale and snr >= 100.0, (what, mx, snr)


def run_file(ab, config, x, fs):
    """engine.py:86-102 through the drop-in package: build@1024, warm-up, one whole call."""
    from audioblocks.engine import file_chain
    chain = file_chain(config, fs, channels_in=x.shape[1])
    out = np.zeros((x.shape[0], 2), np.float32)
    chain.process(x, out)
    return out


def test_golden_blocks(ab):
    z, meta = goldens.load("blocks")
    for name, m in meta.items():
        if name.startswith("spectral"):
            continue
        y = run_file(ab, m["config"], goldens.block_input(m), m["fs"])
        check(y, z[name + "_y"], exact=(name == "delay_fb0"), what=name)


def test_golden_presets_rain_and_synthetic(ab):
    z, meta = goldens.load("presets")
    mono = np.ascontiguousarray(z["rain_mono"], np.float32)
    syn = goldens.syn_input(meta["syn"])
    for name in NATIVE:
        key = meta["presets"][name]
        y = run_file(ab, synth.PRESETS[name], mono, meta["rain"]["fs"])
        check(y, z[f"rain_{key}"], exact=(name == "Slapback Echo"), what=("rain", name))
        q = (np.clip(y, -1.0, 1.0) * 32767).astype(np.int16)
        assert np.max(np.abs(q.astype(np.int32) - z[f"rain_{key}_i16"].astype(np.int32))) <= 1
        y = run_file(ab, synth.PRESETS[name], syn, 48000)
        check(y, z[f"syn_{key}"], exact=(name == "Slapback Echo"), what=("syn", name))


@pytest.mark.parametrize("n", [1, 2, 255, 1024, 1025, 2999, 70001])
def test_ragged_lengths(ab, orc, n):
    cfg = synth.PRESETS["Robot Voice"] + synth.PRESETS["Guitar Filter"]
    x = synth.clip(11, n, 2)
    check(run_file(ab, cfg, x, 48000), orc.run_file_path(cfg, x, 48000), what=n)


@pytest.mark.parametrize("name", NATIVE)
def test_batch_matches_oracle_per_clip(ab, orc, name):
    from audioblocks.engine import file_chain
    cfg = synth.PRESETS[name]
    n, B = 100000, 6
    x = synth.batch(40, B, n)
    y = file_chain(cfg, 48000, channels_in=2).process_batch(x)
    for b in range(B):
        check(y[b], orc.run_file_path(cfg, x[b], 48000), exact=(name == "Slapback Echo"), what=(name, b))


def test_many_clips_fresh_state_and_grid_striding(ab, orc):
    """More clips than resident CTAs: every clip must start from fresh lines."""
    from audioblocks.engine import file_chain
    cfg = synth.PRESETS["Rain Delay"]
    n, B = 20000, 700
    base = synth.batch(60, 4, n)
    x = np.ascontiguousarray(np.tile(base, (B // 4, 1, 1)))
    y = file_chain(cfg, 48000, channels_in=2).process_batch(x)
    want = [orc.run_file_path(cfg, base[k], 48000) for k in range(4)]
    for b in range(B):
        assert np.array_equal(y[b], y[b % 4]), b              # identical inputs -> identical outputs
    for k in range(4):
        check(y[k], want[k], what=k)


@pytest.mark.parametrize("fs", [44100, 22050, 11025])
def test_lower_sample_rates_use_smaller_tiles(ab, orc, fs):
    cfg = synth.PRESETS["Cathedral"]
    x = synth.clip(5, 30000, 2, fs)
    check(run_file(ab, cfg, x, fs), orc.run_file_path(cfg, x, fs), what=fs)


def test_int16_file_path_bit_exact(ab, orc):
    """engine.py:78-84 down-mix in, engine.py:104-105 quantise out; index-only chain."""
    from audioblocks.engine import file_chain
    cfg = synth.PRESETS["Slapback Echo"]
    n = 50000
    rng = np.random.default_rng(5)
    pcm = rng.integers(-32768, 32767, (3, n, 2), dtype=np.int16)
    out = np.empty((3, n, 2), np.int16)
    file_chain(cfg, 48000, channels_in=1).process_batch(pcm, out)
    for b in range(3):
        mono = orc.mono_downmix(pcm[b].astype(np.float32) / np.float32(32768.0))
        want = orc.quantize_i16(np.clip(orc.run_file_path(cfg, mono, 48000), -1.0, 1.0))
        assert np.array_equal(out[b], want)


def test_extensions_distortion_peaking(ab, orc):
    cfg = [{"type": "distortion", "params": {"drive": 4.0, "mix": 0.7}},
           {"type": "filter", "params": {"filter_type": 3, "cutoff_hz": 1000, "q": 1.0, "gain_db": 6.0}},
           {"type": "octaver", "params": {"semitones": -12, "mix": 0.5}},
           {"type": "delay", "params": {"delay_ms": 120, "feedback": 0.3, "offset_ms": 10}}]
    x = synth.clip(2, 60000, 2)
    check(run_file(ab, cfg, x, 48000), orc.run_file_path(cfg, x, 48000))


def test_c2_biquad_cascade_long_clip(ab, orc):
    """BASELINE configs[1]: one long stereo clip through LP/HP/BP/peaking, plus the low-fc stress."""
    n = 48000 * 20
    x = synth.clip(1, n, 2)
    cfg = [{"type": "filter", "params": {"filter_type": 0, "cutoff_hz": 8000, "q": 0.707}},
           {"type": "filter", "params": {"filter_type": 1, "cutoff_hz": 80, "q": 0.707}},
           {"type": "filter", "params": {"filter_type": 2, "cutoff_hz": 1000, "q": 0.8}},
           {"type": "filter", "params": {"filter_type": 3, "cutoff_hz": 1000, "q": 1.0, "gain_db": 6.0}}]
    check(run_file(ab, cfg, x, 48000), orc.run_file_path(cfg, x, 48000), what="cascade")
    cfg = [{"type": "filter", "params": {"filter_type": 0, "cutoff_hz": 40, "q": 5.0}}]
    check(run_file(ab, cfg, x, 48000), orc.run_file_path(cfg, x, 48000), what="lp40q5")


def test_full_size_properties(ab):
    """BASELINE-size clip (10 s): size-independent properties instead of a CPU run.
    Slapback (fb=0): out[n] = clip(x[n] + 0.5*x[n-4800]) exactly; and linearity of the
    un-clipped reverb: reverb(a*x) == a*reverb(x) for a power-of-two gain."""
    from audioblocks.engine import file_chain
    n = 480000
    x = synth.batch(80, 2, n) * np.float32(0.25)
    y = file_chain(synth.PRESETS["Slapback Echo"], 48000, channels_in=2).process_batch(x)
    d = np.zeros_like(x)
    d[:, 4800:] = x[:, :-4800]
    want = np.clip(np.float32(1.0) * x + np.float32(0.5) * d, -1.0, 1.0)
    assert np.array_equal(y, want)
    cfg = [{"type": "reverb", "params": {"rt60_s": 1.0, "mix_wet": 0.25, "mix_dry": 0.25}}]
    x = x * np.float32(0.25)
    y1 = file_chain(cfg, 48000, channels_in=2).process_batch(x)
    y2 = file_chain(cfg, 48000, channels_in=2).process_batch(x * np.float32(0.5))
    assert np.max(np.abs(y1)) < 1.0
    assert np.array_equal(y1 * np.float32(0.5), y2)          # scaling by 2^-1 commutes with every rounding


def test_engine_process_file_arrays(ab, orc):
    import queue
    eng = ab.AudioEngine({"input": queue.Queue(10), "output": queue.Queue(10)})
    eng.last_chain_config = synth.PRESETS["Rain Delay"]
    z, meta = goldens.load("presets")
    mono = np.ascontiguousarray(z["rain_mono"], np.float32)
    stereo = np.repeat(mono, 2, axis=1)
    m, processed, pcm = eng.process_file_arrays(stereo, meta["rain"]["fs"])
    assert np.array_equal(m, mono)
    check(processed, np.clip(z["rain_Rain_Delay"], -1, 1))
    assert np.max(np.abs(pcm.astype(np.int32) - z["rain_Rain_Delay_i16"].astype(np.int32))) <= 1


def test_single_block_c_abi_entries(ab, orc):
    import torch
    from audioblocks import _native
    L = _native.lib()
    n, B = 30000, 3
    x = synth.batch(90, B, n)
    xd = torch.from_numpy(x).cuda()
    yd = torch.empty_like(xd)
    _native.check(L.aes_delay_f32(xd.data_ptr(), yd.data_ptr(), B, n, 4800, 5280, 0.3, 1.0, 0.3, None))
    cfg = [{"type": "delay", "params": {"delay_ms": 100, "feedback": 0.3, "mix_wet": 0.3, "mix_dry": 1.0, "offset_ms": 10}}]
    for b in range(B):
        check(yd[b].cpu().numpy(), orc.run_file_path(cfg, x[b], 48000))
    qd = torch.empty((B, n, 2), dtype=torch.int16, device="cuda")
    _native.check(L.aes_quantize_i16(yd.data_ptr(), qd.data_ptr(), B * n * 2, None))
    torch.cuda.synchronize()
    assert np.array_equal(qd.cpu().numpy(), (np.clip(yd.cpu().numpy(), -1, 1) * 32767).astype(np.int16))
    assert L.aes_launch_count() > 0


@pytest.mark.parametrize("name", NATIVE)
def test_generic_interpreter_kernel_matches_too(ab, orc, name, monkeypatch):
    """The presets normally run on shape-specialised kernels; the generic interpreter
    (used for arbitrary chains) must agree."""
    from audioblocks.engine import file_chain
    monkeypatch.setenv("AES_NO_FAST", "1")
    cfg = synth.PRESETS[name]
    n, B = 60000, 3
    x = synth.batch(50, B, n)
    y = file_chain(cfg, 48000, channels_in=2).process_batch(x)
    for b in range(B):
        check(y[b], orc.run_file_path(cfg, x[b], 48000), exact=(name == "Slapback Echo"), what=(name, b))


UNLISTED = {
    "filter>delay>reverb": [
        {"type": "filter", "params": {"filter_type": 0, "cutoff_hz": 4000, "q": 0.707}},
        {"type": "delay", "params": {"delay_ms": 250, "feedback": 0.35, "mix_wet": 0.5, "mix_dry": 1.0, "offset_ms": 15}},
        {"type": "reverb", "params": {"rt60_s": 1.5, "mix_wet": 0.3, "mix_dry": 0.9}}],
    "octaver>reverb": [
        {"type": "octaver", "params": {"semitones": 7, "mix": 0.4}},
        {"type": "reverb", "params": {"rt60_s": 1.2, "mix_wet": 0.25, "mix_dry": 0.9}}],
    "delay>gate>delay": [
        {"type": "delay", "params": {"delay_ms": 80, "feedback": 0.5, "mix_wet": 0.5, "mix_dry": 1.0, "offset_ms": 0}},
        {"type": "gate", "params": {"threshold_db": -35, "attack_ms": 5, "release_ms": 80}},
        {"type": "delay", "params": {"delay_ms": 300, "feedback": 0.2, "mix_wet": 0.4, "mix_dry": 1.0, "offset_ms": 20}}],
}


@pytest.mark.parametrize("name", sorted(UNLISTED))
@pytest.mark.parametrize("in_fmt,out_fmt", [("f32_stereo", "f32"), ("f32_mono", "i16"), ("i16", "i16")])
def test_chains_without_their_own_kernel_run_as_runs_of_specialised_ones(ab, orc, name, in_fmt, out_fmt, monkeypatch):
    """A chain no kernel was instantiated for is cut at stage boundaries into runs that have one
    (aes_chain.cu, plan_create): every kernel hands float32 from stage to stage, so the cut itself changes
    nothing; the result agrees with the generic interpreter to the kernels' own rounding and sits within
    the parity bar of the oracle."""
    from audioblocks import _native
    from audioblocks.engine import file_chain
    cfg = UNLISTED[name]
    n, B = 70000, 5
    x = synth.batch(60, B, n, 1 if in_fmt == "f32_mono" else 2)
    if in_fmt == "i16":
        xin, fi = (np.clip(x, -1, 1) * 32767).astype(np.int16), _native.FMT_I16_DOWNMIX
    else:
        xin, fi = x, (_native.FMT_F32_MONO if in_fmt == "f32_mono" else _native.FMT_F32_STEREO)
    fo, odt = (_native.FMT_I16_STEREO, np.int16) if out_fmt == "i16" else (_native.FMT_F32_STEREO, np.float32)

    def run():
        ch = file_chain(cfg, 48000, channels_in=1 if in_fmt != "f32_stereo" else 2)
        plan = ch.prepare_batch(n)
        y = np.zeros((B, n, 2), odt)
        plan.run_host(xin, fi, y, fo, B, n)
        kern = plan.info()["kernel"]
        plan.close()
        return y, kern
    y_split, k_split = run()
    monkeypatch.setenv("AES_NO_SPLIT", "1")
    y_whole, k_whole = run()
    assert k_split.startswith("split:") and "generic" in k_whole, (k_split, k_whole)
    # (the specialised reverb kernels sum their comb recurrences as scans: ulp-level differences to the interpreter)
    d = np.max(np.abs(y_split.astype(np.float64) - y_whole.astype(np.float64)))
    assert d <= (1 if out_fmt == "i16" else 2e-6), d
    if in_fmt == "f32_stereo" and "gate" not in name:      # (a gate behind a float block: see test_fuzz_chains.check)
        for b in (0, B - 1):
            check(y_split[b], orc.run_file_path(cfg, x[b], 48000), what=(name, b))


def test_time_parallel_scan_equals_batch_kernel_on_a_single_clip(ab, orc, monkeypatch):
    """One long clip through biquads takes the decoupled-look-back scan kernel
    (aes_biquad_scan.cuh); forcing the one-CTA-per-clip kernel must give the same audio,
    and the carried DF-I state must match the oracle's (filter.py:35-40)."""
    n = 48000 * 8 + 123
    x = synth.clip(6, n, 2)
    cfg = [{"type": "filter", "params": {"filter_type": 0, "cutoff_hz": 300, "q": 2.0}},
           {"type": "filter", "params": {"filter_type": 1, "cutoff_hz": 50, "q": 0.9}}]
    from audioblocks.engine import file_chain
    ch = file_chain(cfg, 48000, channels_in=2)
    y_scan = np.zeros((n, 2), np.float32)
    ch.process(x, y_scan)
    states = [fx._state.copy() for fx in ch.effects]
    monkeypatch.setenv("AES_NO_SCAN", "1")
    y_batch = run_file(ab, cfg, x, 48000)
    want = orc.run_file_path(cfg, x, 48000)
    check(y_scan, want, what="scan")
    check(y_batch, want, what="batch")
    och = orc.build_chain(cfg, 48000, ci=2)
    och.warmup()
    och.process(x, np.zeros((n, 2), np.float32))
    for st, ofx in zip(states, och.fx):
        assert np.max(np.abs(st - ofx.state)) <= 1e-5 * max(1.0, float(np.abs(ofx.state).max()))


def test_truncated_lookback_matches_chained_lookback_and_is_reproducible(ab, orc, monkeypatch):
    """aes_biquad_scan.cuh: filters that forget within a 256-tile window sum a fixed number of tile
    aggregates instead of chaining look-backs.  Same audio as the chained look-back (forced by
    AES_SCAN_CHAINED) to an f32 ulp, bit-identical from run to run, and within the bar of the oracle
    for BASELINE configs[1]'s cascade and the low-cutoff stress pair."""
    n = 48000 * 20 + 77
    x = synth.clip(9, n, 2)
    cascade = [{"type": "filter", "params": {"filter_type": 0, "cutoff_hz": 8000, "q": 0.707}},
               {"type": "filter", "params": {"filter_type": 1, "cutoff_hz": 80, "q": 0.707}},
               {"type": "filter", "params": {"filter_type": 2, "cutoff_hz": 1000, "q": 0.8}},
               {"type": "filter", "params": {"filter_type": 3, "cutoff_hz": 1000, "q": 1.0, "gain_db": 6.0}}]
    stress = [{"type": "filter", "params": {"filter_type": 0, "cutoff_hz": 40, "q": 5.0}},
              {"type": "filter", "params": {"filter_type": 1, "cutoff_hz": 20, "q": 0.707}}]
    for cfg in (cascade, stress):
        y1 = run_file(ab, cfg, x, 48000)
        y2 = run_file(ab, cfg, x, 48000)
        assert np.array_equal(y1, y2)
        check(y1, orc.run_file_path(cfg, x, 48000), what="truncated look-back")
        monkeypatch.setenv("AES_SCAN_CHAINED", "1")
        y3 = run_file(ab, cfg, x, 48000)
        monkeypatch.delenv("AES_SCAN_CHAINED")
        assert np.max(np.abs(y1 - y3)) <= 2.4e-7 * max(1.0, float(np.abs(y3).max()))


def test_scan_sub_batches_on_different_streams_share_one_scratch(ab, orc):
    """A host call with a few long clips is cut into sub-batches that run on the pipeline's four
    streams; every one of them takes the time-parallel scan, whose look-back records, ticket and
    epoch live in ONE per-plan scratch (launches are chained by an event; the last, smaller
    sub-batch re-lays the scratch out).  Twice through the same plan, against the oracle."""
    n, B = 2_600_000, 7
    x = synth.batch(70, B, n)
    cfg = [{"type": "filter", "params": {"filter_type": 1, "cutoff_hz": 120, "q": 0.9}},
           {"type": "filter", "params": {"filter_type": 0, "cutoff_hz": 5000, "q": 0.707}}]
    from audioblocks.engine import file_chain
    chain = file_chain(cfg, 48000, channels_in=2)
    y1 = chain.process_batch(x)
    y2 = chain.process_batch(x)
    assert np.array_equal(y1, y2)
    for b in (0, 3, 6):
        check(y1[b], orc.run_file_path(cfg, x[b], 48000), what=("scan sub-batch", b))


def test_convolution_reverb_3s_ir(ab, orc):
    """BASELINE configs[3] shape at test size: 3 s synthetic IR (144 000 taps, 18 partitions of
    the 16384-point FFT), 5 s clips; float64 fftconvolve oracle (parity unpinned by the reference)."""
    ir = orc.synthetic_ir(144000)
    n, B = 48000 * 5 + 777, 3
    x = synth.batch(70, B, n)
    fx = ab.ConvolutionReverbEffect(ir, mix_dry=0.7, mix_wet=0.5)
    assert fx.plan().info() == {"fft_size": 16384, "partitions": 18}
    y = fx.process_batch(x)
    for b in range(B):
        want = np.zeros_like(x[b])
        orc.OConvReverb(ir, 0.7, 0.5).process_into(x[b], want)
        check(y[b], want, what=b)
    # inside a chain, between fused segments, with a mono file-path input
    from audioblocks.engine import file_chain
    cfg = [{"type": "filter", "params": {"filter_type": 1, "cutoff_hz": 120}},
           {"type": "convreverb", "params": {"ir": ir[:20000], "mix_wet": 0.8}},
           {"type": "delay", "params": {"delay_ms": 100, "feedback": 0.0, "mix_wet": 0.5, "mix_dry": 1.0, "offset_ms": 0}}]
    xm = synth.clip(3, 50000, 1)
    got = np.zeros((50000, 2), np.float32)
    file_chain(cfg, 48000, channels_in=1).process(xm, got)
    check(got, orc.run_file_path(cfg, xm, 48000), what="chain")
    yb = file_chain(cfg, 48000, channels_in=1).process_batch(xm[None])
    assert np.array_equal(yb[0], got)


def test_convolution_reverb_small_fft_sizes(ab, orc):
    for log2n, taps in ((8, 300), (11, 5000)):
        ir = orc.synthetic_ir(taps, rt60=0.05)
        x = synth.batch(75, 2, 30000)
        y = ab.ConvolutionReverbEffect(ir, 0.5, 0.9, block_log2=log2n).process_batch(x)
        for b in range(2):
            want = np.zeros_like(x[b])
            orc.OConvReverb(ir, 0.5, 0.9).process_into(x[b], want)
            check(y[b], want, what=(log2n, b))


def test_spectral_filter_streaming_blocks_match_reference_semantics(ab, orc):
    """spectral.py:44-100 with its live block size: 256-frame blocks, 512-point frames, carried
    analysis buffer / overlap-add accumulator / smoothed mask."""
    fx, ofx = ab.SpectralFilter(-30.0, 0.2, 0.8), orc.OSpectral(-30.0, 0.2, 0.8)
    fx.prepare(48000, 2, 2, 256)
    ofx.prepare(48000, 2, 2, 256)
    x = synth.clip(12, 256 * 12, 2)
    x[256 * 6:] *= np.float32(0.002)                     # second half under the threshold: the mask moves
    for k in range(12):
        blk = np.ascontiguousarray(x[256 * k:256 * (k + 1)])
        got, want = np.zeros((256, 2), np.float32), np.zeros((256, 2), np.float32)
        fx.process_into(blk, got)
        ofx.process_into(blk, want)
        assert np.max(np.abs(got - want)) <= 1e-5, k
    assert np.max(np.abs(fx.mask_smooth - ofx.mask)) <= 1e-5


def test_clean_noise_removal_preset_and_spectral_golden(ab, orc):
    """Whole-file mode: hop == N, one 2N-point frame, output = its zero-padded half:
    the reference's own result is ~1e-8 noise, so only the max-abs bar applies (no SNR)."""
    z, meta = goldens.load("blocks")
    m = meta["spectral_default"]
    y = run_file(ab, m["config"], goldens.block_input(m), m["fs"])
    assert np.max(np.abs(y - z["spectral_default_y"])) <= 1e-5
    zp, mp = goldens.load("presets")
    mono = np.ascontiguousarray(zp["rain_mono"], np.float32)
    cfg = synth.PRESETS["Clean Noise Removal"]
    y = run_file(ab, cfg, mono, mp["rain"]["fs"])
    assert np.max(np.abs(y - zp["rain_Clean_Noise_Removal"])) <= 1e-5
    syn = goldens.syn_input(mp["syn"])
    y = run_file(ab, cfg, syn, 48000)
    assert np.max(np.abs(y - zp["syn_Clean_Noise_Removal"])) <= 1e-5
    # batch entry, awkward frame count (2N = 2 * 30011 is not a power of two), quiet clip so bins gate
    from audioblocks.engine import file_chain
    xb = synth.batch(21, 3, 30011) * np.float32(1e-4)
    yb = file_chain(cfg, 48000, channels_in=2).process_batch(xb)
    for b in range(3):
        assert np.max(np.abs(yb[b] - orc.run_file_path(cfg, xb[b], 48000))) <= 1e-5


def test_engine_process_wav_file_websocket_reply(ab, orc):
    """engine.py:67-129 end to end: data-URL WAV in -> JSON reply with the processed WAV, the
    reply shape the browser expects (assets/02_custom.js:377-390)."""
    import asyncio
    import base64
    import io
    import json
    import queue
    import scipy.io.wavfile

    fs, n = 44100, 30000
    rng = np.random.default_rng(11)
    pcm = (rng.uniform(-0.6, 0.6, (n, 2)) * 32767).astype(np.int16)
    buf = io.BytesIO()
    scipy.io.wavfile.write(buf, fs, pcm)
    url = "data:audio/wav;base64," + base64.b64encode(buf.getvalue()).decode("ascii")

    class FakeSocket:
        def __init__(self): self.sent = []
        async def send(self, msg): self.sent.append(msg)

    eng = ab.AudioEngine({"input": queue.Queue(10), "output": queue.Queue(10)})
    cfg = [dict(c, effect_id=f"e{i}") for i, c in enumerate(synth.PRESETS["Robot Voice"])]
    eng.build_chain(cfg)                                   # live chain: taps + effects + warm-up at 256
    assert set(eng.effects_map) == {"e0", "e1", "e2"}
    eng.update_param("e2", "feedback", 0.9)                # SmoothParam route (engine.py:139-143)
    assert eng.effects_map["e2"].feedback.target == 0.9
    eng.update_param("e2", "mix_wet", 0.25)                # setter route
    assert eng.effects_map["e2"].mix_wet == 0.25
    ws = FakeSocket()
    asyncio.run(eng.process_wav_file(url, ws))
    assert len(ws.sent) == 1 and not eng.is_processing_file
    reply = json.loads(ws.sent[0])
    assert reply["type"] == "file_processed" and reply["sample_rate"] == fs and reply["original_b64"] == url
    assert len(reply["original_samples"]) == n and len(reply["processed_samples"]) == n
    rfs, out_pcm = scipy.io.wavfile.read(io.BytesIO(base64.b64decode(reply["processed_b64"].split(",")[1])))
    assert rfs == fs and out_pcm.shape == (n, 2) and out_pcm.dtype == np.int16
    audio = pcm.astype(np.float32) / np.float32(32768.0)
    mono = orc.mono_downmix(audio)
    want = orc.quantize_i16(np.clip(orc.run_file_path(synth.PRESETS["Robot Voice"], mono, fs), -1.0, 1.0))
    assert np.max(np.abs(out_pcm.astype(np.int32) - want.astype(np.int32))) <= 1
    assert np.max(np.abs(np.array(reply["original_samples"], np.float32) - mono[:, 0])) == 0.0


@pytest.mark.parametrize("name", NATIVE)
def test_blocks_of_whole_clip_size_continue_the_clip_like_the_reference(ab, orc, name):
    """The reference carries its rings from call to call at ANY block size (core.py:123-161).  Here the
    first long block takes the whole-clip kernels, which do not write their lines back; when a second
    block follows without prepare(), the first is replayed through the streaming kernel and the clip
    goes on -- results as if it had been one call (block-wise processing with blocks of 2048
    frames or more used to raise on the second block)."""
    from audioblocks.engine import file_chain
    cfg = synth.PRESETS[name]
    bs, nblk = 4096, 3
    x = synth.clip(21, bs * nblk, 2)
    chain = file_chain(cfg, 48000, channels_in=2, blocksize=bs)       # built and warmed at the block size: no re-prepare
    got = np.zeros((bs * nblk, 2), np.float32)
    for k in range(nblk):
        chain.process(x[k * bs:(k + 1) * bs], got[k * bs:(k + 1) * bs])
    ref = orc.build_chain(cfg, 48000, ci=2, bs=bs)
    ref.warmup()
    want = np.zeros_like(got)
    for k in range(nblk):
        ref.process(x[k * bs:(k + 1) * bs], want[k * bs:(k + 1) * bs])
    check(got, want, exact=(name == "Slapback Echo"), what=name)


def test_single_effect_second_block_and_reprepare(ab, orc):
    fx = ab.StereoDelayEffect()
    fx.prepare(48000, 2, 2, 4096)
    x = synth.batch(1, 2, 4096)
    out = np.zeros((2, 4096, 2), np.float32)
    fx.process_into(x[0], out[0])
    fx.process_into(x[1], out[1])                           # continues the lines of the first block
    o = orc.ODelay()
    o.prepare(48000, 2, 2, 4096)
    want = np.zeros_like(out)
    o.process_into(x[0], want[0])
    o.process_into(x[1], want[1])
    check(out.reshape(-1, 2), want.reshape(-1, 2), what="second block")
    fx.prepare(48000, 2, 2, 4096)                           # fresh lines again
    fx.process_into(x[1], out[1])
    o.prepare(48000, 2, 2, 4096)
    o.process_into(x[1], want[1])
    check(out[1], want[1], what="after prepare")


@pytest.mark.parametrize("fs", [8000, 192000])
def test_sample_rates_the_whole_clip_kernels_refuse_fall_back_to_streaming(ab, orc, fs):
    """At 8 kHz the default reverb's shortest comb is 235 samples (< the smallest tile), at 192 kHz its
    rings exceed shared memory: aes_chain_plan_create answers AES_ERR_UNSUPPORTED and the host side
    takes the streaming kernel instead of raising (such files never got a reply)."""
    from audioblocks.engine import file_chain
    cfg = synth.PRESETS["Rain Delay"]
    n = 9000 if fs == 8000 else 20000
    x = synth.clip(5, n, 1, fs)
    got = np.zeros((n, 2), np.float32)
    file_chain(cfg, fs, channels_in=1).process(x, got)
    check(got, orc.run_file_path(cfg, x, fs), what=("file", fs))
    xb = synth.batch(6, 3, n, 2, fs)
    yb = file_chain(cfg, fs, channels_in=2).process_batch(xb)
    for b in range(3):
        check(yb[b], orc.run_file_path(cfg, xb[b], fs), what=("batch", fs, b))
    q = np.zeros((3, n, 2), np.int16)
    file_chain(cfg, fs, channels_in=2).process_batch(xb, q)
    wq = (np.clip(yb, -1.0, 1.0) * np.float32(32767.0)).astype(np.int16)
    assert np.array_equal(q, wq)


@pytest.mark.parametrize("name", NATIVE)
def test_live_256_frame_blocks_carry_state_like_the_reference(ab, orc, name):
    """engine.py:38-65,156-163: chain built at blocksize 256, warmed up, then driven block by block;
    rings, filter state, gate gain and octaver phase are carried from call to call."""
    cfg = synth.PRESETS[name]
    ours = ab.EffectsChain(48000, 1, 2, 256)
    for c in cfg:
        ours.add(ab.engine.make_effect(c))
    ours.warmup()
    ref = orc.build_chain(cfg, 48000, ci=1, bs=256)
    ref.warmup()
    x = synth.clip(33, 256 * 40, 1)
    x = np.ascontiguousarray(np.roll(x, 3000, axis=0))
    for k in range(40):
        blk = np.ascontiguousarray(x[256 * k:256 * (k + 1)])
        got, want = np.zeros((256, 2), np.float32), np.zeros((256, 2), np.float32)
        ours.process(blk, got)
        ref.process(blk, want)
        assert np.max(np.abs(got - want)) <= 1e-5, (name, k)


def test_exact_1024_frames_keeps_the_warmup_state(ab, orc):
    """N == 1024: no re-prepare, the warm-up state of every block carries over (SURVEY 3.1 edge);
    pinned by a golden vector from the reference."""
    z, _ = goldens.load("presets")
    x = synth.clip(9, 1024, 1, 48000)
    y = run_file(ab, synth.PRESETS["Robot Voice"], x, 48000)
    check(y, z["n1024_Robot_Voice"], what="n1024")


def test_streaming_then_reprepare_then_whole_clip(ab, orc):
    cfg = synth.PRESETS["Rain Delay"]
    ours = ab.EffectsChain(48000, 2, 2, 512)
    for c in cfg:
        ours.add(ab.engine.make_effect(c))
    ref = orc.build_chain(cfg, 48000, ci=2, bs=512)
    x = synth.clip(14, 512 * 6, 2)
    for k in range(6):                                   # dirty the lines through the streaming path
        blk = np.ascontiguousarray(x[512 * k:512 * (k + 1)])
        got, want = np.zeros((512, 2), np.float32), np.zeros((512, 2), np.float32)
        ours.process(blk, got)
        ref.process(blk, want)
        assert np.max(np.abs(got - want)) <= 1e-5, k
    big = synth.clip(15, 50000, 2)                       # new frame count: re-prepare, whole-clip kernel
    got, want = np.zeros((50000, 2), np.float32), np.zeros((50000, 2), np.float32)
    ours.process(big, got)
    ref.process(big, want)
    check(got, want, what="whole clip after streaming")


@pytest.mark.parametrize("fs", [44100, 40000])
@pytest.mark.parametrize("name", NATIVE)
def test_presets_at_other_sample_rates(ab, orc, name, fs):
    """44.1 kHz takes the second compile-time reverb topology and misaligned delay lags; 40 kHz has
    no baked topology, so the reverb shapes run with descriptor-driven lengths (same kernels)."""
    cfg = synth.PRESETS[name]
    x = synth.clip(51, 40000, 2, fs)
    check(run_file(ab, cfg, x, fs), orc.run_file_path(cfg, x, fs), exact=(name == "Slapback Echo"), what=(name, fs))


REVERB_VARIANTS = synth.REVERB_VARIANTS


@pytest.mark.parametrize("variant", sorted(REVERB_VARIANTS))
def test_reverb_parameter_corners(ab, orc, variant):
    cfg = [{"type": "reverb", "params": dict(REVERB_VARIANTS[variant])}]
    x = synth.clip(61, 60000, 2, 48000)
    check(run_file(ab, cfg, x, 48000), orc.run_file_path(cfg, x, 48000), what=variant)


DELAY_VARIANTS = synth.DELAY_VARIANTS


@pytest.mark.parametrize("variant", sorted(DELAY_VARIANTS))
def test_delay_parameter_corners(ab, orc, variant):
    cfg = [{"type": "delay", "params": dict(DELAY_VARIANTS[variant])}]
    x = synth.clip(62, 150000, 2, 48000)
    check(run_file(ab, cfg, x, 48000), orc.run_file_path(cfg, x, 48000), what=variant)


@pytest.mark.parametrize("B", [1, 2, 297, 593])
def test_batch_sizes_around_the_grid(ab, orc, B):
    """One clip, fewer clips than CTAs, one more than a full grid (296 resident CTAs), two grids + 1."""
    from audioblocks.engine import file_chain
    cfg = synth.PRESETS["Rain Delay"]
    n = 6000
    base = synth.batch(70, 3, n)
    x = np.ascontiguousarray(base[np.arange(B) % 3])
    y = file_chain(cfg, 48000, channels_in=2).process_batch(x)
    want = [orc.run_file_path(cfg, base[k], 48000) for k in range(3)]
    for b in range(B):
        check(y[b], want[b % 3], what=(B, b))


@pytest.mark.parametrize("variant", sorted(synth.OCTAVER_VARIANTS))
def test_octaver_parameter_corners(ab, orc, variant):
    cfg = [{"type": "octaver", "params": dict(synth.OCTAVER_VARIANTS[variant])}]
    x = synth.clip(63, 100000, 2, 48000)
    check(run_file(ab, cfg, x, 48000), orc.run_file_path(cfg, x, 48000), what=variant)


def test_spectral_whole_file_with_partial_gating_pairs_and_odd_batches(ab, orc):
    """Whole-file SpectralFilter on clips scaled so that about half of the bins fall under the gate:
    the output (leakage of the gated bins into the zero half of the frame) is then far above rounding
    noise, and clips travel two per complex transform (odd batch: the last one alone).  A bin whose
    magnitude sits on the threshold may flip between the f32 device FFT and numpy's, so the bar is
    relative: 1 % of the output's peak; a pairing mix-up would be an error of order one."""
    from audioblocks.engine import file_chain
    cfg = [{"type": "spectral", "params": {"threshold_db": -40.0, "reduction": 0.1, "smoothing": 0.5}}]
    for B, n in ((3, 20011), (2, 16384), (1, 9000)):
        x = synth.batch(31, B, n) * np.float32(4e-4)
        x[0] *= np.float32(3.0)                         # a louder partner inside the first pair
        y = file_chain(cfg, 48000, channels_in=2).process_batch(x)
        for b in range(B):
            want = orc.run_file_path(cfg, x[b], 48000)
            peak = float(np.max(np.abs(want)))
            assert peak > 1e-6, (B, n, b, peak)         # the case really exercises the gate
            assert np.max(np.abs(y[b] - want)) <= 1e-2 * peak, (B, n, b, peak)


@pytest.mark.parametrize("seed", range(10 * int(__import__("os").environ.get("AES_FUZZ_SCALE", "1"))))
def test_convolution_reverb_random_sizes(ab, orc, seed):
    rng = np.random.default_rng(8100 + seed)
    log2n = int(rng.choice([8, 11, 14]))
    taps = int(rng.integers(1, {8: 700, 11: 6000, 14: 40000}[log2n]))
    n = int(rng.integers(1, 60000))
    B = int(rng.integers(1, 4))
    ir = orc.synthetic_ir(taps, rt60=float(rng.uniform(0.01, 0.5)))
    x = synth.batch(700 + seed, B, n)
    dry, wet = float(rng.uniform(0.0, 1.0)), float(rng.uniform(0.1, 1.0))
    y = ab.ConvolutionReverbEffect(ir, dry, wet, block_log2=log2n).process_batch(x)
    for b in range(B):
        want = np.zeros_like(x[b])
        orc.OConvReverb(ir, dry, wet).process_into(x[b], want)
        check(y[b], want, what=(seed, log2n, taps, n, b))


RACE_CHAINS = dict(synth.PRESETS)
RACE_CHAINS.pop("Clean Noise Removal")
RACE_CHAINS["c2-biquads"] = [
    {"type": "filter", "params": {"filter_type": 0, "cutoff_hz": 8000, "q": 0.707}},
    {"type": "filter", "params": {"filter_type": 1, "cutoff_hz": 80, "q": 0.707}},
    {"type": "filter", "params": {"filter_type": 2, "cutoff_hz": 1000, "q": 0.8}},
    {"type": "filter", "params": {"filter_type": 3, "cutoff_hz": 1000, "q": 1.0, "gain_db": 6.0}}]
RACE_CHAINS["gate-filter-gate"] = [
    {"type": "gate", "params": {"threshold_db": -35}}, {"type": "filter", "params": {"cutoff_hz": 300.0}},
    {"type": "gate", "params": {"threshold_db": -25}}]
RACE_CHAINS["filter-dist-filter"] = [
    {"type": "filter", "params": {"cutoff_hz": 2000.0}}, {"type": "distortion", "params": {"drive": 3.0}},
    {"type": "filter", "params": {"filter_type": 1, "cutoff_hz": 200.0}}]


@pytest.mark.parametrize("name", sorted(RACE_CHAINS))
def test_copies_of_one_clip_come_out_identical_on_every_cta(ab, name):
    """Race detector for the barrier-light kernels: 900 copies of two clips run on 296 CTAs at
    different phases (three waves, two CTAs sharing each SM); every copy must match bit for bit."""
    from audioblocks.engine import file_chain
    cfg = RACE_CHAINS[name]
    n, B = 12000, 900
    base = synth.batch(95, 2, n)
    x = np.ascontiguousarray(base[np.arange(B) % 2])
    chain = file_chain(cfg, 48000, channels_in=2)
    for rep in range(3):
        y = chain.process_batch(x) if rep == 0 else file_chain(cfg, 48000, channels_in=2).process_batch(x)
        for k in range(2):
            same = (y[k::2] == y[k]).all(axis=(1, 2))
            assert same.all(), (name, rep, k, int(np.argmin(same)))


def test_host_cache_release_and_reuse(ab, orc):
    """The host pipeline keeps its staging per thread across plans; releasing it must leave the next
    call working (and re-allocating)."""
    from audioblocks import _native
    from audioblocks.engine import file_chain
    cfg = synth.PRESETS["Rain Delay"]
    x = synth.batch(12, 3, 9000)
    y1 = file_chain(cfg, 48000, channels_in=2).process_batch(x)
    _native.release_host_cache()
    _native.release_host_cache()                      # idempotent
    y2 = file_chain(cfg, 48000, channels_in=2).process_batch(x)
    assert np.array_equal(y1, y2)
    check(y2[1], orc.run_file_path(cfg, x[1], 48000), what="after release")
