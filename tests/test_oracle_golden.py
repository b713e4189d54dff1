"""Pins the CPU oracle (oracle/) to the reference: every golden vector was
produced by running the unmodified reference (tests/golden/make_golden.py).
The reference kernels are numba fastmath=True, so its own output is only
defined up to LLVM's reassociation/contraction; index-only paths must be
exact, floating-point recurrences agree to well inside the 1e-5 contract."""
import ctypes as C

import numpy as np
import pytest

import goldens
import synth
from oracle import oracle as orc

TOL = 2e-6          # oracle (strict IEEE order) vs reference (fastmath LLVM)


def _f(a):
    return np.ascontiguousarray(a, np.float32)


@pytest.fixture(scope="module")
def K():
    return goldens.load("kernels")


def test_delay_kernel_golden(K):
    z, meta = K
    L = orc.lib()
    for i in range(5):
        m = meta[f"delay{i}"]
        x, buf = _f(z[f"delay{i}_x"]), _f(z[f"delay{i}_buf0"]).copy()
        wet = np.zeros_like(x)
        w1 = L.orc_delay_kernel(orc._p(buf), m["w0"], m["size"], orc._p(x), 1, orc._p(wet), 1,
                                x.shape[0], m["dS"], m["fb"])
        assert w1 == m["w1"]
        if m["fb"] == 0.0:
            assert np.array_equal(wet, z[f"delay{i}_wet"])           # pure index path: bit-exact
            assert np.array_equal(buf, z[f"delay{i}_buf1"])
        assert synth.err_stats(wet, z[f"delay{i}_wet"])[0] <= TOL
        assert synth.err_stats(buf, z[f"delay{i}_buf1"])[0] <= TOL


def test_biquad_kernel_golden(K):
    z, meta = K
    L = orc.lib()
    for i in range(5):
        m = meta[f"biquad{i}"]
        x, st = _f(z[f"biquad{i}_x"]), _f(z[f"biquad{i}_st0"]).copy()
        y = np.zeros_like(x)
        L.orc_biquad_kernel(orc._p(x), orc._p(y), x.shape[0], 2, *m["coeffs"], orc._p(st))
        scale = max(1.0, float(np.max(np.abs(z[f"biquad{i}_y"]))))
        assert synth.err_stats(y, z[f"biquad{i}_y"])[0] <= TOL * scale
        assert np.max(np.abs(st - z[f"biquad{i}_st1"])) <= TOL * scale


def test_pitch_kernel_golden(K):
    z, meta = K
    L = orc.lib()
    for i in range(4):
        m = meta[f"pitch{i}"]
        x, buf = _f(z[f"pitch{i}_x"]), _f(z[f"pitch{i}_buf0"]).copy()
        y = np.zeros_like(x)
        w, ph = C.c_int64(m["w0"]), C.c_double(m["ph0"])
        L.orc_pitch_shift_kernel(orc._p(buf), C.byref(w), m["size"], orc._p(x), 1, orc._p(y), 1,
                                 x.shape[0], C.byref(ph), m["step"])
        assert w.value == m["w1"]
        assert abs(ph.value - m["ph1"]) < 1e-12
        assert np.array_equal(buf, z[f"pitch{i}_buf1"])               # ring holds raw input
        assert synth.err_stats(y, z[f"pitch{i}_y"])[0] <= TOL


def test_reverb_kernels_golden(K):
    z, meta = K
    L = orc.lib()
    for i in range(3):
        m = meta[f"pure{i}"]
        x, buf = _f(z[f"pure{i}_x"]), _f(z[f"pure{i}_buf0"]).copy()
        y = np.zeros_like(x)
        w1 = L.orc_pure_delay_kernel(orc._p(buf), m["w0"], m["size"], orc._p(x), 1, orc._p(y), 1,
                                     x.shape[0], m["dS"])
        assert w1 == m["w1"]
        assert np.array_equal(y, z[f"pure{i}_y"]) and np.array_equal(buf, z[f"pure{i}_buf1"])
    for i in range(4):
        m = meta[f"comb{i}"]
        x, buf = _f(z[f"comb{i}_x"]), _f(z[f"comb{i}_buf0"]).copy()
        y = np.zeros_like(x)
        lp = C.c_double(m["lp0"])
        w1 = L.orc_comb_damped_kernel(orc._p(buf), m["w0"], m["L"] + 1, orc._p(x), 1, orc._p(y), 1,
                                      x.shape[0], m["L"], m["g"], m["h"], C.byref(lp))
        assert w1 == m["w1"]
        assert abs(lp.value - m["lp1"]) <= TOL
        assert synth.err_stats(y, z[f"comb{i}_y"])[0] <= TOL
        assert synth.err_stats(buf, z[f"comb{i}_buf1"])[0] <= TOL
    for i in range(3):
        m = meta[f"ap{i}"]
        x, buf = _f(z[f"ap{i}_x"]), _f(z[f"ap{i}_buf0"]).copy()
        y = np.zeros_like(x)
        w1 = L.orc_allpass_kernel(orc._p(buf), m["w0"], m["L"] + 1, orc._p(x), 1, orc._p(y), 1,
                                  x.shape[0], m["L"], m["a"])
        assert w1 == m["w1"]
        assert synth.err_stats(y, z[f"ap{i}_y"])[0] <= TOL
        assert synth.err_stats(buf, z[f"ap{i}_buf1"])[0] <= TOL


def test_gate_kernel_golden(K):
    z, meta = K
    L = orc.lib()
    for i in range(2):
        m = meta[f"gate{i}"]
        x = _f(z[f"gate{i}_x"])
        y = np.zeros_like(x)
        g1 = L.orc_gate_kernel(orc._p(x), orc._p(y), x.shape[0], 2, m["g0"], m["thr"], m["att"], m["rel"])
        assert abs(g1 - m["g1"]) < 1e-12
        assert synth.err_stats(y, z[f"gate{i}_y"])[0] <= TOL


def test_blocks_golden():
    z, meta = goldens.load("blocks")
    assert len(meta) >= 20
    for name, m in meta.items():
        x = goldens.block_input(m)
        y = orc.run_file_path(m["config"], x, m["fs"])
        want = z[name + "_y"]
        mx, snr = synth.err_stats(y, want)
        scale = max(1.0, float(np.max(np.abs(want))))
        assert mx <= TOL * scale, (name, mx, snr)
        if name == "delay_fb0":
            assert np.array_equal(y, want), name                       # index-only path: bit-exact


def test_presets_golden():
    z, meta = goldens.load("presets")
    mono = _f(z["rain_mono"])
    syn = goldens.syn_input(meta["syn"])
    for name, key in meta["presets"].items():
        cfg = synth.PRESETS[name]
        y = orc.run_file_path(cfg, mono, meta["rain"]["fs"])
        mx, snr = synth.err_stats(y, z[f"rain_{key}"])
        assert mx <= TOL, (name, "rain", mx, snr)
        q = orc.quantize_i16(y)
        # int16 truncation is exact given equal f32; the f32 may differ in the last bits
        assert np.max(np.abs(q.astype(np.int32) - z[f"rain_{key}_i16"].astype(np.int32))) <= 1
        y2 = orc.run_file_path(cfg, syn, 48000)
        mx2, snr2 = synth.err_stats(y2, z[f"syn_{key}"])
        assert mx2 <= TOL, (name, "syn", mx2, snr2)
    x1024 = synth.clip(9, 1024, 1, 48000)
    y = orc.run_file_path(synth.PRESETS["Robot Voice"], x1024, 48000)
    assert synth.err_stats(y, z["n1024_Robot_Voice"])[0] <= TOL


def test_quantize_exact():
    rng = np.random.default_rng(3)
    y = (1.2 * rng.uniform(-1, 1, (4096, 2))).astype(np.float32)
    want = (np.clip(y, -1.0, 1.0) * 32767).astype(np.int16)            # engine.py:104-105
    assert np.array_equal(orc.quantize_i16(y), want)


def test_c_chain_driver_matches_block_wrappers():
    """orc_chain_batch (the timed baseline) == the block-wrapper route."""
    x = synth.batch(0, 3, 20000)
    for name in ("Rain Delay", "Robot Voice", "Guitar Filter", "Cathedral", "Slapback Echo"):
        cfg = synth.PRESETS[name]
        yb = orc.run_batch_c(cfg, x, 48000, threads=2)
        for b in range(3):
            y = orc.run_file_path(cfg, x[b], 48000)
            assert np.array_equal(yb[b], y), (name, b)
        yf = orc.run_batch_c(cfg, x, 48000, threads=2, fast=True)
        assert synth.err_stats(yf, yb)[0] <= TOL, name
