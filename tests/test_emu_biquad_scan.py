"""Time-parallel biquad cascade (csrc/aes_biquad_scan.cuh: one CTA per tile, decoupled
look-back across CTAs) on the CPU emulator, against the oracle's sequential DF-I loop."""
import ctypes as C

import numpy as np
import pytest

import emu
import synth
from oracle import oracle as orc

CASCADE = [{"type": "filter", "params": {"filter_type": 0, "cutoff_hz": 8000, "q": 0.707}},
           {"type": "filter", "params": {"filter_type": 1, "cutoff_hz": 80, "q": 0.707}},
           {"type": "filter", "params": {"filter_type": 2, "cutoff_hz": 1000, "q": 0.8}},
           {"type": "filter", "params": {"filter_type": 3, "cutoff_hz": 1000, "q": 1.0, "gain_db": 6.0}}]
STRESS = [{"type": "filter", "params": {"filter_type": 0, "cutoff_hz": 40, "q": 5.0}},
          {"type": "filter", "params": {"filter_type": 1, "cutoff_hz": 20, "q": 0.707}}]


def coeffs_of(cfg, fs, n):
    co = []
    for c in cfg:
        f = orc.OFilter(**c["params"])
        f.prepare(fs, 2, 2, n)
        co += list(f.coeffs())
    return np.array(co, np.float64)


def scan(x, co, n_stages, skip=0, state=None):
    L = emu.lib()
    L.emu_biquad_scan.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong, C.c_longlong, C.c_int, C.c_void_p,
                                  C.c_void_p, C.c_int]
    y = np.full_like(x, 7.0)
    sp = state.ctypes.data if state is not None else None
    assert L.emu_biquad_scan(x.ctypes.data, y.ctypes.data, x.shape[0], x.shape[1], n_stages, co.ctypes.data, sp, skip) == 0
    return y


@pytest.mark.parametrize("cfg", [CASCADE, STRESS])
@pytest.mark.parametrize("n,skip", [(1000, 0), (1024, 0), (9001, 0), (40 * 1024 + 5, 1000), (37 * 1024 + 5, 35)])
def test_scan_matches_sequential_reference_loop(cfg, n, skip):
    """skip > 0 hides inclusive prefixes from the look-back so the aggregate path and the
    multi-window (> 32 predecessors) path are exercised too."""
    x = synth.batch(3, 2, n)
    y = scan(x, coeffs_of(cfg, 48000, n), len(cfg), skip)
    for b in range(2):
        want = orc.run_file_path(cfg, x[b], 48000)
        mx, snr = synth.err_stats(y[b], want)
        assert mx <= 1e-5 * max(1.0, float(np.abs(want).max())) and snr >= 100.0, (n, skip, b, mx, snr)


def lookback_depth(co, n_stages):
    L = emu.lib()
    L.emu_biquad_lookback_depth.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
    out = np.zeros(n_stages, np.int32)
    assert L.emu_biquad_lookback_depth(n_stages, co.ctypes.data, out.ctypes.data) == 0
    return out.tolist()


def test_truncated_lookback_depth_and_agreement_with_the_chained_lookback():
    """A stable biquad forgets: once |A^(1024 i)| < 2^-44 the look-back stops at i tiles and sums
    aggregates only (no chain across tiles).  The depths follow the pole radius; the chained path
    (skip > 0 forces it) and the truncated one agree to an f32 ulp; a filter that remembers more
    than a 256-tile window keeps the chained look-back."""
    n = 40 * 1024 + 5
    x = synth.batch(3, 1, n)
    co = coeffs_of(CASCADE, 48000, n)
    depth = lookback_depth(co, 4)
    assert depth[0] == 1 and 3 <= depth[1] <= 6 and all(1 <= d <= 6 for d in depth), depth
    y_trunc = scan(x, co, 4, 0)
    y_chain = scan(x, co, 4, 1000)
    assert np.max(np.abs(y_trunc - y_chain)) <= 2.4e-7 * max(1.0, float(np.abs(y_chain).max()))
    co_s = coeffs_of(STRESS, 48000, n)
    d_s = lookback_depth(co_s, 2)
    assert 30 <= d_s[0] <= 120 and 2 <= d_s[1] <= 40, d_s
    slow = coeffs_of([{"type": "filter", "params": {"filter_type": 0, "cutoff_hz": 20, "q": 10.0}}], 96000, n)
    assert lookback_depth(slow, 1) == [0]


def test_scan_honours_a_carried_dfi_state():
    """filter.py:17-20: the kernel starts from state[c] = [x1, x2, y1, y2]."""
    n = 9000
    x = synth.batch(8, 1, n)
    cfg = CASCADE[:2]
    co = coeffs_of(cfg, 48000, n)
    rng = np.random.default_rng(1)
    st = (0.1 * rng.standard_normal((2, 2, 4))).astype(np.float32)
    y = scan(x, co, 2, 0, st.astype(np.float64))
    L = orc.lib()
    cur = x[0].copy()
    for s in range(2):
        out = np.zeros_like(cur)
        L.orc_biquad_kernel(orc._p(cur), orc._p(out), n, 2, *co[5 * s:5 * s + 5], orc._p(st[s].copy()))
        cur = out
    mx, snr = synth.err_stats(y[0], cur)
    assert mx <= 1e-5 and snr >= 100.0, (mx, snr)
