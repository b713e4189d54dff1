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


# ---- the sequential batch kernel (csrc/aes_biquad_seq.cuh): one thread per (clip, segment) ----------------
def seq(x, co, n_stages, K, warm, state=None, max_ctas=0):
    L = emu.lib()
    L.emu_biquad_seq.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong, C.c_longlong, C.c_int, C.c_void_p, C.c_void_p,
                                 C.c_int, C.c_longlong, C.c_int]
    y = np.full_like(x, 7.0)
    sp = state.ctypes.data if state is not None else None
    assert L.emu_biquad_seq(x.ctypes.data, y.ctypes.data, x.shape[0], x.shape[1], n_stages, co.ctypes.data, sp, K, warm,
                            max_ctas) == 0
    return y


@pytest.mark.parametrize("n,K", [(2, 1), (62, 1), (1000, 1), (4096, 1), (50000, 3), (70002, 5)])
def test_sequential_batch_kernel_matches_the_reference_loop(n, K):
    """Whole clips per thread (K = 1: the reference's loop itself) and clips cut into segments that warm up over
    the cascade's memory (the scan's look-back depths); ragged lengths, more threads than one warp."""
    x = synth.batch(5, 37, n)
    co = coeffs_of(CASCADE, 48000, n)
    warm = 1024 * sum(lookback_depth(co, 4))
    y = seq(x, co, 4, K, warm)
    for b in (0, 17, 36):
        want = orc.run_file_path(CASCADE, x[b], 48000)
        mx, snr = synth.err_stats(y[b], want)
        assert mx <= 2e-7 * max(1.0, float(np.abs(want).max())) and snr >= 120.0, (n, K, b, mx, snr)


@pytest.mark.parametrize("ns", [1, 2, 3])
def test_sequential_batch_kernel_stage_counts_and_in_place(ns):
    cfg = CASCADE[:ns]
    x = synth.batch(6, 3, 3000)
    co = coeffs_of(cfg, 48000, 3000)
    L = emu.lib()
    y = seq(x, co, ns, 1, 0)
    z = x.copy()
    assert L.emu_biquad_seq(z.ctypes.data, z.ctypes.data, 3, 3000, ns, co.ctypes.data, None, 1, 0, 0) == 0
    assert np.array_equal(y, z)                        # in place: every chunk is read before it is written
    want = orc.run_file_path(cfg, x[1], 48000)
    assert synth.err_stats(y[1], want)[0] <= 2e-7 * max(1.0, float(np.abs(want).max()))


def test_sequential_batch_kernel_honours_a_carried_dfi_state():
    cfg = CASCADE[:2]
    x = synth.batch(9, 2, 4000)
    co = coeffs_of(cfg, 48000, 4000)
    full = seq(x, co, 2, 1, 0)
    # the state after the first 1000 frames, computed by the oracle's filters, continues the clip
    st = np.zeros((2, 8))
    sig = x[0, :1000].copy()
    for s, c in enumerate(cfg):
        f = orc.OFilter(**c["params"])
        f.prepare(48000, 2, 2, 1000)
        out = np.zeros_like(sig)
        f.process_into(sig, out)
        st[s] = np.asarray(f.state, np.float64).reshape(-1)[:8]
        sig = out
    rest = seq(np.ascontiguousarray(x[:1, 1000:]), co, 2, 1, 0, st)
    # (the reference keeps its state as float32, filter.py:35-40: resuming from it is not the unbroken run to the bit)
    assert np.max(np.abs(rest[0] - full[0, 1000:])) <= 5e-6


@pytest.mark.timeout(300)
@pytest.mark.parametrize("ns,n,K", [(4, 1500, 1), (2, 40000, 3), (1, 999 * 2, 1)])
def test_sequential_batch_kernel_persistent_threads_take_several_items(ns, n, K):
    """One CTA for 150 (clip, segment) items: every thread runs two or three of them back to back through the same
    buffer ring and mbarriers (a phase that drifts between items shows up as a hang or as stale data)."""
    cfg = CASCADE[:ns]
    x = synth.batch(11, 150 // K, n)
    co = coeffs_of(cfg, 48000, n)
    warm = 1024 * sum(lookback_depth(co, ns))
    y = seq(x, co, ns, K, warm, max_ctas=1)
    assert np.array_equal(y, seq(x, co, ns, K, warm))                     # same bits as one item per thread
    for b in (0, x.shape[0] // 2, x.shape[0] - 1):
        want = orc.run_file_path(cfg, x[b], 48000)
        assert synth.err_stats(y[b], want)[0] <= 2e-7 * max(1.0, float(np.abs(want).max()))
