"""CPU ORACLE -- test infrastructure, NOT the product.

A restatement of the reference's offline effect-chain path
(javierdrp/audio-effects-simulator, ``src/audioblocks``) on the CPU: the
per-sample loops live in ``aes_oracle.c`` (plain C, one function per numba
kernel), the block wrappers and the chain protocol are restated here in
numpy.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this module; the
product package never does.

Parity status: the reference has no tests and no golden vectors
(SURVEY.md section 8c), so this oracle is pinned against outputs of the
reference itself: ``tests/golden/make_golden.py`` imports the Python
reference from ``/root/reference`` in the build container and commits the
vectors under ``tests/golden/``; ``tests/test_oracle_golden.py`` holds the
oracle to them.  Blocks with no reference implementation (distortion,
peaking biquad, convolution reverb) are OUR definitions: parity unpinned.

Every function cites the reference file:line it follows (paths relative to
``/root/reference``).
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS: dict[str, C.CDLL] = {}

_F = C.POINTER(C.c_float)
_i64 = C.c_int64


def _build():
    subprocess.check_call(["make", "-s", "-C", _HERE], stdout=subprocess.DEVNULL)


def lib(fast: bool = False) -> C.CDLL:
    """Load (building if needed) the strict checker or the fast baseline build."""
    name = "liboracle_fast.so" if fast else "liboracle.so"
    if name in _LIBS:
        return _LIBS[name]
    path = os.path.join(_HERE, name)
    src = os.path.join(_HERE, "aes_oracle.c")
    if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
        _build()
    L = C.CDLL(path)
    L.orc_delay_kernel.restype = _i64
    L.orc_delay_kernel.argtypes = [_F, _i64, _i64, _F, _i64, _F, _i64, _i64, _i64, C.c_double]
    L.orc_biquad_kernel.restype = None
    L.orc_biquad_kernel.argtypes = [_F, _F, _i64, _i64] + [C.c_double] * 5 + [_F]
    L.orc_pitch_shift_kernel.restype = None
    L.orc_pitch_shift_kernel.argtypes = [_F, C.POINTER(_i64), _i64, _F, _i64, _F, _i64, _i64,
                                         C.POINTER(C.c_double), C.c_double]
    L.orc_pure_delay_kernel.restype = _i64
    L.orc_pure_delay_kernel.argtypes = [_F, _i64, _i64, _F, _i64, _F, _i64, _i64, _i64]
    L.orc_comb_damped_kernel.restype = _i64
    L.orc_comb_damped_kernel.argtypes = [_F, _i64, _i64, _F, _i64, _F, _i64, _i64, _i64,
                                         C.c_double, C.c_double, C.POINTER(C.c_double)]
    L.orc_allpass_kernel.restype = _i64
    L.orc_allpass_kernel.argtypes = [_F, _i64, _i64, _F, _i64, _F, _i64, _i64, _i64, C.c_double]
    L.orc_gate_kernel.restype = C.c_double
    L.orc_gate_kernel.argtypes = [_F, _F, _i64, _i64] + [C.c_double] * 4
    L.orc_mix_clip.restype = None
    L.orc_mix_clip.argtypes = [_F, _i64, _F, _i64, _F, _i64, _i64, C.c_float, C.c_float]
    L.orc_mono_mean2.restype = None
    L.orc_mono_mean2.argtypes = [_F, _F, _i64]
    L.orc_octaver_mix.restype = None
    L.orc_octaver_mix.argtypes = [_F, _F, _F, _i64, _i64, C.c_float, C.c_float]
    L.orc_accumulate.restype = None
    L.orc_accumulate.argtypes = [_F, _F, _i64]
    L.orc_quantize_i16.restype = None
    L.orc_quantize_i16.argtypes = [_F, C.POINTER(C.c_int16), _i64]
    L.orc_distortion.restype = None
    L.orc_distortion.argtypes = [_F, _F, _i64, C.c_float, C.c_float]
    L.orc_chain_clip.restype = None
    L.orc_chain_clip.argtypes = [C.c_void_p, C.c_int32, _F, _F, _F, _i64]
    L.orc_chain_batch.restype = None
    L.orc_chain_batch.argtypes = [C.c_void_p, C.c_int32, _F, _F, _i64, _i64, C.c_int32]
    _LIBS[name] = L
    return L


def _p(a: np.ndarray, offset: int = 0):
    """float* to element `offset` of a C-contiguous f32 array."""
    assert a.dtype == np.float32 and a.flags.c_contiguous
    return C.cast(a.ctypes.data + 4 * offset, _F)


# --------------------------------------------------------------------------
# Block wrappers.  State layout and (re)allocation rules follow the reference.
# Parameters are constants: in the offline path every SmoothParam has
# current == target (core.py:56-59), so step_towards is a no-op.
# --------------------------------------------------------------------------

class ODelay:
    """delay.py:43-96 StereoDelayEffect + delay.py:24-41 DelayLine."""

    def __init__(self, max_delay_ms=1500.0, mix_dry=0.8, mix_wet=0.8, offset_ms=30.0,
                 delay_ms=375.0, feedback=0.2, fb_step=0.02, step_samples=2.0):
        self.max_delay_ms, self.mix_dry, self.mix_wet = max_delay_ms, mix_dry, mix_wet
        self.offset_ms, self.delay_ms, self.feedback = offset_ms, float(delay_ms), float(feedback)
        self.fs = 48000
        self.rings = [np.zeros(1, np.float32), np.zeros(1, np.float32)]
        self.w = [0, 0]

    def prepare(self, fs, ci, co, bs):
        self.fs = fs
        size = int(fs * self.max_delay_ms / 1000.0) + 1          # delay.py:33
        self.rings = [np.zeros(size, np.float32), np.zeros(size, np.float32)]
        self.w = [0, 0]

    def lags(self):
        """(size, dS_L, dS_R) with the reference's own expressions (delay.py:38-40,84)."""
        size = self.rings[0].shape[0]
        d_ms = (self.delay_ms, min(self.delay_ms + self.offset_ms, self.max_delay_ms - 1.0))
        out = []
        for ms in d_ms:
            dS = int(self.fs * ms / 1000.0)
            out.append(min(dS, size - 1))
        return size, out[0], out[1]

    def process_into(self, x, out, L=None):
        L = L or lib()
        n = x.shape[0]
        size, dL, dR = self.lags()
        wet = np.empty(n, np.float32)
        for c, dS in ((0, dL), (1, dR)):
            self.w[c] = L.orc_delay_kernel(_p(self.rings[c]), self.w[c], size, _p(x, c), 2,
                                           _p(wet), 1, n, dS, float(self.feedback))
            L.orc_mix_clip(_p(x, c), 2, _p(wet), 1, _p(out, c), 2, n,
                           np.float32(self.mix_dry), np.float32(self.mix_wet))


class OFilter:
    """filter.py:42-113 FilterEffect.  filter_type 3 (RBJ peaking, +gain_db) is
    OUR extension; the recurrence is the reference's."""

    def __init__(self, filter_type=0.0, cutoff_hz=1000.0, q=0.707, gain_db=0.0):
        self.filter_type, self.cutoff_hz, self.q = float(filter_type), float(cutoff_hz), float(q)
        self.gain_db = float(gain_db)
        self.state = np.zeros((1, 4), np.float32)
        self.fs = 48000.0

    def prepare(self, fs, ci, co, bs):
        self.fs = float(fs)
        if self.state.shape[0] != co:                              # filter.py:59-60
            self.state = np.zeros((co, 4), np.float32)

    def coeffs(self):
        """filter.py:62-98 (RBJ cookbook), normalised by a0."""
        w0 = 2.0 * math.pi * self.cutoff_hz / self.fs
        cw, sw = math.cos(w0), math.sin(w0)
        alpha = sw / (2.0 * self.q)
        t = int(round(self.filter_type))
        if t == 3:      # extension: peaking EQ
            A = 10.0 ** (self.gain_db / 40.0)
            b0, b1, b2 = 1 + alpha * A, -2 * cw, 1 - alpha * A
            a0, a1, a2 = 1 + alpha / A, -2 * cw, 1 - alpha / A
        elif t == 0:
            b0, b1, b2 = (1 - cw) / 2, 1 - cw, (1 - cw) / 2
            a0, a1, a2 = 1 + alpha, -2 * cw, 1 - alpha
        elif t == 1:
            b0, b1, b2 = (1 + cw) / 2, -(1 + cw), (1 + cw) / 2
            a0, a1, a2 = 1 + alpha, -2 * cw, 1 - alpha
        else:
            b0, b1, b2 = alpha, 0, -alpha
            a0, a1, a2 = 1 + alpha, -2 * cw, 1 - alpha
        return (b0 / a0, b1 / a0, b2 / a0, a1 / a0, a2 / a0)

    def process_into(self, x, out, L=None):
        L = L or lib()
        b0, b1, b2, a1, a2 = self.coeffs()
        L.orc_biquad_kernel(_p(x), _p(out), x.shape[0], x.shape[1], b0, b1, b2, a1, a2,
                            _p(self.state))


class OOctaver:
    """octaver.py:84-150 OctaverEffect."""

    def __init__(self, semitones=-12.0, mix=0.5, window_ms=40.0):
        self.semitones, self.mix, self.window_ms = float(semitones), float(mix), float(window_ms)
        self.buf = np.zeros(1, np.float32)
        self.w, self.phasor, self.size, self.fs = 0, 0.0, 1, 48000

    def prepare(self, fs, ci, co, bs):
        self.fs = fs
        req = max(int(fs * self.window_ms / 1000.0), 16)           # octaver.py:106
        if req != self.size:                                       # state kept otherwise
            self.size, self.buf, self.w, self.phasor = req, np.zeros(req, np.float32), 0, 0.0

    def step(self):
        return (1.0 - 2.0 ** (self.semitones / 12.0)) / self.size  # octaver.py:121-122

    def process_into(self, x, out, L=None):
        L = L or lib()
        n, ch = x.shape
        if ch > 1:
            mono = np.empty(n, np.float32)
            L.orc_mono_mean2(_p(x), _p(mono), n)
        else:
            mono = np.ascontiguousarray(x[:, 0])
        wet = np.zeros(n, np.float32)
        w, ph = _i64(self.w), C.c_double(self.phasor)
        L.orc_pitch_shift_kernel(_p(self.buf), C.byref(w), self.size, _p(mono), 1, _p(wet), 1, n,
                                 C.byref(ph), self.step())
        self.w, self.phasor = w.value, ph.value
        L.orc_octaver_mix(_p(x), _p(wet), _p(out), n, out.shape[1],
                          np.float32(1.0 - self.mix), np.float32(self.mix))


class OReverb:
    """reverb.py:72-277 ReverbEffect (Schroeder/Moorer network per side)."""

    def __init__(self, *, comb_times_ms=(29.7, 37.1, 41.1, 43.7), allpass_times_ms=(5.0, 1.7),
                 allpass_gain=0.6, jitter_ms=0.3, max_delay_ms=200.0, max_pre_delay_ms=100.0,
                 mix_dry=0.7, mix_wet=0.5, rt60_s=1.5, damp=0.3, pre_delay_ms=0.0,
                 step_samples=2.0, rt60_step=0.05, damp_step=0.02):
        self.comb_ms = tuple(float(v) for v in comb_times_ms)
        self.ap_ms = tuple(float(v) for v in allpass_times_ms)
        self.a, self.jitter = float(allpass_gain), float(jitter_ms)
        self.max_ms, self.max_pre_ms = float(max_delay_ms), float(max_pre_delay_ms)
        self.mix_dry, self.mix_wet = float(mix_dry), float(mix_wet)
        self.rt60, self.damp, self.pre_ms = float(rt60_s), float(damp), float(pre_delay_ms)
        self.fs = 48000
        self.sides = []

    def side_lengths(self, fs, jitter):
        """reverb.py:158-177: integer line lengths of one side."""
        combs = [max(1, int(fs * min(ms + jitter, self.max_ms - 1.0) / 1000.0)) for ms in self.comb_ms]
        aps = [max(1, int(fs * min(ms + jitter * 0.2, self.max_ms - 1.0) / 1000.0)) for ms in self.ap_ms]
        return combs, aps

    def prepare(self, fs, ci, co, bs):
        self.fs = int(fs)
        self.sides = []
        pre_size = max(1, int(self.fs * self.max_pre_ms / 1000.0) + 1)   # reverb.py:192-193
        for jit in (+self.jitter, -self.jitter):
            cl, al = self.side_lengths(self.fs, jit)
            self.sides.append({
                "comb": [{"L": n, "buf": np.zeros(n + 1, np.float32), "w": 0, "lp": 0.0} for n in cl],
                "ap": [{"L": n, "buf": np.zeros(n + 1, np.float32), "w": 0} for n in al],
                "pre": np.zeros(pre_size, np.float32), "pre_w": 0,
            })

    def gain(self, L_samples):
        return 10.0 ** (-3.0 * (float(L_samples) / float(self.fs)) / max(1e-3, self.rt60))  # :205-206

    def pre_dS(self):
        d = int(self.fs * self.pre_ms / 1000.0)                   # reverb.py:223-225
        size = self.sides[0]["pre"].shape[0]
        return min(d, size - 1)

    def process_into(self, x, out, L=None):
        L = L or lib()
        n = x.shape[0]
        pre_dS = self.pre_dS()
        pre, tmp, acc = (np.empty(n, np.float32) for _ in range(3))
        for s, side in enumerate(self.sides):
            ring = side["pre"]
            side["pre_w"] = L.orc_pure_delay_kernel(_p(ring), side["pre_w"], ring.shape[0],
                                                    _p(x, s), 2, _p(pre), 1, n, pre_dS)
            acc.fill(0.0)
            for c in side["comb"]:
                lp = C.c_double(c["lp"])
                c["w"] = L.orc_comb_damped_kernel(_p(c["buf"]), c["w"], c["L"] + 1, _p(pre), 1,
                                                  _p(tmp), 1, n, c["L"], self.gain(c["L"]),
                                                  self.damp, C.byref(lp))
                c["lp"] = lp.value
                L.orc_accumulate(_p(acc), _p(tmp), n)
            src, dst = acc, tmp
            for a in side["ap"]:
                a["w"] = L.orc_allpass_kernel(_p(a["buf"]), a["w"], a["L"] + 1, _p(src), 1,
                                              _p(dst), 1, n, a["L"], self.a)
                src, dst = dst, src
            L.orc_mix_clip(_p(x, s), 2, _p(src), 1, _p(out, s), 2, n,
                           np.float32(self.mix_dry), np.float32(self.mix_wet))


class OGate:
    """gate.py:45-90 NoiseGateEffect."""

    def __init__(self, threshold_db=-40.0, attack_ms=10.0, release_ms=100.0):
        self.th_db, self.att_ms, self.rel_ms = float(threshold_db), float(attack_ms), float(release_ms)
        self.gain, self.fs = 0.0, 48000.0

    def prepare(self, fs, ci, co, bs):
        self.fs = float(fs)

    def coeff(self, ms):
        t = max(1e-3, ms * 1e-3)
        return float(1.0 - np.exp(-2.2 / (t * self.fs)))           # gate.py:63-69

    def consts(self):
        return 10.0 ** (self.th_db / 20.0), self.coeff(self.att_ms), self.coeff(self.rel_ms)

    def process_into(self, x, out, L=None):
        L = L or lib()
        thr, att, rel = self.consts()
        self.gain = L.orc_gate_kernel(_p(x), _p(out), x.shape[0], x.shape[1], self.gain, thr, att, rel)


class OSpectral:
    """spectral.py:5-100 SpectralFilter (numpy FFT, restated as-is incl. its
    one-block latency; in whole-file mode hop == N, so the block emits the
    zero-padded half of a single 2N-point frame)."""

    def __init__(self, threshold_db=-40.0, reduction=0.5, smoothing=0.8):
        self.th_db, self.reduction, self.alpha = float(threshold_db), float(reduction), smoothing
        self._alloc(256)

    def _alloc(self, hop):
        self.hop, self.n_fft = hop, 2 * hop
        self.window = np.hanning(self.n_fft).astype(np.float32)
        self.inbuf = np.zeros(self.n_fft, np.float32)
        self.acc = np.zeros(self.n_fft, np.float32)
        self.mask = np.ones(self.n_fft // 2 + 1, np.float32)

    def prepare(self, fs, ci, co, bs):
        if bs != self.hop:                                         # spectral.py:34-42
            self._alloc(bs)

    def process_into(self, x, out, L=None):
        h = self.hop
        thr = 10.0 ** (self.th_db / 20.0)
        self.inbuf[:-h] = self.inbuf[h:]
        self.inbuf[-h:] = np.mean(x, axis=1)
        spec = np.fft.rfft(self.inbuf * self.window)
        mag, ph = np.abs(spec), np.angle(spec)
        cur = np.where(mag > thr, 1.0, self.reduction)
        self.mask = self.alpha * self.mask + (1.0 - self.alpha) * cur
        self.acc += np.fft.irfft(mag * self.mask * np.exp(1j * ph))
        for c in range(out.shape[1]):
            out[:, c] = self.acc[:h]
        self.acc[:-h] = self.acc[h:]
        self.acc[-h:] = 0.0


class ODistortion:
    """OUR definition (no reference implementation):
    out = clip((1-mix)*x + mix*tanh(drive*x), -1, 1), f32 elementwise."""

    def __init__(self, drive=4.0, mix=1.0):
        self.drive, self.mix = float(drive), float(mix)

    def prepare(self, fs, ci, co, bs):
        pass

    def process_into(self, x, out, L=None):
        L = L or lib()
        L.orc_distortion(_p(x), _p(out), x.size, np.float32(self.drive), np.float32(self.mix))


class OConvReverb:
    """OUR definition (no reference implementation; SURVEY 0.2 / BASELINE configs[3]):
    out[:, c] = clip(f32(dry)*x[:, c] + f32(wet)*f32(conv(x[:, c], ir[:, c])[:N]), -1, 1) with the
    convolution evaluated in float64 (scipy.signal.fftconvolve).  Parity unpinned by the reference."""

    def __init__(self, ir, mix_dry=0.7, mix_wet=0.5):
        self.ir = np.asarray(ir, np.float32).reshape(-1, 2)
        self.mix_dry, self.mix_wet = float(mix_dry), float(mix_wet)

    def prepare(self, fs, ci, co, bs):
        pass

    def process_into(self, x, out, L=None):
        from scipy.signal import fftconvolve
        n = x.shape[0]
        for c in range(2):
            wet = fftconvolve(x[:, c].astype(np.float64), self.ir[:, c].astype(np.float64))[:n].astype(np.float32)
            out[:, c] = np.clip(np.float32(self.mix_dry) * x[:, c] + np.float32(self.mix_wet) * wet, -1.0, 1.0)


def synthetic_ir(n_taps, fs=48000, rt60=1.5, seed=7):
    """rng(7) noise x 10^(-3 t / rt60) envelope, L/R independent, sum|h| = 4 per channel."""
    rng = np.random.default_rng(seed)
    t = np.arange(n_taps) / fs
    h = rng.standard_normal((n_taps, 2)) * (10.0 ** (-3.0 * t / rt60))[:, None]
    h *= 4.0 / np.abs(h).sum(axis=0, keepdims=True)
    return h.astype(np.float32)


KINDS = {"convreverb": OConvReverb, "delay": ODelay, "reverb": OReverb, "gate": OGate, "spectral": OSpectral,
         "octaver": OOctaver, "filter": OFilter, "distortion": ODistortion}


class OChain:
    """core.py:109-161 EffectsChain: ping-pong driver with re-prepare on a frame
    count change and a 2-block zero warm-up."""

    def __init__(self, fs, ci, co, bs):
        self.fs, self.ci, self.co, self.bs = fs, ci, co, bs
        self.fx = []

    def add(self, fx):
        fx.prepare(self.fs, self.ci, self.co, self.bs)
        self.fx.append(fx)

    def warmup(self):
        zi, zo = np.zeros((self.bs, self.ci), np.float32), np.zeros((self.bs, self.co), np.float32)
        for _ in range(2):
            self.process(zi, zo)

    def process(self, xin, out):
        n = xin.shape[0]
        if n != self.bs:
            self.bs = n
            for fx in self.fx:
                fx.prepare(self.fs, self.ci, self.co, n)
        a = np.zeros((n, self.co), np.float32)
        b = np.zeros((n, self.co), np.float32)
        if self.ci == 1 and self.co == 2:
            a[:, 0] = xin[:, 0]
            a[:, 1] = xin[:, 0]
        else:
            k = min(self.ci, self.co)
            a[:, :k] = xin[:, :k]
        for fx in self.fx:
            fx.process_into(a, b)
            a, b = b, a
        out[:, :] = a


def build_chain(config, fs, ci=1, co=2, bs=1024):
    """engine.py:86-98: construct the chain from a preset config list."""
    ch = OChain(fs, ci, co, bs)
    for cfg in config:
        cls = KINDS.get(cfg.get("type"))
        if cls is None:
            continue
        ch.add(cls(**cfg.get("params", {})))
    return ch


def run_file_path(config, x, fs):
    """engine.py:86-102: build at blocksize 1024, warm up, one whole-clip call.
    x is (N, ci) f32; returns (N, 2) f32 before the final clip/int16 stage."""
    x = np.ascontiguousarray(x, np.float32)
    ch = build_chain(config, fs, ci=x.shape[1])
    ch.warmup()
    out = np.zeros((x.shape[0], 2), np.float32)
    ch.process(x, out)
    return out


def mono_downmix(audio):
    """engine.py:81-84."""
    if audio.ndim > 1:
        return audio.mean(axis=1, keepdims=True)
    return audio.reshape(-1, 1)


def quantize_i16(y):
    """engine.py:104-105."""
    y = np.ascontiguousarray(y, np.float32)
    q = np.empty(y.shape, np.int16)
    lib().orc_quantize_i16(_p(y), q.ctypes.data_as(C.POINTER(C.c_int16)), y.size)
    return q


# --------------------------------------------------------------------------
# Whole-chain C driver (timed CPU baseline; also a second route to the same
# numbers for the tests).  Fresh delay/reverb state, carried octaver phase.
# --------------------------------------------------------------------------

class _Op(C.Structure):
    _fields_ = [("kind", C.c_int32), ("pad", C.c_int32), ("p", C.c_double * 32), ("q", _i64 * 32)]


def plan_ops(config, fs, n_frames):
    """Resolve a preset into the op table orc_chain_clip understands, replaying
    the build@1024 -> warm-up -> re-prepare@N protocol for the carried state."""
    ch = build_chain(config, fs, ci=2)
    ch.warmup()
    if n_frames != ch.bs:
        for fx in ch.fx:
            fx.prepare(fs, 2, 2, n_frames)
    ops = (_Op * max(1, len(ch.fx)))()
    for op, fx in zip(ops, ch.fx):
        if isinstance(fx, ODelay):
            size, dL, dR = fx.lags()
            op.kind = 1
            op.q[0], op.q[1], op.q[2] = size, dL, dR
            op.p[0], op.p[1], op.p[2] = fx.feedback, fx.mix_dry, fx.mix_wet
        elif isinstance(fx, OReverb):
            op.kind = 2
            op.q[0], op.q[1] = len(fx.comb_ms), len(fx.ap_ms)
            op.q[2], op.q[3] = fx.sides[0]["pre"].shape[0], fx.pre_dS()
            op.p[0], op.p[1], op.p[2], op.p[3] = fx.mix_dry, fx.mix_wet, fx.damp, fx.a
            for s, side in enumerate(fx.sides):
                for c, comb in enumerate(side["comb"]):
                    op.q[4 + 8 * s + c] = comb["L"]
                    op.p[4 + 8 * s + c] = fx.gain(comb["L"])
                for k, ap in enumerate(side["ap"]):
                    op.q[20 + 4 * s + k] = ap["L"]
        elif isinstance(fx, OFilter):
            op.kind = 3
            for i, v in enumerate(fx.coeffs()):
                op.p[i] = v
        elif isinstance(fx, OGate):
            op.kind = 4
            op.p[0], op.p[1], op.p[2] = fx.consts()
            op.p[3] = fx.gain
        elif isinstance(fx, OOctaver):
            op.kind = 5
            op.q[0], op.q[1] = fx.size, fx.w
            op.p[0], op.p[1], op.p[2] = fx.phasor, fx.step(), fx.mix
        elif isinstance(fx, ODistortion):
            op.kind = 6
            op.p[0], op.p[1] = fx.drive, fx.mix
        else:
            raise NotImplementedError(f"no C driver for {type(fx).__name__}")
    return ops, len(ch.fx)


def run_batch_c(config, x, fs, threads=1, fast=False):
    """x: (B, N, 2) f32 -> (B, N, 2) f32 through orc_chain_batch."""
    x = np.ascontiguousarray(x, np.float32)
    B, N, _ = x.shape
    ops, n = plan_ops(config, fs, N)
    y = np.empty_like(x)
    lib(fast).orc_chain_batch(C.cast(ops, C.c_void_p), n, _p(x), _p(y), B, N, threads)
    return y


# ---- plot-side analysis (assets/02_custom.js:65-154); float64 like the page's JavaScript ----------------
def chroma_from_magnitudes(mag, sample_rate, n_fft, threshold_scale=1.0):
    """assets/02_custom.js:65-106 `calculateChroma`.  `threshold_scale` moves the 15 % gate by a hair so a
    test can bracket the bins a float32 transform may put on the other side of it."""
    mag = np.asarray(mag, np.float64)
    chroma = np.zeros(12)
    threshold = mag.max() * 0.15 * threshold_scale
    for k in range(1, len(mag)):
        freq = k * (sample_rate / n_fft)
        if freq < 70:
            continue
        weighting = 1.0
        if freq > 800:
            weighting *= 0.5
        if freq > 1500:
            weighting *= 0.1
        if freq > 5000:
            continue
        if mag[k] * weighting < threshold:
            continue
        midi = 12 * np.log2(freq / 440) + 69
        nearest = np.floor(midi + 0.5)                  # Math.round
        if abs(midi - nearest) > 0.50:
            continue
        chroma[(int(nearest) % 12 + 12) % 12] += mag[k] * weighting
    v = chroma / (chroma.max() + 1e-9)
    return v * v * v


def spectrum_and_chroma(signal, sample_rate, threshold_scale=1.0):
    """assets/02_custom.js:108-154 `calculateSpectrumAndChroma` on one signal of n_fft samples:
    (freqs, magnitudesDB, chroma, peakFreq, magnitudesLin)."""
    x = np.asarray(signal, np.float64)
    n = len(x)
    i = np.arange(n)
    w = (0.35875 - 0.48829 * np.cos(2 * np.pi * i / (n - 1)) + 0.14128 * np.cos(4 * np.pi * i / (n - 1))
         - 0.01168 * np.cos(6 * np.pi * i / (n - 1)))
    mag = np.abs(np.fft.fft(x * w)[: n // 2 + 1])
    freqs = np.arange(n // 2 + 1) * (sample_rate / n)
    db = 20 * np.log10(mag / n + 1e-9)
    peak_mag, peak_freq = -np.inf, 0.0
    for k in range(n // 2 + 1):
        if freqs[k] > 60 and db[k] > peak_mag:
            peak_mag, peak_freq = db[k], freqs[k]
    return freqs, db, chroma_from_magnitudes(mag, sample_rate, n, threshold_scale), peak_freq, mag
