"""Chain driver of the drop-in `audioblocks` package.

Same public surface as the reference's src/audioblocks/core.py (SmoothParam
core.py:56-78, Effect :81-86, PlotDataTap :89-106, EffectsChain :109-161,
pick_devices :12-53); the signal path is the fused sm_100a kernel behind
libaesim.so instead of per-effect numpy/numba passes.
"""
from __future__ import annotations

import queue
import threading

import numpy as np

from . import _native

try:                                    # optional, exactly as in the reference (core.py:6-9)
    import sounddevice as sd
except (ImportError, OSError):
    sd = None


def pick_devices(ch_in=1, ch_out=2, in_hint=("usb", "mic"), out_hint=("system",)):
    """(in_idx, out_idx) preferring JACK, else Pulse, else PortAudio defaults
    (reference core.py:12-53).  Without sounddevice there is nothing to pick."""
    if sd is None:
        return None, None
    try:
        apis, devices = sd.query_hostapis(), sd.query_devices()
    except Exception:
        return None, None

    def api_id(tag):
        return next((i for i, a in enumerate(apis) if tag in a["name"]), None)

    def find(api, need_in, need_out, tokens):
        for i, d in enumerate(devices):
            name = d["name"].lower()
            if d["hostapi"] != api or not all(t.lower() in name for t in tokens):
                continue
            if need_in and d["max_input_channels"] < ch_in:
                continue
            if need_out and d["max_output_channels"] < ch_out:
                continue
            return i
        return None

    jack, pulse = api_id("JACK"), api_id("Pulse")
    if jack is not None:
        i, o = find(jack, True, False, in_hint), find(jack, False, True, out_hint)
        if i is not None and o is not None:
            return i, o
    if pulse is not None:
        p = next((i for i, d in enumerate(devices) if d["hostapi"] == pulse), None)
        if p is not None:
            return p, p
    return None, None


class SmoothParam:
    """A clamped target that `current` slews towards, one bounded step per block.
    The constructor does not clamp (reference core.py:57-59); set_target/nudge do."""

    def __init__(self, value, lo=-np.inf, hi=np.inf):
        self.current = float(value)
        self.target = float(value)
        self.lo = float(lo)
        self.hi = float(hi)
        self._lock = threading.Lock()

    def _clamped(self, v):
        return min(max(float(v), self.lo), self.hi)

    def set_target(self, v):
        with self._lock:
            self.target = self._clamped(v)

    def nudge(self, dv):
        with self._lock:
            self.target = self._clamped(self.target + float(dv))

    def step_towards(self, max_step=1.0):
        if max_step < 0:
            raise ValueError("max_step must be >= 0")
        with self._lock:
            gap = self.target - self.current
            self.current += min(max(gap, -max_step), max_step)
            return self.current


class Effect:
    """Base effect: prepare() (re)allocates, process_into() writes `out`."""

    def prepare(self, sample_rate: int, channels_in: int, channels_out: int, blocksize: int):
        pass

    def process_into(self, x_in: np.ndarray, out: np.ndarray) -> None:
        raise NotImplementedError


# Blocks of at least this many frames through freshly prepared delay lines take the whole-clip
# kernels; everything else (warm-up blocks, live 256-frame blocks, continuing a clip) takes the
# block-streaming kernel with carried state.
WHOLE_CLIP_MIN_FRAMES = 2048


class NativeEffect(Effect):
    """An effect executed by the CUDA library.  Subclasses resolve their parameters into
    aes_stage_desc records (`_stages`), size their delay-line state (`_blob_floats`,
    `_stream_fields`) and mirror the carried scalars on the host (`_advance`, `_absorb`); the
    signal never touches a CPU implementation."""

    _sr = 48000
    _dirty = False          # a non-silent block went through the streaming path: lines are not zero
    _stale = False          # a whole-clip call consumed the lines without writing them back
    _n_total = 0            # frames processed since prepare()
    _blob = None            # device memory of the delay lines (streaming path), allocated lazily
    _replay = None          # input + pre-call state of the last whole-clip call (see _replay_whole_clip)

    def _stages(self, frames: int) -> list:
        raise NotImplementedError

    def _blob_floats(self) -> int:
        return 0

    def _stream_fields(self, desc):
        """Fill the streaming-only descriptor fields (q[28] blob, q[29] frames so far, q[30] ring size)."""
        desc.q[29] = self._n_total

    def _reset_lines(self):
        """prepare() rebuilt the delay lines: forget everything that went through them."""
        self._dirty = False
        self._stale = False
        self._n_total = 0
        self._blob_zero = True
        self._replay = None

    def _stream_desc(self, frames: int):
        descs = self._stages(frames)
        n = self._blob_floats()
        if n:
            if self._blob is None or self._blob.n_floats != n:
                if self._blob is not None:
                    self._blob.close()
                self._blob = _native.DeviceBlob(n)
                self._blob_zero = False
            elif getattr(self, "_blob_zero", False):
                self._blob.zero()
                self._blob_zero = False
        for d in descs:
            if n:
                d.q[28] = self._blob.ptr
            self._stream_fields(d)
        return descs

    def _advance(self, frames: int, silent: bool, final=None):
        """Host-visible state after a whole-clip call of `frames`; `final` holds the stage's carried
        scalars read back from the device (None when nothing was run)."""
        self._n_total += frames
        if not silent:
            self._stale = True

    def _absorb(self, desc, frames: int, silent: bool):
        """Host-visible state after a streamed block: `desc` carries the final scalars."""
        self._n_total += frames
        if not silent:
            self._dirty = True

    def _at_rest(self) -> bool:
        """Nothing in this effect's memory can make sound: a silent block comes out silent."""
        return not (self._dirty or self._stale)

    def _rest_block(self, frames: int):
        """State after a silent block through an effect at rest, in closed form (no launch): the
        parameters took their per-block smoothing step in _stages(); lines stay zero."""
        self._absorb_rest(frames)

    def _absorb_rest(self, frames: int):
        self._n_total += frames

    def _require_fresh(self):
        if self._dirty or self._stale:
            raise _native.AesimError(
                f"{type(self).__name__}: the whole-clip CUDA path starts every call from freshly "
                "prepared delay lines; call prepare() (the chain does so whenever the frame count "
                "changes) before processing another block")

    def _require_usable(self):
        if self._stale:
            raise _native.AesimError(
                f"{type(self).__name__}: a whole-clip call left the delay lines behind (they are not written "
                "back) and its input was too large to keep for a replay; call prepare() before streaming "
                "further blocks through this effect")

    def process_into(self, x_in: np.ndarray, out: np.ndarray) -> None:
        run_native([self], self._sr, x_in, out)


# A whole-clip call keeps a copy of its input (up to this many floats) so that a caller who goes on
# processing block after block at the same size -- legal in the reference, whose rings carry over
# (core.py:123-161) -- can be served: the block is replayed through the streaming kernel first.
REPLAY_MAX_FLOATS = 64 << 20

_SNAP_TYPES = (int, float, bool, str, type(None), tuple)


def _snapshot(fx):
    """Host-visible scalar state of an effect (write pointers, phasors, gains, filter memories)."""
    snap = {}
    for k, v in fx.__dict__.items():
        if k in ("_blob", "_replay"):
            continue
        if isinstance(v, _SNAP_TYPES):
            snap[k] = v
        elif isinstance(v, np.ndarray):
            snap[k] = v.copy()
    return snap


def _stream_block(effects, x, y, silent):
    if silent and all(fx._at_rest() for fx in effects):
        # the reference's warm-up (engine.py:96-99: two zero blocks) and any silent stretch: zero in,
        # zero out; only write pointers, phasors, gate gain and the smoothed parameters move
        for fx in effects:
            fx._stages(x.shape[0])
            fx._rest_block(x.shape[0])
        y[...] = 0.0
        return
    per_fx = []
    for fx in effects:
        fx._require_usable()
        per_fx.append(fx._stream_desc(x.shape[0]))
    arr = _native.desc_array([d for ds in per_fx for d in ds])
    _native.stream_process(arr, x, y)
    k = 0
    for fx, ds in zip(effects, per_fx):
        fx._absorb(arr[k], x.shape[0], silent)
        k += len(ds)


def _replay_whole_clip(effects):
    """The lines a whole-clip call left behind are not written back by its kernels.  When another block
    follows without a prepare(), rebuild them: restore every effect's state from before that call and
    push the kept input through the streaming kernel (slow, exact, and only on this path)."""
    group = getattr(effects[0], "_replay", None)
    if group is None or group["ids"] != [id(fx) for fx in effects] or any(getattr(fx, "_replay", None) is not group for fx in effects):
        return False
    for fx, snap in zip(effects, group["snaps"]):
        for k, v in snap.items():
            setattr(fx, k, v.copy() if isinstance(v, np.ndarray) else v)
        fx._stale = False
        fx._blob_zero = True                    # fresh lines, as they were before the whole-clip call
        fx._replay = None
    x = group["x"]
    _stream_block(effects, x, np.empty((x.shape[0], 2), np.float32), not x.any())
    return True


def run_native(effects, sample_rate, x_in: np.ndarray, out: np.ndarray):
    """One block of (frames, channels) host arrays through a run of native effects: one fused
    whole-clip launch when the block is long and the lines are fresh, else the streaming kernel
    (also when the whole-clip kernels do not support the chain, e.g. comb lines shorter than a tile)."""
    frames = x_in.shape[0]
    x = np.ascontiguousarray(x_in, np.float32)
    if x.shape[1] not in (1, 2):
        raise ValueError("audioblocks (B200) processes mono or stereo blocks")
    if out.shape != (frames, 2):
        raise ValueError("output block must be (frames, 2)")
    y = out if (out.dtype == np.float32 and out.flags.c_contiguous) else np.empty((frames, 2), np.float32)
    silent = not x.any()
    if frames > 0 and any(fx._stale for fx in effects):
        _replay_whole_clip(effects)             # (if it cannot, _require_usable below says why)
    whole = frames >= WHOLE_CLIP_MIN_FRAMES and not any(fx._dirty or fx._stale for fx in effects)
    plan = None
    if whole:
        snaps = [_snapshot(fx) for fx in effects]
        descs = []
        for fx in effects:
            descs.extend(fx._stages(frames))
        try:
            plan = _native.ChainPlan(descs, sample_rate)
        except _native.AesimError as e:
            if e.code != _native.ERR_UNSUPPORTED:
                raise
            for fx, snap in zip(effects, snaps):    # e.g. an 8 kHz reverb: its shortest comb is 235 samples
                for k, v in snap.items():
                    setattr(fx, k, v)
            whole = False
    if frames == 0:
        pass
    elif whole:
        fmt_in = _native.FMT_F32_MONO if x.shape[1] == 1 else _native.FMT_F32_STEREO
        try:
            plan.run_host(x, fmt_in, y, _native.FMT_F32_STEREO, 1, frames)
            finals = [plan.final_state(s) for s in range(len(descs))]
        finally:
            plan.close()
        for fx, final in zip(effects, finals):          # one stage per effect
            fx._advance(frames, silent, final)
        if not silent and x.size <= REPLAY_MAX_FLOATS:
            group = {"x": x.copy() if x is x_in or np.shares_memory(x, x_in) else x, "snaps": snaps,
                     "ids": [id(fx) for fx in effects]}
            for fx in effects:
                fx._replay = group
    else:
        _stream_block(effects, x, y, silent)
    if y is not out:
        out[:, :] = y


class PlotDataTap(Effect):
    """Transparent tap that copies blocks into a queue for the UI (reference core.py:89-106)."""

    def __init__(self, data_queue: queue.Queue):
        self.queue = data_queue

    def process_into(self, x_in: np.ndarray, out: np.ndarray) -> None:
        out[:] = x_in
        try:
            self.queue.put_nowait(x_in.copy())
        except queue.Full:
            pass


class EffectsChain:
    """Ping-pong driver with the reference's protocol (core.py:109-161): effects are
    prepared on add(); a frame-count change re-prepares all of them; warmup() pushes
    two silent blocks.  Consecutive native effects run as ONE fused kernel launch."""

    def __init__(self, sample_rate: int, channels_in: int, channels_out: int, blocksize: int):
        self.sr, self.ci, self.co, self.bs = sample_rate, channels_in, channels_out, blocksize
        self.effects: list[Effect] = []

    def add(self, effect: Effect):
        effect.prepare(self.sr, self.ci, self.co, self.bs)
        self.effects.append(effect)

    def _ensure_blocksize(self, frames: int):
        if frames != self.bs:
            self.bs = frames
            for e in self.effects:
                e.prepare(self.sr, self.ci, self.co, frames)

    def warmup(self):
        zi = np.zeros((self.bs, self.ci), np.float32)
        zo = np.zeros((self.bs, self.co), np.float32)
        for _ in range(2):
            self.process(zi, zo)

    def _segments(self):
        seg = []
        for e in self.effects:
            if isinstance(e, NativeEffect):
                seg.append(e)
            else:
                if seg:
                    yield seg
                    seg = []
                yield e
        if seg:
            yield seg

    def process(self, in_block: np.ndarray, out_block: np.ndarray):
        """in_block (frames, ci) float32 -> out_block (frames, co) float32."""
        frames = in_block.shape[0]
        self._ensure_blocksize(frames)
        if self.co != 2:
            raise ValueError("audioblocks (B200) chains are stereo-out, like the reference engine")
        src = in_block
        fanned = False
        for seg in self._segments():
            if isinstance(seg, list):
                # the native path fans mono out to L=R itself (core.py:147-149)
                dst = np.empty((frames, self.co), np.float32)
                run_native(seg, self.sr, src if src.shape[1] in (1, 2) else src[:, :2], dst)
                fanned = True
            else:
                if not fanned:
                    src = self._fan(src, frames)
                    fanned = True
                dst = np.empty((frames, self.co), np.float32)
                seg.process_into(src, dst)
            src = dst
        if not fanned:
            src = self._fan(src, frames)
        out_block[:, :] = src

    def _fan(self, x, frames):
        buf = np.zeros((frames, self.co), np.float32)
        if self.ci == 1 and self.co == 2:
            buf[:, 0] = x[:, 0]
            buf[:, 1] = x[:, 0]
        else:
            k = min(self.ci, self.co)
            buf[:, :k] = x[:, :k]
        return buf

    # ---- batched entry (new; the reference is strictly one clip per call) -------------
    def stage_descs(self, frames: int):
        """Resolved aes_stage_desc list for the whole chain at its current state."""
        descs = []
        for e in self.effects:
            if isinstance(e, PlotDataTap):
                continue
            if not isinstance(e, NativeEffect):
                raise TypeError(f"{type(e).__name__} has no CUDA implementation; cannot batch")
            e._require_fresh()
            descs.extend(e._stages(frames))
        return descs

    def prepare_batch(self, frames: int):
        """Bring the chain to the state it has when a file of `frames` frames arrives
        (engine.py:99-102: after warm-up, re-prepared at the file's frame count) and
        return the compiled plan."""
        self._ensure_blocksize(frames)
        return _native.ChainPlan(self.stage_descs(frames), self.sr)

    def device_pipeline(self, frames: int):
        """Compile the chain for device-resident batches of (B, frames, 2) float32 clips.

        Returns `run(x_ptr, y_ptr, tmp_ptr, n_clips, stream=0)` and the list of plan objects it keeps
        alive.  x/y/tmp are CUDA device pointers of n_clips*frames*2 floats (tmp is only touched when
        the chain has more than one segment); runs of fusable effects are one kernel launch each,
        whole-clip FFT effects (SpectralFilter, ConvolutionReverbEffect) run between them."""
        from .convreverb import ConvolutionReverbEffect
        from .spectral import SpectralFilter
        self._ensure_blocksize(frames)
        steps, cur, plans = [], [], []

        def flush():
            if cur:
                descs = []
                for e in cur:
                    e._require_fresh()
                    descs.extend(e._stages(frames))
                plan = _native.ChainPlan(descs, self.sr)
                plans.append(plan)
                steps.append(lambda xp, yp, B, st, plan=plan: plan.run_device(
                    xp, _native.FMT_F32_STEREO, yp, _native.FMT_F32_STEREO, B, frames, st))
                cur.clear()

        for e in self.effects:
            if isinstance(e, PlotDataTap):
                continue
            if isinstance(e, NativeEffect):
                cur.append(e)
            elif isinstance(e, SpectralFilter):
                flush()
                thr, red = e._params()
                sp = e._plan(2 * frames)
                plans.append(sp)
                steps.append(lambda xp, yp, B, st, sp=sp, thr=thr, red=red, al=float(e.alpha_param):
                             sp.run_device(xp, yp, B, frames, thr, red, al, st))
            elif isinstance(e, ConvolutionReverbEffect):
                flush()
                cp = e.plan()
                plans.append(cp)
                steps.append(lambda xp, yp, B, st, cp=cp, e=e: cp.run_device(xp, yp, B, frames, e.mix_dry, e.mix_wet, st))
            else:
                raise TypeError(f"{type(e).__name__} has no CUDA implementation")
        flush()
        if not steps:
            plan = _native.ChainPlan([], self.sr)
            plans.append(plan)
            steps.append(lambda xp, yp, B, st, plan=plan: plan.run_device(
                xp, _native.FMT_F32_STEREO, yp, _native.FMT_F32_STEREO, B, frames, st))

        def run(x_ptr, y_ptr, tmp_ptr, n_clips, stream=0):
            src = x_ptr
            for k, step in enumerate(steps):
                # ping-pong so that the last segment lands in y
                dst = y_ptr if (len(steps) - 1 - k) % 2 == 0 else tmp_ptr
                step(src, dst, n_clips, stream)
                src = dst

        run.n_segments = len(steps)
        return run, plans

    @staticmethod
    def _stream_batch(seg, data, dst):
        """Batch through the streaming kernel, one clip at a time, every clip from the same freshly
        prepared state (chains the whole-clip kernels refuse); formats converted on the host like
        engine.py:78-84 / :104-105."""
        snaps = [_snapshot(fx) for fx in seg]
        for b in range(data.shape[0]):
            for fx, snap in zip(seg, snaps):
                for k, v in snap.items():
                    setattr(fx, k, v.copy() if isinstance(v, np.ndarray) else v)
                fx._dirty = fx._stale = False
                fx._blob_zero = True
            xb = data[b]
            if xb.dtype == np.int16:                 # int16 -> /32768 -> mean over channels, exact in f32
                xb = ((xb[:, :1].astype(np.int32) + xb[:, 1:2].astype(np.int32)).astype(np.float32) * np.float32(1.0 / 65536.0))
            xb = np.ascontiguousarray(xb, np.float32)
            yb = dst[b] if dst.dtype == np.float32 else np.empty((xb.shape[0], 2), np.float32)
            _stream_block(seg, xb, yb, not xb.any())
            if dst.dtype == np.int16:
                dst[b] = (np.clip(yb, -1.0, 1.0) * np.float32(32767.0)).astype(np.int16)
        for fx, snap in zip(seg, snaps):             # a batch leaves the chain as it found it
            for k, v in snap.items():
                setattr(fx, k, v)
            fx._dirty = fx._stale = False
            fx._blob_zero = True

    def process_batch(self, x: np.ndarray, out: np.ndarray | None = None) -> np.ndarray:
        """x: (B, frames, 1|2) float32 host array (or (B, frames, 2) int16 PCM, which is
        down-mixed like engine.py:78-84) -> (B, frames, 2) float32 or int16 (if `out`
        is int16: clip, *32767, truncate, engine.py:104-105).  Every clip starts from
        the same freshly prepared state.  Runs of fusable effects are one kernel launch each;
        whole-clip FFT effects (ConvolutionReverbEffect) run between them."""
        from .convreverb import ConvolutionReverbEffect
        from .spectral import SpectralFilter
        B, frames, ch = x.shape
        self._ensure_blocksize(frames)
        segs, cur = [], []
        for e in self.effects:
            if isinstance(e, PlotDataTap):
                continue
            if isinstance(e, NativeEffect):
                cur.append(e)
            elif isinstance(e, (ConvolutionReverbEffect, SpectralFilter)):
                if cur:
                    segs.append(cur)
                    cur = []
                segs.append(e)
            else:
                raise TypeError(f"{type(e).__name__} has no CUDA implementation; cannot batch")
        if cur or not segs:
            segs.append(cur)
        want_i16 = out is not None and out.dtype == np.int16
        data = x if x.dtype == np.int16 else np.ascontiguousarray(x, np.float32)
        for k, seg in enumerate(segs):
            last = k == len(segs) - 1
            if isinstance(seg, list):
                descs = []
                for e in seg:
                    e._require_fresh()
                    descs.extend(e._stages(frames))
                if data.dtype == np.int16:
                    fmt_in = _native.FMT_I16_DOWNMIX
                else:
                    fmt_in = _native.FMT_F32_MONO if data.shape[2] == 1 else _native.FMT_F32_STEREO
                if last and out is not None:
                    dst = out
                else:
                    dst = np.empty((B, frames, 2), np.int16 if (last and want_i16) else np.float32)
                fmt_out = _native.FMT_I16_STEREO if dst.dtype == np.int16 else _native.FMT_F32_STEREO
                try:
                    plan = _native.ChainPlan(descs, self.sr)
                except _native.AesimError as err:
                    if err.code != _native.ERR_UNSUPPORTED:
                        raise
                    plan = None                      # e.g. comb lines shorter than a tile (8 kHz): stream clip by clip
                if plan is None:
                    self._stream_batch(seg, data, dst)
                else:
                    try:
                        plan.run_host(np.ascontiguousarray(data), fmt_in, dst, fmt_out, B, frames)
                    finally:
                        plan.close()
                data = dst
            else:
                if data.dtype == np.int16 or data.shape[2] != 2:      # down-mix / fan out through an empty chain
                    data = EffectsChain(self.sr, self.ci, self.co, frames).process_batch(data)
                dst = out if (last and out is not None and out.dtype == np.float32) else None
                data = seg.process_batch(data, dst)
        if out is None:
            return data
        if data is not out:
            if want_i16 and data.dtype != np.int16:
                tmp = EffectsChain(self.sr, 2, 2, frames)
                tmp.process_batch(data, out)
            else:
                out[...] = data
        return out
