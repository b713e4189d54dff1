"""ctypes binding of libaesim.so (the C ABI declared in include/aesim.h).

There is no CPU fallback: if the CUDA library is missing or no device is present,
the first call that needs it raises.  Importing the package stays possible on a
GPU-less machine so that parameter resolution (pure host logic) can be tested.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("AESIM_LIB", os.path.join(os.path.dirname(_PKG), "lib", "libaesim.so"))

# aes_stage_kind / aes_format (include/aesim.h)
DELAY, REVERB, BIQUAD, GATE, OCTAVER, DISTORTION = 1, 2, 3, 4, 5, 6
FMT_F32_STEREO, FMT_F32_MONO, FMT_I16_DOWNMIX, FMT_I16_STEREO = 0, 1, 2, 3


class StageDesc(C.Structure):
    """aes_stage_desc"""
    _fields_ = [("kind", C.c_int32), ("flags", C.c_int32), ("p", C.c_double * 32), ("q", C.c_int64 * 32)]

    def key(self):
        return (self.kind, tuple(self.p), tuple(self.q))


class AesimError(RuntimeError):
    code = 0


ERR_UNSUPPORTED = -4        # AES_ERR_UNSUPPORTED: valid in the reference, not on the whole-clip kernels


_lib = None
_lock = threading.Lock()


def lib() -> C.CDLL:
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise AesimError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  audioblocks has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        vp, i64, ci = C.c_void_p, C.c_int64, C.c_int
        L.aes_last_error.restype = C.c_char_p
        L.aes_launch_count.restype = i64
        L.aes_chain_plan_create.argtypes = [C.POINTER(StageDesc), ci, ci, C.POINTER(vp)]
        L.aes_chain_plan_destroy.argtypes = [vp]
        L.aes_chain_run.argtypes = [vp, vp, ci, vp, ci, i64, i64, vp]
        L.aes_chain_process_host.argtypes = [vp, vp, ci, vp, ci, i64, i64]
        L.aes_chain_final_state.argtypes = [vp, ci, C.POINTER(C.c_double)]
        L.aes_chain_plan_kernel_name.restype = C.c_char_p
        L.aes_chain_plan_kernel_name.argtypes = [vp]
        L.aes_chain_plan_info.argtypes = [vp, C.POINTER(ci), C.POINTER(ci), C.POINTER(ci), C.POINTER(i64)]
        L.aes_malloc.argtypes = [C.POINTER(vp), C.c_size_t]
        L.aes_free.argtypes = [vp]
        L.aes_host_alloc.argtypes = [C.POINTER(vp), C.c_size_t]
        L.aes_host_free.argtypes = [vp]
        L.aes_memcpy_h2d.argtypes = [vp, vp, C.c_size_t, vp]
        L.aes_memcpy_d2h.argtypes = [vp, vp, C.c_size_t, vp]
        L.aes_memset.argtypes = [vp, ci, C.c_size_t, vp]
        L.aes_stream_create.argtypes = [C.POINTER(vp)]
        L.aes_stream_destroy.argtypes = [vp]
        L.aes_stream_sync.argtypes = [vp]
        L.aes_device_count.argtypes = [C.POINTER(ci)]
        L.aes_set_device.argtypes = [ci]
        L.aes_device_sm_count.argtypes = [C.POINTER(ci)]
        L.aes_delay_f32.argtypes = [vp, vp, i64, i64, i64, i64, C.c_double, C.c_double, C.c_double, vp]
        L.aes_biquad_cascade_f32.argtypes = [vp, vp, i64, i64, ci, C.POINTER(C.c_double), vp]
        L.aes_quantize_i16.argtypes = [vp, vp, i64, vp]
        L.aes_convreverb_plan_create.argtypes = [vp, i64, ci, C.POINTER(vp)]
        L.aes_convreverb_plan_destroy.argtypes = [vp]
        L.aes_convreverb_run.argtypes = [vp, vp, vp, i64, i64, C.c_double, C.c_double, vp]
        L.aes_convreverb_process_host.argtypes = [vp, vp, vp, i64, i64, C.c_double, C.c_double]
        L.aes_convreverb_plan_info.argtypes = [vp, C.POINTER(ci), C.POINTER(ci)]
        L.aes_spectral_plan_create.argtypes = [i64, C.POINTER(vp)]
        L.aes_spectral_plan_destroy.argtypes = [vp]
        L.aes_spectral_frames_host.argtypes = [vp, vp, vp, vp, ci, C.c_double, C.c_double, C.c_double]
        L.aes_spectral_run.argtypes = [vp, vp, vp, i64, i64, C.c_double, C.c_double, C.c_double, vp]
        L.aes_spectral_process_host.argtypes = [vp, vp, vp, i64, i64, C.c_double, C.c_double, C.c_double]
        L.aes_stream_process_host.argtypes = [C.POINTER(StageDesc), ci, vp, ci, vp, i64]
        L.aes_spectrum_chroma.argtypes = [vp, vp, i64, i64, ci, C.c_double, vp, vp, vp, vp, vp]
        L.aes_spectrum_chroma_host.argtypes = [vp, vp, i64, i64, ci, C.c_double, vp, vp, vp, vp]
        for fn in (L.aes_json_float_list, L.aes_json_stereo_mean_list):
            fn.restype = i64
            fn.argtypes = [vp, i64, vp, i64, ci]
        L.aes_json_float_list_bound.restype = i64
        L.aes_json_float_list_bound.argtypes = [i64]
        if L.aes_abi_version() != 1:
            raise AesimError("libaesim.so ABI version mismatch")
        _lib = L
        return L


def check(rc: int):
    if rc != 0:
        err = AesimError(f"aesim error {rc}: {lib().aes_last_error().decode(errors='replace')}")
        err.code = rc
        raise err


def desc_array(descs):
    arr = (StageDesc * max(1, len(descs)))()
    for i, d in enumerate(descs):
        arr[i] = d
    return arr


class ChainPlan:
    """RAII wrapper of aes_chain_plan."""

    def __init__(self, descs, sample_rate: int):
        self._h = C.c_void_p()
        self._descs = desc_array(descs)
        check(lib().aes_chain_plan_create(self._descs, len(descs), int(sample_rate), C.byref(self._h)))

    def info(self):
        t, s, o, sc = C.c_int(), C.c_int(), C.c_int(), C.c_int64()
        check(lib().aes_chain_plan_info(self._h, C.byref(t), C.byref(s), C.byref(o), C.byref(sc)))
        return {"tile_frames": t.value, "smem_bytes": s.value, "ctas_per_sm": o.value,
                "scratch_bytes_per_cta": sc.value,
                "kernel": lib().aes_chain_plan_kernel_name(self._h).decode()}

    def run_device(self, x_ptr: int, in_fmt: int, y_ptr: int, out_fmt: int, n_clips: int, n_frames: int,
                   stream: int = 0):
        """x_ptr / y_ptr: CUDA device pointers (e.g. torch tensor .data_ptr())."""
        check(lib().aes_chain_run(self._h, C.c_void_p(x_ptr), in_fmt, C.c_void_p(y_ptr), out_fmt,
                                  n_clips, n_frames, C.c_void_p(stream)))

    def run_host(self, x: np.ndarray, in_fmt: int, y: np.ndarray, out_fmt: int, n_clips: int, n_frames: int):
        assert x.flags.c_contiguous and y.flags.c_contiguous
        check(lib().aes_chain_process_host(self._h, C.c_void_p(x.ctypes.data), in_fmt,
                                           C.c_void_p(y.ctypes.data), out_fmt, n_clips, n_frames))

    def final_state(self, stage: int):
        """Carried scalars (16 doubles) of `stage` after a single-clip run_host."""
        out = (C.c_double * 16)()
        check(lib().aes_chain_final_state(self._h, stage, out))
        return list(out)

    def close(self):
        if self._h:
            lib().aes_chain_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DeviceBlob:
    """Zero-initialised device memory holding an effect's delay lines for the streaming path."""

    def __init__(self, n_floats: int):
        self.n_floats = int(n_floats)
        self._p = C.c_void_p()
        check(lib().aes_malloc(C.byref(self._p), max(4, 4 * self.n_floats)))
        self.zero()

    @property
    def ptr(self) -> int:
        return self._p.value

    def zero(self):
        check(lib().aes_memset(self._p, 0, max(4, 4 * self.n_floats), None))
        check(lib().aes_stream_sync(None))

    def close(self):
        if self._p:
            lib().aes_free(self._p)
            self._p = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def stream_process(descs, x: np.ndarray, y: np.ndarray):
    """aes_stream_process_host; `descs` (ctypes array) is updated in place."""
    assert x.dtype == np.float32 and y.dtype == np.float32 and x.flags.c_contiguous and y.flags.c_contiguous
    check(lib().aes_stream_process_host(descs, len(descs), C.c_void_p(x.ctypes.data), x.shape[1],
                                        C.c_void_p(y.ctypes.data), x.shape[0]))


class ConvReverbPlan:
    """RAII wrapper of aes_convreverb_plan (IR partition spectra live on the device)."""

    def __init__(self, ir: np.ndarray, block_log2: int = 0):
        ir = np.ascontiguousarray(ir, np.float32)
        if ir.ndim != 2 or ir.shape[1] != 2:
            raise ValueError("impulse response must be (n_taps, 2)")
        self._h = C.c_void_p()
        check(lib().aes_convreverb_plan_create(C.c_void_p(ir.ctypes.data), ir.shape[0], block_log2, C.byref(self._h)))

    def info(self):
        n, p = C.c_int(), C.c_int()
        check(lib().aes_convreverb_plan_info(self._h, C.byref(n), C.byref(p)))
        return {"fft_size": n.value, "partitions": p.value}

    def run_device(self, x_ptr, y_ptr, n_clips, n_frames, dry, wet, stream=0):
        check(lib().aes_convreverb_run(self._h, C.c_void_p(x_ptr), C.c_void_p(y_ptr), n_clips, n_frames,
                                       float(dry), float(wet), C.c_void_p(stream)))

    def run_host(self, x: np.ndarray, y: np.ndarray, n_clips, n_frames, dry, wet):
        assert x.flags.c_contiguous and y.flags.c_contiguous and x.dtype == np.float32 and y.dtype == np.float32
        check(lib().aes_convreverb_process_host(self._h, C.c_void_p(x.ctypes.data), C.c_void_p(y.ctypes.data),
                                                n_clips, n_frames, float(dry), float(wet)))

    def close(self):
        if self._h:
            lib().aes_convreverb_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class SpectralPlan:
    """RAII wrapper of aes_spectral_plan for one frame length M = 2 * hop."""

    def __init__(self, frame_len: int):
        self.frame_len = int(frame_len)
        self._h = C.c_void_p()
        check(lib().aes_spectral_plan_create(self.frame_len, C.byref(self._h)))

    def frames_host(self, in_buffers: np.ndarray, mask: np.ndarray, thr: float, red: float, alpha: float) -> np.ndarray:
        """in_buffers (nb, M) f32 raw analysis buffers, mask (nb, M/2+1) f32 updated in place ->
        (nb, M) f32 irfft of the gated spectrum."""
        assert in_buffers.dtype == np.float32 and mask.dtype == np.float32
        assert in_buffers.flags.c_contiguous and mask.flags.c_contiguous
        y = np.empty_like(in_buffers)
        check(lib().aes_spectral_frames_host(self._h, C.c_void_p(in_buffers.ctypes.data), C.c_void_p(mask.ctypes.data),
                                             C.c_void_p(y.ctypes.data), in_buffers.shape[0], thr, red, alpha))
        return y

    def process_host(self, x: np.ndarray, y: np.ndarray, thr: float, red: float, alpha: float):
        """Whole clips from the freshly re-initialised state: x, y (B, N, 2) f32 host arrays, M == 2N."""
        check(lib().aes_spectral_process_host(self._h, C.c_void_p(x.ctypes.data), C.c_void_p(y.ctypes.data),
                                              x.shape[0], x.shape[1], thr, red, alpha))

    def run_device(self, x_ptr, y_ptr, n_clips, n_frames, thr, red, alpha, stream=0):
        check(lib().aes_spectral_run(self._h, C.c_void_p(x_ptr), C.c_void_p(y_ptr), n_clips, n_frames,
                                     thr, red, alpha, C.c_void_p(stream)))

    def close(self):
        if self._h:
            lib().aes_spectral_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def pinned_empty(shape, dtype=np.float32) -> np.ndarray:
    """A numpy array over page-locked host memory, freed when the array and every view of it are gone
    (the owner rides on the ctypes buffer numpy keeps as the array's base)."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    p = C.c_void_p()
    check(lib().aes_host_alloc(C.byref(p), max(n, 1)))
    buf = (C.c_char * max(n, 1)).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    class _Owner:
        def __init__(self, ptr): self.ptr = ptr
        def __del__(self):
            try: lib().aes_host_free(self.ptr)
            except Exception: pass
    buf._aes_owner = _Owner(p)
    return arr


def json_float_list(x: np.ndarray, stereo_mean: bool = False, threads: int = 0) -> str:
    """The text `json.dumps(x.flatten().tolist())` would produce for a float32 array, or, with
    `stereo_mean`, `json.dumps(x.mean(axis=1).flatten().tolist())` for a (frames, 2) float32 array
    (reference engine.py:119-120), written by libaesim's parallel formatter."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    if stereo_mean:
        if x.ndim != 2 or x.shape[1] != 2:
            raise AesimError("json_float_list(stereo_mean=True) needs a (frames, 2) array")
        n, fn = x.shape[0], lib().aes_json_stereo_mean_list
    else:
        n, fn = x.size, lib().aes_json_float_list
    cap = lib().aes_json_float_list_bound(n)
    buf = np.empty(cap, dtype=np.uint8)
    got = fn(C.c_void_p(x.ctypes.data), n, C.c_void_p(buf.ctypes.data), cap, int(threads))
    if got < 0:
        check(int(got))
    return str(memoryview(buf)[:got], "ascii")


def release_host_cache():
    """Free the calling thread's cached pipeline buffers (aes_release_host_cache): streams, device
    and pinned staging, delay-line scratch kept between host-buffer calls."""
    check(lib().aes_release_host_cache())
