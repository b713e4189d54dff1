"""File-route reply serialisation: libaesim's float-list writer must produce the
exact text the reference produces with `json.dumps(arr.tolist())` (engine.py:115-122).
Host-only code in the C-ABI library -- runs without a GPU."""
import json

import numpy as np
import pytest

from audioblocks import _native
from audioblocks.engine import file_processed_message


def want(x):
    return json.dumps(np.asarray(x, dtype=np.float32).flatten().tolist())


def test_special_values_and_notation_switches():
    x = np.array([0.0, -0.0, 1.0, -1.0, 0.1, 1e-5, 1e-4, 9.999e-5, 1e16, 1e15, 123456789.0, 3.4028235e38,
                  1.4e-45, 1.17549435e-38, np.inf, -np.inf, np.nan, 0.5, 16777216.0, 1e22, 1e21, 2.5e-7,
                  -3.0517578e-05, 32767.0 / 32768.0], dtype=np.float32)
    assert _native.json_float_list(x) == want(x)


def test_random_bit_patterns_match_python_repr():
    rng = np.random.default_rng(11)
    x = rng.integers(0, 2 ** 32, 300_000, dtype=np.uint64).astype(np.uint32).view(np.float32)
    assert _native.json_float_list(x) == want(x)


@pytest.mark.parametrize("n", [0, 1, 2, 4095, 4096, 4097, 20_001])
@pytest.mark.parametrize("threads", [0, 1, 3])
def test_lengths_and_thread_counts(n, threads):
    x = np.random.default_rng(n + 1).standard_normal(n).astype(np.float32)
    assert _native.json_float_list(x, threads=threads) == want(x)


def test_stereo_mean_is_numpys_float32_mean():
    rng = np.random.default_rng(5)
    y = (rng.standard_normal((50_000, 2)) * 0.3).astype(np.float32)
    y[:4] = [[1.0, 1e-8], [3.4e38, 3.4e38], [-0.0, 0.0], [1e-45, 1e-45]]
    with np.errstate(over="ignore"):
        ref = json.dumps(y.mean(axis=1).flatten().tolist())
    assert _native.json_float_list(y, stereo_mean=True) == ref


def test_whole_reply_is_json_dumps_of_the_reference_dict():
    rng = np.random.default_rng(9)
    mono = (rng.standard_normal((3000, 1)) * 0.2).astype(np.float32)
    processed = (rng.standard_normal((3000, 2)) * 0.2).astype(np.float32)
    contents = 'data:audio/wav;base64,AAAA"\\/+=='
    url = "data:audio/wav;base64,QUJD"
    ref = json.dumps({
        "type": "file_processed",
        "original_b64": contents,
        "processed_b64": url,
        "sample_rate": 44100,
        "original_samples": mono.flatten().tolist(),
        "processed_samples": processed.mean(axis=1).flatten().tolist(),
    })
    got = file_processed_message(contents, url, 44100, mono, processed)
    assert got == ref
    assert json.loads(got)["sample_rate"] == 44100


def test_buffer_too_small_is_refused():
    import ctypes as C
    x = np.ones(10, np.float32)
    buf = np.empty(16, np.uint8)
    rc = _native.lib().aes_json_float_list(C.c_void_p(x.ctypes.data), 10, C.c_void_p(buf.ctypes.data), 16, 1)
    assert rc < 0
