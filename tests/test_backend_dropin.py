"""Drop-in test of the transport: the reference's own src/backend.py (unmodified, read from
/root/reference when that mount exists, i.e. in the build container) is started against THIS
repository's `audioblocks` package and spoken to over its WebSocket protocol.
An empty chain needs no CUDA device (no effect => no kernel), so the server side of the boundary
-- imports, AudioEngine construction, build_chain, process_file, the reply JSON -- is covered on
the CPU; chains with effects are covered by tests/test_gpu_parity.py."""
import asyncio
import base64
import io
import json
import os
import socket
import subprocess
import sys
import time

import numpy as np
import pytest

REF_BACKEND = "/root/reference/src/backend.py"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "audio-effects-simulator_b200")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.skipif(not os.path.exists(REF_BACKEND), reason="reference checkout not mounted")
def test_reference_backend_runs_on_our_audioblocks():
    websockets = pytest.importorskip("websockets")
    import scipy.io.wavfile
    port = _free_port()
    boot = ("import sys, runpy; sys.path.insert(0, %r); import audioblocks; "
            "assert audioblocks.__file__.startswith(%r), audioblocks.__file__; "
            "runpy.run_path(%r, run_name='__main__')") % (PKG, PKG, REF_BACKEND)
    env = dict(os.environ, PORT=str(port))
    proc = subprocess.Popen([sys.executable, "-c", boot], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    try:
        fs, n = 48000, 2000
        pcm = (np.random.default_rng(0).uniform(-0.5, 0.5, (n, 2)) * 32767).astype(np.int16)
        buf = io.BytesIO()
        scipy.io.wavfile.write(buf, fs, pcm)
        url = "data:audio/wav;base64," + base64.b64encode(buf.getvalue()).decode("ascii")

        async def talk():
            for _ in range(100):
                try:
                    ws = await websockets.connect(f"ws://127.0.0.1:{port}", max_size=None)
                    break
                except OSError:
                    await asyncio.sleep(0.2)
            else:
                raise RuntimeError("backend did not come up")
            async with ws:
                await ws.send(json.dumps({"command": "build_chain", "config": []}))
                await ws.send(json.dumps({"command": "process_file", "contents": url, "filename": "t.wav"}))
                while True:
                    msg = json.loads(await asyncio.wait_for(ws.recv(), 60))
                    if msg.get("type") == "file_processed":
                        return msg

        reply = asyncio.run(talk())
        assert reply["sample_rate"] == fs and len(reply["processed_samples"]) == n
        rfs, out = scipy.io.wavfile.read(io.BytesIO(base64.b64decode(reply["processed_b64"].split(",")[1])))
        # empty chain: mono mean fanned out to L=R, clipped, *32767, truncated (engine.py:81-105)
        mono = (pcm.astype(np.float32) / np.float32(32768.0)).mean(axis=1, keepdims=True)
        want = (np.clip(np.repeat(mono, 2, axis=1), -1.0, 1.0) * 32767).astype(np.int16)
        assert rfs == fs and np.array_equal(out, want)
    finally:
        proc.terminate()
        try:
            proc.wait(timeout=5)
        except Exception:
            proc.kill()
