#!/usr/bin/env python
"""Wall-clock breakdown of one `process_file` request (reference engine.py:74-126) through this
package on cuda:0: WAV decode, chain, WAV encode + base64, reply serialisation -- the last one
both ways: the reference's `json.dumps` over `.tolist()` and libaesim's writer.

usage: python profiles/tools/file_route_timing.py [frames] [sample_rate]   (default: the shipped
clip's size, 892 775 stereo int16 frames at 48 kHz, synthetic content)
"""
import base64
import io
import json
import os
import sys
import time

import numpy as np
import scipy.io.wavfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "audio-effects-simulator_b200"), os.path.join(ROOT, "tests")]
import audioblocks as ab                                  # noqa: E402
from audioblocks import engine as eng                     # noqa: E402
import synth                                              # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 892_775
    fs = int(sys.argv[2]) if len(sys.argv) > 2 else 48_000
    x = synth.clip(77, n, 2, fs)
    pcm_in = (np.clip(x, -1, 1) * 32767).astype(np.int16)
    with io.BytesIO() as fh:
        scipy.io.wavfile.write(fh, fs, pcm_in)
        contents = "data:audio/wav;base64," + base64.b64encode(fh.getvalue()).decode("ascii")
    import queue
    e = ab.AudioEngine({"input": queue.Queue(maxsize=10), "output": queue.Queue(maxsize=10)})
    e.build_chain(synth.PRESETS["Rain Delay"])
    out = {}
    for rep in range(3):                                   # first pass warms the plan cache / CUDA context
        t = [time.perf_counter()]
        _, payload = contents.split(",")
        audio, fs_in = eng.read_wav(base64.b64decode(payload)); t.append(time.perf_counter())
        mono, processed, pcm = e.process_file_arrays(audio, fs_in); t.append(time.perf_counter())
        with io.BytesIO() as out_io:
            scipy.io.wavfile.write(out_io, fs_in, pcm)
            url = "data:audio/wav;base64," + base64.b64encode(out_io.getvalue()).decode("ascii")
        t.append(time.perf_counter())
        msg = eng.file_processed_message(contents, url, fs_in, mono, processed); t.append(time.perf_counter())
        ref = json.dumps({"type": "file_processed", "original_b64": contents, "processed_b64": url,
                          "sample_rate": fs_in, "original_samples": mono.flatten().tolist(),
                          "processed_samples": processed.mean(axis=1).flatten().tolist()})
        t.append(time.perf_counter())
        d = np.diff(t) * 1e3
        out = {"frames": n, "sample_rate": fs, "host_threads": os.cpu_count(), "reply_bytes": len(msg),
               "identical_reply": msg == ref,
               "ms": {"wav_decode": d[0], "chain_incl_h2d_d2h": d[1], "wav_encode_base64": d[2],
                      "reply_native_writer": d[3], "reply_json_dumps_tolist": d[4]},
               "request_ms_native": float(d[:4].sum()), "request_ms_reference_style": float(d[:3].sum() + d[4])}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
