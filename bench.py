#!/usr/bin/env python
"""bench.py -- throughput of the offline effect-chain hot path.

    python bench.py --gpus N --steps K --warmup W            # B200 arm (this repo)
    python bench.py --impl reference --gpus N --steps K ...  # reference arm: the CPU algorithm on host cores

Default workload = BASELINE configs[4]: 8192 synthetic 10 s 48 kHz stereo clips, batch-sharded
over the N ranks (STRONG scaling: 8192 / N clips per GPU).  A "step" is one pass of the fused
'Rain Delay' preset chain over the rank's shard, already resident in HBM (`value`, `roofline`).
The same line carries
  * `pipelined_steps`: the same passes with consecutive passes on two streams / two output buffers
    (the partial last wave of a pass filled by the next pass; beside `value`, never instead of it);
  * `sweep`: every preset of app.py:41-71 over the same shard, one at a time and all six at once
    on separate streams (per-preset and aggregate Msamples/s), plus two chains that have no kernel
    of their own;
  * `e2e`: a bounded sample of the shard pushed through the reference-facing host-buffer call with
    H2D/D2H inside the timed region, next to a copy-only ceiling of the same bytes;
  * `gather` (N > 1): the int16 results of a bounded sample gathered over NCCL, alone and
    overlapped with the next chunk's compute;
  * `cpu_baseline` (N = 1): the unmodified reference (numba) on the box's host cores, and the C port.
1 sample = one float32 value of one channel of output; algorithmic traffic is 8 bytes
per sample (read once, write once).  No data-path collective: clips are independent.
`--clips K` switches to K clips per GPU (weak scaling) for side experiments.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "audio-effects-simulator_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

FS = 48000
METRIC = "Msamples/s through preset chain"
UNIT = "Msamples/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--preset", default="Rain Delay")
    ap.add_argument("--total-clips", type=int, default=8192, help="clips in the whole job, sharded over the ranks (BASELINE configs[4])")
    ap.add_argument("--clips", type=int, default=0, help="clips PER GPU instead (weak scaling; side experiments)")
    ap.add_argument("--seconds", type=float, default=10.0, help="clip length")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--e2e-clips", type=int, default=1184, help="bounded sample of the shard for the end-to-end / copy-ceiling legs")
    ap.add_argument("--no-sweep", action="store_true")
    ap.add_argument("--sweep-steps", type=int, default=2)
    ap.add_argument("--no-gather", action="store_true")
    ap.add_argument("--gather-clips", type=int, default=512, help="clips per rank in the NCCL gather leg")
    ap.add_argument("--cpu-clips", type=int, default=0, help="clips in the CPU baseline sample (0 = auto)")
    return ap.parse_args()


EXTRA_CHAINS = {
    # BASELINE configs[1]: LP / HP / BP / peaking cascade
    "c2-biquad-cascade": [
        {"type": "filter", "params": {"filter_type": 0, "cutoff_hz": 8000, "q": 0.707}},
        {"type": "filter", "params": {"filter_type": 1, "cutoff_hz": 80, "q": 0.707}},
        {"type": "filter", "params": {"filter_type": 2, "cutoff_hz": 1000, "q": 0.8}},
        {"type": "filter", "params": {"filter_type": 3, "cutoff_hz": 1000, "q": 1.0, "gain_db": 6.0}}],
    # BASELINE configs[2]: distortion + octaver + feedback delay
    "c3-dist-octaver-delay": [
        {"type": "distortion", "params": {"drive": 4.0}},
        {"type": "octaver", "params": {"semitones": -12, "mix": 0.5}},
        {"type": "delay", "params": {"delay_ms": 120, "feedback": 0.3, "offset_ms": 10}}],
    # single blocks with their constructor defaults: per-block cost
    "block-delay": [{"type": "delay", "params": {}}],
    "block-reverb": [{"type": "reverb", "params": {}}],
    "block-filter": [{"type": "filter", "params": {}}],
    "block-gate": [{"type": "gate", "params": {}}],
    "block-octaver": [{"type": "octaver", "params": {}}],
    "block-distortion": [{"type": "distortion", "params": {"drive": 4.0}}],
}


def conv_arm(args, torch, dist, _native, rank, world, local, dev):
    """BASELINE configs[3]: IR-convolution reverb, 3 s synthetic IR, synthetic 30 s stereo clips."""
    import numpy as np
    from oracle import oracle as orc
    n_frames = int(args.seconds * FS)
    B = args.clips
    ir = orc.synthetic_ir(int(3.0 * FS))
    plan = _native.ConvReverbPlan(ir)
    x = synth_device(torch, B, n_frames, rank * B, dev)
    y = torch.empty_like(x)
    sptr = torch.cuda.current_stream().cuda_stream
    L = _native.lib()
    for _ in range(max(3, args.warmup)):
        plan.run_device(x.data_ptr(), y.data_ptr(), B, n_frames, 0.7, 0.5, sptr)
    torch.cuda.synchronize()
    l0 = L.aes_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        plan.run_device(x.data_ptr(), y.data_ptr(), B, n_frames, 0.7, 0.5, sptr)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    n_chk = min(n_frames, 200000)
    fx = orc.OConvReverb(ir, 0.7, 0.5)
    xs = x[0, :n_chk].cpu().numpy()
    want = np.zeros_like(xs)
    fx.process_into(xs, want)
    import synth
    mx, snr = synth.err_stats(y[0, :n_chk].cpu().numpy(), want)
    info = plan.info()
    samples = B * n_frames * 2
    nblk = -(-n_frames // (info["fft_size"] // 2))
    mac_flops = B * nblk * info["fft_size"] * info["partitions"] * 8.0       # one complex multiply-add per bin, partition and block
    peak = 6547.8
    pp = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pp):
        peak = float(json.load(open(pp))["hbm_gbs"])
    ach = samples * 8 / (ms * 1e-3) / 1e9
    print(json.dumps({
        "metric": METRIC, "value": samples / (ms * 1e-3) / 1e6, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"IR-convolution reverb, 3 s synthetic IR ({ir.shape[0]} taps), {B} synthetic "
                               f"{args.seconds:g} s 48 kHz stereo clips (BASELINE configs[3])", **info},
        "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                     "kernel": "aesc_fwd + aesc_mac (TMA-staged IR tiles) + aesc_inv", "mac_fp32_tflops": mac_flops / (ms * 1e-3) / 1e12},
        "gpu_launches": int(L.aes_launch_count() - l0), "parity": {"max_abs_err": mx, "snr_db": snr, "frames": n_chk},
    }))
    plan.close()


def chain_config(name):
    import synth
    if name in synth.PRESETS:
        return synth.PRESETS[name]
    return EXTRA_CHAINS[name]


def shard_of(args, rank, world):
    """(clips of this rank, clips of the whole job, scaling mode)."""
    if args.clips > 0:
        return args.clips, args.clips * world, "weak"
    from audioblocks.sharding import shard_range
    lo, hi = shard_range(args.total_clips, rank, world)
    return hi - lo, args.total_clips, "strong"


def workload(args, world=1):
    n_frames = int(args.seconds * FS)
    if args.clips > 0:
        what = f"{args.clips} synthetic {args.seconds:g} s 48 kHz stereo float32 clips per GPU (weak scaling; side experiment)"
        return {"workload": f"'{args.preset}' preset chain (app.py:41-71) on {what}", "preset": args.preset,
                "clips_per_gpu": args.clips, "frames_per_clip": n_frames, "sample_rate": FS,
                "l2": "inputs larger than L2 (no flush needed)", "parallelism": "clip-sharded"}
    return {"workload": f"BASELINE configs[4]: {args.total_clips} synthetic {args.seconds:g} s 48 kHz stereo float32 clips, "
                        f"batch-sharded over the GPUs (strong scaling); headline = '{args.preset}' preset chain "
                        f"(app.py:41-71), `sweep` = every preset over the same shard",
            "preset": args.preset, "total_clips": args.total_clips, "clips_per_gpu": args.total_clips // max(1, world),
            "frames_per_clip": n_frames, "sample_rate": FS, "l2": "inputs larger than L2 (no flush needed)",
            "parallelism": f"clip-sharded x{world}"}


# ------------------------------------------------------------------ reference arm / CPU baseline
def cpu_run(preset, n_clips, n_frames, threads, reps=1):
    """Time the CPU port of the reference algorithm (oracle, fast build) on `n_clips` clips."""
    import numpy as np
    import synth
    from oracle import oracle as orc
    x = np.stack([synth.clip(i % 16, n_frames, 2) for i in range(min(n_clips, 16))])
    if n_clips > x.shape[0]:
        x = np.concatenate([x] * ((n_clips + x.shape[0] - 1) // x.shape[0]))[:n_clips]
    x = np.ascontiguousarray(x)
    cfg = chain_config(preset)
    if any(c["type"] in ("spectral", "convreverb") for c in cfg):
        # no C driver for the numpy-FFT blocks: time the oracle's block wrappers, one clip at a time
        n = min(2, x.shape[0])
        t0 = time.perf_counter()
        for b in range(n):
            orc.run_file_path(cfg, x[b], FS)
        dt = time.perf_counter() - t0
        return n * x.shape[1] * 2 / dt / 1e6, dt
    orc.run_batch_c(cfg, x[:threads], FS, threads=threads, fast=True)      # page in / warm caches
    best = float("inf")
    for _ in range(reps):
        t0 = time.perf_counter()
        orc.run_batch_c(cfg, x, FS, threads=threads, fast=True)
        best = min(best, time.perf_counter() - t0)
    return x.size / best / 1e6, best


def reference_pool(preset, n_frames, cores):
    """The unmodified reference package on `cores` worker processes (baseline/ref_runner.py), or None when
    it is not installed under baseline/_ref or the chain uses a block the reference does not have."""
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    import ref_runner
    cfg = chain_config(preset)
    if not ref_runner.available() or not ref_runner.supports(cfg):
        return None
    try:
        return ref_runner.Pool(cfg, FS, n_frames, cores)
    except Exception as e:                                  # numba missing, worker crash: fall back to the port
        print(f"bench.py: reference package unavailable ({e}); timing the C port", file=sys.stderr)
        return None


def reference_arm(args):
    """The reference's own implementation on the box's host cores: the UNMODIFIED numba package from
    baseline/_ref, one clip per worker process like its single-threaded kernels (kind "reference");
    the C port under oracle/ when the package is absent or the chain has blocks it lacks (kind "port")."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cores = os.cpu_count() or 1
    n_frames = int(args.seconds * FS)
    pool = reference_pool(args.preset, n_frames, cores)
    times = []
    if pool is not None:
        kind, per_worker = "reference", 4                   # ~0.1 s per 10 s clip and core: ~0.4 s per step
        n_clips = per_worker * cores
        for _ in range(max(1, args.warmup)):
            pool.run(1)
        for _ in range(args.steps):
            _, dt = pool.run(per_worker)
            times.append(dt)
        pool.close()
        sample = f"{n_clips} clips x {args.seconds:g} s per step, {cores} worker processes (unmodified reference, numba)"
    else:
        kind = "port"
        n_clips = args.cpu_clips or max(cores, min(20 * cores, 512))
        cpu_run(args.preset, cores, n_frames, cores)
        for _ in range(args.warmup):
            cpu_run(args.preset, n_clips, n_frames, cores)
        for _ in range(args.steps):
            _, dt = cpu_run(args.preset, n_clips, n_frames, cores)
            times.append(dt)
        sample = f"{n_clips} clips x {args.seconds:g} s per step on {cores} threads (C port of the reference, one clip per thread)"
    total = sum(times)
    value = n_clips * n_frames * 2 * len(times) / total / 1e6
    _, _, scaling = shard_of(args, 0, world) if args.clips > 0 else (0, 0, "strong")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / max(1, len(times)),
        "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload(args, world),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------ B200 arm
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self, t_from=None, t_to=None):
        """Median SM clock and throttle reasons over the samples inside [t_from, t_to] (epoch
        seconds; the load window: warm-up + timed steps), all samples if none fall inside."""
        import datetime
        if self.proc:
            time.sleep(0.05)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        rows = [r for r in self.rows if len(r) > 8]
        if t_from is not None:
            inside = []
            for r in rows:
                try:
                    ts = datetime.datetime.strptime(r[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                except ValueError:
                    continue
                if t_from - 0.02 <= ts <= t_to + 0.02:
                    inside.append(r)
            if inside:
                rows = inside
        self.rows = rows
        sm = sorted(float(r[1]) for r in self.rows if len(r) > 8 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) > 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) > 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx[0] if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def synth_device(torch, n_clips, n_frames, first_clip, device):
    """Synthetic batch on the device with the shape of SURVEY 8d: white noise + a per-clip
    tone + a 50 ms full-scale burst every 2 s; (B, N, 2) float32."""
    g = torch.Generator(device=device)
    g.manual_seed(1000 + first_clip)
    x = torch.empty((n_clips, n_frames, 2), dtype=torch.float32, device=device)
    t = torch.arange(n_frames, device=device, dtype=torch.float32) / FS
    burst = ((torch.arange(n_frames, device=device) % (2 * FS)) < int(0.05 * FS))
    sq = 0.95 * torch.sign(torch.sin(2 * torch.pi * 997.0 * t) + 1e-12)
    step = 64
    for b0 in range(0, n_clips, step):
        nb = min(step, n_clips - b0)
        idx = torch.arange(first_clip + b0, first_clip + b0 + nb, device=device) % 48
        f = 110.0 * torch.pow(torch.tensor(2.0, device=device), idx.float() / 12.0)
        tone = 0.35 * torch.sin(2 * torch.pi * f[:, None] * t[None, :])
        noise = 0.25 * (2.0 * torch.rand((nb, n_frames, 2), generator=g, device=device) - 1.0)
        blk = noise + tone[:, :, None]
        blk = torch.where(burst[None, :, None], sq[None, :, None], blk)
        x[b0:b0 + nb] = blk
    return x


PRESET_ORDER = ("Robot Voice", "Cathedral", "Slapback Echo", "Clean Noise Removal", "Guitar Filter", "Rain Delay")


def timed_steps(torch, stream, barrier, fn, steps):
    """K launches of fn() on `stream`, one CUDA-event pair per step; returns (total ms, per-step ms)."""
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    barrier()
    ev[0].record(stream)
    for k in range(steps):
        fn()
        ev[k + 1].record(stream)
    barrier()
    return ev[0].elapsed_time(ev[-1]), [ev[k].elapsed_time(ev[k + 1]) for k in range(steps)]


def max_ms(torch, dist, world, dev, ms):
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


UNLISTED_CHAINS = {
    "filter>delay>reverb": [
        {"type": "filter", "params": {"filter_type": 0, "cutoff_hz": 4000, "q": 0.707}},
        {"type": "delay", "params": {"delay_ms": 250, "feedback": 0.35, "mix_wet": 0.5, "mix_dry": 1.0, "offset_ms": 15}},
        {"type": "reverb", "params": {"rt60_s": 1.5, "mix_wet": 0.3, "mix_dry": 0.9}},
    ],
    "gate>filter>octaver>delay>reverb": [
        {"type": "gate", "params": {"threshold_db": -45, "attack_ms": 5, "release_ms": 150}},
        {"type": "filter", "params": {"filter_type": 1, "cutoff_hz": 120, "q": 0.707}},
        {"type": "octaver", "params": {"semitones": 7, "mix": 0.4}},
        {"type": "delay", "params": {"delay_ms": 180, "feedback": 0.25, "mix_wet": 0.4, "mix_dry": 1.0, "offset_ms": 5}},
        {"type": "reverb", "params": {"rt60_s": 1.2, "mix_wet": 0.25, "mix_dry": 0.9}},
    ],
}


def sweep_leg(args, torch, dist, _native, file_chain, x, y, B, n_frames, world, dev, barrier, peak):
    """Every preset of app.py:41-71 over this rank's shard (BASELINE configs[4]): each one alone, then all six
    at once on separate streams so that the partial last wave of one preset's launch is filled by the others."""
    import synth
    res, pipes, keep = {}, {}, []
    stream = torch.cuda.current_stream()
    tmp = None
    samples_rank = B * n_frames * 2
    for name in PRESET_ORDER:
        chain = file_chain(synth.PRESETS[name], FS, channels_in=2)
        pipe, plans = chain.device_pipeline(n_frames)
        keep.append(plans)
        pipes[name] = pipe
        if pipe.n_segments > 1 and tmp is None:
            tmp = torch.empty_like(y)
        tp = tmp.data_ptr() if tmp is not None else y.data_ptr()
        run = lambda pipe=pipe, tp=tp: pipe(x.data_ptr(), y.data_ptr(), tp, B, stream.cuda_stream)
        run()                                                   # warm-up (plans allocate their scratch here)
        tot, _ = timed_steps(torch, stream, barrier, run, args.sweep_steps)
        ms = max_ms(torch, dist, world, dev, tot) / args.sweep_steps
        v = world * samples_rank / (ms * 1e-3) / 1e6
        res[name] = {"ms": ms, "value": v, "frac_of_hbm_roofline": (v / world) * 8e6 / 1e9 / peak}
    # chains a user may build that no kernel was specialised for (app.py:396-482 builds arbitrary ones):
    # what the generic interpreter kernel costs next to the listed shapes
    unlisted = {}
    for name, cfg in UNLISTED_CHAINS.items():
        chain = file_chain(cfg, FS, channels_in=2)
        pipe, plans = chain.device_pipeline(n_frames)
        keep.append(plans)
        run = lambda pipe=pipe: pipe(x.data_ptr(), y.data_ptr(), (tmp if tmp is not None else y).data_ptr(), B, stream.cuda_stream)
        run()
        tot, _ = timed_steps(torch, stream, barrier, run, args.sweep_steps)
        ms = max_ms(torch, dist, world, dev, tot) / args.sweep_steps
        v = world * samples_rank / (ms * 1e-3) / 1e6
        unlisted[name] = {"ms": ms, "value": v, "frac_of_hbm_roofline": (v / world) * 8e6 / 1e9 / peak,
                          "stages": [c["type"] for c in cfg],
                          "kernels": [p.info()["kernel"] for p in plans if hasattr(p, "info")]}
    # all six concurrently, chunked so that six output chunks fit beside the shard
    chunk = min(B, 1184)
    streams = [torch.cuda.Stream(device=dev) for _ in PRESET_ORDER]
    outs = [torch.empty((chunk, n_frames, 2), dtype=torch.float32, device=dev) for _ in PRESET_ORDER]
    tmps = {n: torch.empty((chunk, n_frames, 2), dtype=torch.float32, device=dev)
            for n in PRESET_ORDER if pipes[n].n_segments > 1}
    fsz = n_frames * 2 * 4

    def all_at_once():
        for b0 in range(0, B, chunk):
            nb = min(chunk, B - b0)
            for k, name in enumerate(PRESET_ORDER):
                tp = tmps[name].data_ptr() if name in tmps else outs[k].data_ptr()
                pipes[name](x.data_ptr() + b0 * fsz, outs[k].data_ptr(), tp, nb, streams[k].cuda_stream)

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    all_at_once()
    barrier()
    e0.record(stream)
    for st in streams:
        st.wait_stream(stream)
    for _ in range(args.sweep_steps):
        all_at_once()
    for st in streams:
        stream.wait_stream(st)
    e1.record(stream)
    barrier()
    ms = max_ms(torch, dist, world, dev, e0.elapsed_time(e1)) / args.sweep_steps
    seq_ms = sum(r["ms"] for r in res.values())
    out = {"presets": res, "steps": args.sweep_steps,
           "one_at_a_time": {"ms": seq_ms, "value": 6 * world * samples_rank / (seq_ms * 1e-3) / 1e6},
           "six_streams": {"ms": ms, "value": 6 * world * samples_rank / (ms * 1e-3) / 1e6, "chunk_clips": chunk},
           "unlisted_chains": unlisted,
           "unit": UNIT, "samples": "6 presets x the whole job's clips x frames x 2 channels per pass"}
    for plans in keep:
        for p in plans:
            p.close()
    del outs, tmps, tmp
    return out


def baseline_configs_leg(args, torch, _native, file_chain, dev, peak):
    """BASELINE configs[1], [2], [3] at their own sizes and the plot-side analysis, one GPU, inputs
    resident in HBM, each checked against the oracle: the numbers DESIGN quotes, measured by the default run."""
    import ctypes as C
    import numpy as np
    import synth
    from oracle import oracle as orc
    stream = torch.cuda.current_stream()
    sptr = stream.cuda_stream
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def median_ms(run, steps, flush_l2):
        run()
        torch.cuda.synchronize()
        ms = []
        for _ in range(steps):
            if flush_l2:
                flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            run()
            e1.record(stream)
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        return sorted(ms)[len(ms) // 2]

    def chain_case(name, B, seconds, steps, flush_l2, check):
        n = int(seconds * FS)
        cfg = chain_config(name)
        x = synth_device(torch, B, n, 0, dev)
        y = torch.empty_like(x)
        pipe, plans = file_chain(cfg, FS, channels_in=2).device_pipeline(n)
        ms = median_ms(lambda: pipe(x.data_ptr(), y.data_ptr(), y.data_ptr(), B, sptr), steps, flush_l2)
        worst, wsnr = 0.0, float("inf")
        for b in check:
            want = orc.run_file_path(cfg, np.ascontiguousarray(x[b].cpu().numpy()), FS)
            mx, snr = synth.err_stats(y[b].cpu().numpy(), want)
            worst, wsnr = max(worst, mx / max(1.0, float(np.max(np.abs(want))))), min(wsnr, snr)
        v = B * n * 2 / (ms * 1e-3) / 1e6
        r = {"clips": B, "seconds": seconds, "ms": ms, "value": v, "unit": UNIT, "frac_of_hbm_roofline": v * 8e6 / 1e9 / peak,
             "l2": "flushed before every timed launch" if flush_l2 else "inputs larger than L2",
             "kernels": [p.info()["kernel"] for p in plans if hasattr(p, "info")],
             "parity": {"max_abs_err": worst, "snr_db": wsnr if np.isfinite(wsnr) else None, "clips": list(check), "frames": n}}
        for p_ in plans:
            p_.close()
        return r

    out = {}
    c2 = chain_case("c2-biquad-cascade", 1, 60.0, 10, True, [0])
    c2["us_per_clip"] = c2["ms"] * 1e3
    out["configs[1] one 60 s clip, LP/HP/BP/peaking biquad cascade"] = c2
    c2["kernels"] = ["aes_biquad_scan_kernel (one CTA per 1024-frame tile, truncated look-back)"]
    c2b = chain_case("c2-biquad-cascade", 2368, 10.0, 5, False, [0, 2367])
    c2b["kernels"] = ["aes_biquad_seq_kernel (one thread per (clip, segment), the plain recurrence)"]
    out["configs[1] chain on a batch: 2368 clips x 10 s"] = c2b
    out["configs[2] 1024 clips x 10 s, distortion > octaver > delay"] = chain_case("c3-dist-octaver-delay", 1024, 10.0, 5, False, [0, 1023])

    # configs[3]: IR-convolution reverb, 3 s IR, 256 clips x 30 s
    n, B = 30 * FS, 256
    ir = orc.synthetic_ir(int(3.0 * FS))
    plan = _native.ConvReverbPlan(ir)
    x = synth_device(torch, B, n, 0, dev)
    y = torch.empty_like(x)
    ms = median_ms(lambda: plan.run_device(x.data_ptr(), y.data_ptr(), B, n, 0.7, 0.5, sptr), 5, False)
    n_chk = 200000
    xs = x[0, :n_chk].cpu().numpy()
    want = np.zeros_like(xs)
    orc.OConvReverb(ir, 0.7, 0.5).process_into(xs, want)
    mx, snr = synth.err_stats(y[0, :n_chk].cpu().numpy(), want)
    info = plan.info()
    v = B * n * 2 / (ms * 1e-3) / 1e6
    out["configs[3] 256 clips x 30 s, convolution reverb with a 3 s IR"] = {
        "clips": B, "seconds": 30.0, "ms": ms, "value": v, "unit": UNIT, "frac_of_hbm_roofline": v * 8e6 / 1e9 / peak, **info,
        "mac_fp32_tflops": B * -(-n // (info["fft_size"] // 2)) * info["fft_size"] * info["partitions"] * 8.0 / (ms * 1e-3) / 1e12,
        "kernels": ["aesc_fwd_kernel", "aesc_mac_kernel (TMA-staged IR tiles)", "aesc_inv_kernel"],
        "parity": {"max_abs_err": mx, "snr_db": snr, "clips": [0], "frames": n_chk, "oracle": "float64 fftconvolve restatement (ours)"}}
    plan.close()
    del x, y

    # plot-side analysis (02_custom.js:65-154): 4096 (original, processed) pairs of 16384 samples
    n_pairs, n_fft = 4096, 16384
    a = torch.rand(n_pairs, n_fft, device=dev) - 0.5
    t = torch.arange(n_fft, device=dev, dtype=torch.float32) / FS
    b = 0.4 * torch.sin(2 * torch.pi * 440.0 * t)[None, :] + 0.1 * a
    nb = n_fft // 2 + 1
    db, lin = torch.empty(n_pairs, 2, nb, device=dev), torch.empty(n_pairs, 2, nb, device=dev)
    ch, pk = torch.empty(n_pairs, 2, 12, device=dev), torch.empty(n_pairs, 2, device=dev)
    L = _native.lib()
    ms = median_ms(lambda: _native.check(L.aes_spectrum_chroma(
        C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()), n_pairs, n_fft, n_fft, float(FS), C.c_void_p(db.data_ptr()),
        C.c_void_p(lin.data_ptr()), C.c_void_p(ch.data_ptr()), C.c_void_p(pk.data_ptr()), C.c_void_p(sptr))), 5, False)
    _, want_db, want_ch, want_pk, _ = orc.spectrum_and_chroma(b[7].cpu().numpy(), float(FS))
    got_db = db[7, 1].cpu().numpy()
    loud = want_db > -80.0
    out["plot-side spectrum + chromagram, 4096 pairs x 16384 samples"] = {
        "ms": ms, "spectra_per_s": 2 * n_pairs / (ms * 1e-3), "value": 2 * n_pairs * n_fft / (ms * 1e-3) / 1e6, "unit": UNIT,
        "kernel": "aesa_kernel (one complex FFT per pair)",
        "parity": {"max_db_err_above_-80dB": float(np.max(np.abs(got_db[loud] - want_db[loud]))),
                   "max_chroma_err": float(np.max(np.abs(ch[7, 1].cpu().numpy() - want_ch))),
                   "peak_freq": [float(pk[7, 1]), float(want_pk)], "oracle": "float64 restatement of the page's JavaScript (ours)"}}
    return out


def copy_ceiling(torch, dist, world, dev, barrier, xh, yh, reps):
    """What the PCIe path alone delivers: H2D of `xh` and D2H into `yh` (pinned, same bytes as the end-to-end
    leg), both directions at once on two streams, all ranks at the same time, no kernel."""
    xd = torch.empty(xh.shape, dtype=torch.from_numpy(xh[:1]).dtype, device=dev)
    yd = torch.empty(yh.shape, dtype=torch.from_numpy(yh[:1]).dtype, device=dev)
    tx, ty = torch.from_numpy(xh), torch.from_numpy(yh)
    s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    def once():
        with torch.cuda.stream(s1):
            xd.copy_(tx, non_blocking=True)
        with torch.cuda.stream(s2):
            ty.copy_(yd, non_blocking=True)
    once()
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        once()
    torch.cuda.synchronize()
    dt = max_ms(torch, dist, world, dev, (time.perf_counter() - t0) * 1e3) * 1e-3 / reps
    return {"h2d_gbs_per_gpu": xh.nbytes / dt / 1e9, "d2h_gbs_per_gpu": yh.nbytes / dt / 1e9, "seconds": dt,
            "pinned": bool(tx.is_pinned() and ty.is_pinned())}


def gather_leg(args, torch, dist, _native, plan, x, B, n_frames, rank, world, dev, barrier):
    """NCCL over NVLink, used only to gather results (north_star): int16 PCM of a bounded sample per rank,
    all_gather_into_tensor into a preallocated buffer -- alone, and with chunk k's gather on a side stream
    while chunk k+1 is computed."""
    from audioblocks.sharding import all_gather_pcm
    Bg = min(B, args.gather_clips)
    t = torch.tensor([Bg], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    Bg = int(t.item())
    nchunk = 4
    per = max(1, Bg // nchunk)
    Bg = per * nchunk
    stream = torch.cuda.current_stream()
    comm = torch.cuda.Stream(device=dev)
    yq = [torch.empty((per, n_frames, 2), dtype=torch.int16, device=dev) for _ in range(2)]
    full = [torch.empty((world * per, n_frames, 2), dtype=torch.int16, device=dev) for _ in range(2)]
    fsz = n_frames * 2 * 4

    def compute(k):
        plan.run_device(x.data_ptr() + k * per * fsz, _native.FMT_F32_STEREO, yq[k & 1].data_ptr(), _native.FMT_I16_STEREO,
                        per, n_frames, stream.cuda_stream)

    def gather(k, st):
        with torch.cuda.stream(st):
            all_gather_pcm(yq[k & 1], out=full[k & 1])

    for k in range(2):                                      # warm-up: communicator, plan scratch
        compute(k)
        gather(k, stream)
    e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    # (1) gather alone
    barrier()
    e[0].record(stream)
    for k in range(nchunk):
        gather(k, stream)
    e[1].record(stream)
    barrier()
    g_ms = max_ms(torch, dist, world, dev, e[0].elapsed_time(e[1]))
    # (2) compute alone, (3) compute of chunk k+1 overlapped with the gather of chunk k
    barrier()
    e[0].record(stream)
    for k in range(nchunk):
        compute(k)
    e[1].record(stream)
    barrier()
    c_ms = max_ms(torch, dist, world, dev, e[0].elapsed_time(e[1]))
    done = [torch.cuda.Event() for _ in range(nchunk)]
    freed = [torch.cuda.Event() for _ in range(nchunk)]
    barrier()
    e[2].record(stream)
    for k in range(nchunk):
        if k >= 2:
            stream.wait_event(freed[k - 2])                 # the gather of chunk k-2 has read this buffer
        compute(k)
        done[k].record(stream)
        comm.wait_event(done[k])
        gather(k, comm)
        freed[k].record(comm)
    stream.wait_stream(comm)
    e[3].record(stream)
    barrier()
    o_ms = max_ms(torch, dist, world, dev, e[2].elapsed_time(e[3]))
    recv = (world - 1) * Bg * n_frames * 2 * 2              # bytes every rank receives from its peers
    same = bool(torch.equal(full[(nchunk - 1) & 1][rank * per:(rank + 1) * per], yq[(nchunk - 1) & 1]))
    return {"clips_per_rank": Bg, "chunks": nchunk, "dtype": "int16 PCM (as bytes)", "collective": "all_gather_into_tensor",
            "bytes_received_per_rank": recv, "ms": g_ms, "gbs": recv / (g_ms * 1e-3) / 1e9,
            "busbw_gbs": recv / (g_ms * 1e-3) / 1e9,   # all-gather: bus bandwidth = received bytes / time
            "compute_only_ms": c_ms, "compute_plus_gather_overlapped_ms": o_ms,
            "gather_hidden_frac": max(0.0, min(1.0, (c_ms + g_ms - o_ms) / g_ms)) if g_ms > 0 else None,
            "own_shard_round_trips": same}


def b200_arm(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import audioblocks  # noqa: F401
    from audioblocks import _native
    from audioblocks.engine import file_chain
    import synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the B200 arm has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = _native.lib()
    _native.check(L.aes_set_device(local))
    numa_cpus = None
    if world > 1:                   # node-local staging memory: bind before any pinned allocation
        from audioblocks.sharding import bind_host_to_gpu
        pr = torch.cuda.get_device_properties(local)
        if hasattr(pr, "pci_bus_id"):
            numa_cpus = bind_host_to_gpu("%04x:%02x:%02x.0" % (getattr(pr, "pci_domain_id", 0), pr.pci_bus_id,
                                                               getattr(pr, "pci_device_id", 0)))
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()              # nvidia-smi takes a while to start: launch it before the data is built

    if args.preset == "c4-convreverb":
        if args.clips <= 0:
            args.clips = 256
        return conv_arm(args, torch, dist, _native, rank, world, local, dev)
    n_frames = int(args.seconds * FS)
    B, total_clips, scaling = shard_of(args, rank, world)
    first_clip = rank * B if args.clips > 0 else __import__("audioblocks.sharding", fromlist=["shard_range"]).shard_range(args.total_clips, rank, world)[0]
    cfg = chain_config(args.preset)
    chain = file_chain(cfg, FS, channels_in=2)            # build@1024 + warm-up, as engine.py:86-99
    pipe, plans = chain.device_pipeline(n_frames)         # re-prepared at the file's frame count
    fused = pipe.n_segments == 1 and isinstance(plans[0], _native.ChainPlan)
    plan = plans[0] if fused else None
    info = plan.info() if fused else {"tile_frames": None, "smem_bytes": None, "ctas_per_sm": None,
                                      "kernel": f"{pipe.n_segments} segments (whole-clip FFT block between fused runs)"}

    x = synth_device(torch, B, n_frames, first_clip, dev)
    y = torch.empty_like(x)
    tmp = torch.empty_like(x) if pipe.n_segments > 1 else y
    stream = torch.cuda.current_stream()
    sptr = stream.cuda_stream

    def step():
        pipe(x.data_ptr(), y.data_ptr(), tmp.data_ptr(), B, sptr)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    t_load0 = time.time()
    for _ in range(max(3, args.warmup)):
        step()
    barrier()
    launches0 = L.aes_launch_count()
    # timing rule: inputs larger than L2, or L2 flushed between timed iterations.  The default workload
    # (tens of GB per step) is the former; a side workload that fits the 126 MB L2 (one 60 s clip: 46 MB)
    # gets a 256 MB buffer written before every timed step, and the steps are timed one by one.
    l2_fits = B * n_frames * 16 < 2 * 126e6
    if l2_fits:
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        eva = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
        evb = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
        barrier()
        for k in range(args.steps):
            flush.zero_()
            eva[k].record(stream)
            step()
            evb[k].record(stream)
        barrier()
        per_launch_ms = [eva[k].elapsed_time(evb[k]) for k in range(args.steps)]
        total_ms = sum(per_launch_ms)
        del flush
    else:
        total_ms, per_launch_ms = timed_steps(torch, stream, barrier, step, args.steps)
    launches = L.aes_launch_count() - launches0
    clk = clocks.stop(t_load0, time.time()) if rank == 0 else None
    total_ms = max_ms(torch, dist, world, dev, total_ms)

    # The same K passes with consecutive passes alternating between two streams and two output buffers: the
    # partial last wave of one pass (8192 / N clips over 296 resident CTAs) is filled by the next pass's first
    # CTAs, as it is in a job that has more than one batch.  Reported beside `value`, never instead of it.
    pipelined = None
    if fused and not l2_fits and 3 * x.numel() * 4 < 150e9:
        y2 = torch.empty_like(x)
        # (a plan owns the feedback-delay lines of its CTAs: launches that overlap need a plan each)
        pipe_b, plans_b = file_chain(cfg, FS, channels_in=2).device_pipeline(n_frames)
        pipe_b(x.data_ptr(), y2.data_ptr(), y2.data_ptr(), B, sptr)        # allocates its scratch
        sa, sb = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        sa.wait_stream(stream); sb.wait_stream(stream)
        for k in range(args.steps):
            st_k, y_k, pipe_k = (sa, y, pipe) if k % 2 == 0 else (sb, y2, pipe_b)
            pipe_k(x.data_ptr(), y_k.data_ptr(), y_k.data_ptr(), B, st_k.cuda_stream)
        stream.wait_stream(sa); stream.wait_stream(sb)
        e1.record(stream)
        barrier()
        pms = max_ms(torch, dist, world, dev, e0.elapsed_time(e1))
        same = bool(torch.equal(y[-1], y2[-1])) if args.steps > 1 else None
        pipelined = {"value": total_clips * n_frames * 2 * args.steps / (pms * 1e-3) / 1e6, "unit": UNIT,
                     "ms_per_step": pms / args.steps, "streams": 2, "steps": args.steps, "buffers_identical": same,
                     "what": "the timed passes again, pass k on stream k mod 2 (a plan and an output buffer per stream)"}
        del y2
        for p_ in plans_b:
            p_.close()

    # parity of the timed buffers against the oracle: the FIRST and the LAST clip of this rank's shard, whole clips
    parity = None
    if rank == 0 and not args.no_cpu:
        from oracle import oracle as orc
        worst, wsnr = 0.0, float("inf")
        which = sorted({0, B - 1})
        for b in which:
            xs = np.ascontiguousarray(x[b].cpu().numpy())
            want = orc.run_file_path(cfg, xs, FS)
            mx, snr = synth.err_stats(y[b].cpu().numpy(), want)
            worst, wsnr = max(worst, mx), min(wsnr, snr)
        parity = {"max_abs_err": worst, "snr_db": wsnr if np.isfinite(wsnr) else None, "clips": which, "frames": n_frames,
                  "bar": "max-abs 1e-5 of full scale, SNR 100 dB (north_star)", "ok": bool(worst <= 1e-5 and wsnr >= 100.0)}
        if not parity["ok"]:
            print(f"bench.py: PARITY FAILED against the oracle: {parity}", file=sys.stderr)

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"

    # ---- every preset over the same shard
    sweep = None
    if not args.no_sweep and args.clips <= 0 and args.preset == "Rain Delay":
        sweep = sweep_leg(args, torch, dist, _native, file_chain, x, y, B, n_frames, world, dev, barrier, peak)

    # ---- BASELINE's other configurations at their own sizes (N = 1 only: none of them shards further)
    others = None
    if sweep is not None and world == 1 and not args.no_cpu:
        others = baseline_configs_leg(args, torch, _native, file_chain, dev, peak)

    # ---- results gathered over NCCL (N > 1 only)
    gather = None
    if world > 1 and fused and not args.no_gather:
        gather = gather_leg(args, torch, dist, _native, plan, x, B, n_frames, rank, world, dev, barrier)

    # ---- end to end through the host-buffer call (pinned host memory, H2D + D2H timed)
    e2e = None
    Be = min(B, args.e2e_clips)
    if not args.no_e2e and not fused:
        xh = _native.pinned_empty((Be, n_frames, 2), np.float32)
        yh = np.empty((Be, n_frames, 2), np.float32)
        torch.from_numpy(xh).copy_(x[:Be])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            file_chain(cfg, FS, channels_in=2).process_batch(xh, yh)
        e2e_s = time.perf_counter() - t0
        e2e = {"value": world * Be * n_frames * 2 * args.e2e_steps / e2e_s / 1e6, "unit": UNIT,
               "h2d_bytes_per_step": int(xh.nbytes), "d2h_bytes_per_step": int(yh.nbytes), "steps": args.e2e_steps,
               "api": "EffectsChain.process_batch (host buffers, one H2D/D2H per chain segment)"}
    elif not args.no_e2e:
        xh = _native.pinned_empty((Be, n_frames, 2), np.float32)
        yh = _native.pinned_empty((Be, n_frames, 2), np.float32)
        torch.from_numpy(xh).copy_(x[:Be])
        torch.cuda.synchronize()
        fmt = _native.FMT_F32_STEREO
        plan.run_host(xh, fmt, yh, fmt, Be, n_frames)     # warm-up: allocates the staging slots
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            plan.run_host(xh, fmt, yh, fmt, Be, n_frames)
        torch.cuda.synchronize()
        e2e_s = max_ms(torch, dist, world, dev, (time.perf_counter() - t0) * 1e3) * 1e-3
        e2e_ok = bool(np.array_equal(yh[0, :4096], y[0, :4096].cpu().numpy()))
        e2e = {"value": world * Be * n_frames * 2 * args.e2e_steps / e2e_s / 1e6, "unit": UNIT,
               "h2d_bytes_per_step": int(xh.nbytes), "d2h_bytes_per_step": int(yh.nbytes),
               "steps": args.e2e_steps, "matches_device_path": e2e_ok,
               "sample": f"{Be} clips of the shard per GPU and step (bounded: pinned host memory)",
               "host_cpus_bound_to_gpu_node": len(numa_cpus) if numa_cpus else None,
               "api": "EffectsChain.prepare_batch(...).run_host -> aes_chain_process_host (pinned host buffers)"}
        ceil = copy_ceiling(torch, dist, world, dev, barrier, xh, yh, args.e2e_steps)
        ceil["value"] = world * Be * n_frames * 2 / ceil["seconds"] / 1e6
        ceil["unit"] = UNIT
        ceil["what"] = "cudaMemcpyAsync H2D + D2H of the same pinned bytes, both directions at once, all ranks at once, no kernel"
        e2e["copy_ceiling"] = ceil
        e2e["frac_of_copy_ceiling"] = e2e["value"] / ceil["value"]
        # the same call fed like the WAV-file route feeds it (engine.py:78-84,104-105): int16 stereo
        # PCM in (down-mixed on the device), int16 stereo PCM out -- half the PCIe bytes per frame
        xq = _native.pinned_empty((Be, n_frames, 2), np.int16)
        yq = _native.pinned_empty((Be, n_frames, 2), np.int16)
        np.multiply(xh, 32767.0, out=yh)
        xq[...] = yh.astype(np.int16)
        plan.run_host(xq, _native.FMT_I16_DOWNMIX, yq, _native.FMT_I16_STEREO, Be, n_frames)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            plan.run_host(xq, _native.FMT_I16_DOWNMIX, yq, _native.FMT_I16_STEREO, Be, n_frames)
        pcm_s = max_ms(torch, dist, world, dev, (time.perf_counter() - t0) * 1e3) * 1e-3
        pceil = copy_ceiling(torch, dist, world, dev, barrier, xq, yq, args.e2e_steps)
        pv = world * Be * n_frames * 2 * args.e2e_steps / pcm_s / 1e6
        pcv = world * Be * n_frames * 2 / pceil["seconds"] / 1e6
        # Headline = the WAV-file route's own formats (engine.py:78-84 reads int16 PCM and down-mixes it,
        # engine.py:104-110 writes int16 PCM): 4 bytes in + 4 bytes out per stereo frame.  The float32-buffer
        # variant (8 + 8 bytes per frame, what EffectsChain.process takes) is kept beside it.
        f32 = {k: e2e[k] for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step", "copy_ceiling",
                                   "frac_of_copy_ceiling", "matches_device_path")}
        e2e.update({"value": pv, "h2d_bytes_per_step": int(xq.nbytes), "d2h_bytes_per_step": int(yq.nbytes),
                    "formats": "int16 stereo PCM in (down-mixed on the device like engine.py:81-84), int16 stereo PCM out "
                               "(clip, x32767, truncate like engine.py:104-105)",
                    "copy_ceiling": {"value": pcv, "unit": UNIT, "h2d_gbs_per_gpu": pceil["h2d_gbs_per_gpu"],
                                     "d2h_gbs_per_gpu": pceil["d2h_gbs_per_gpu"], "pinned": pceil["pinned"],
                                     "what": ceil["what"]},
                    "frac_of_copy_ceiling": pv / pcv, "f32_buffers": f32})
        e2e.pop("matches_device_path", None)
        del xh, yh, xq, yq

    if rank == 0:
        samples_per_step = world * B * n_frames * 2 if scaling == "weak" else total_clips * n_frames * 2
        value = samples_per_step * args.steps / (total_ms * 1e-3) / 1e6
        k_ms = sum(per_launch_ms) / len(per_launch_ms)
        achieved = B * n_frames * 2 * 8 / (k_ms * 1e-3) / 1e9
        traffic, traffic_src = NCU_DRAM_TRAFFIC.get((args.preset, n_frames), (None, None))
        if traffic is not None:
            traffic = traffic * B                            # recorded per clip: the capture ran another batch size
        kernel = info["kernel"]
        if args.preset == "c2-biquad-cascade" and info["ctas_per_sm"] and not os.environ.get("AES_NO_SCAN") and \
                B < info["ctas_per_sm"] * torch.cuda.get_device_properties(0).multi_processor_count:
            # fewer clips than resident CTAs: launch_chain (aes_chain.cu) takes the time-parallel scan
            kernel = "aes_biquad_scan_kernel (one CTA per 1024-frame tile, %s look-back)" % (
                "chained" if os.environ.get("AES_SCAN_CHAINED") else "truncated")
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {**workload(args, world), "clips_this_rank": B, "tile_frames": info["tile_frames"],
                       "smem_bytes_per_cta": info["smem_bytes"], "ctas_per_sm": info["ctas_per_sm"],
                       **({"l2": "fits L2: a 256 MB buffer is written before every timed step, steps timed one by one"}
                          if l2_fits else {})},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": peak_src,
                         "kernel": kernel, "algorithmic_bytes_per_launch": B * n_frames * 16,
                         "launch_ms": k_ms},
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clk, "parity": parity,
        }
        if pipelined is not None:
            line["pipelined_steps"] = pipelined
        if sweep is not None:
            line["sweep"] = sweep
        if gather is not None:
            line["gather"] = gather
        if others is not None:
            line["baseline_configs"] = others
        if not args.no_cpu and world == 1:          # the CPU baseline is reported at N = 1 only
            cores = os.cpu_count() or 1
            n_cpu = args.cpu_clips or max(cores, min(20 * cores, 512))     # ~20 s of CPU work on all threads
            v, dt = cpu_run(args.preset, n_cpu, n_frames, cores, reps=2)
            numpy_fft = any(c["type"] in ("spectral", "convreverb") for c in cfg)
            port = {"value": v, "unit": UNIT, "cores": 1 if numpy_fft else cores, "kind": "port",
                    "sample": (f"2 clips x {args.seconds:g} s, 1 thread (numpy FFT block), {dt:.2f} s" if numpy_fft
                               else f"{n_cpu} clips x {args.seconds:g} s, {cores} threads, {dt:.2f} s")}
            line["cpu_baseline"] = port
            pool = reference_pool(args.preset, n_frames, cores)
            if pool is not None:                    # the unmodified reference (numba), one clip per worker process
                pool.run(1)
                per_worker = 8
                rv, rdt = pool.run(per_worker)
                pool.close()
                line["cpu_baseline"] = {"value": rv, "unit": UNIT, "cores": cores, "kind": "reference",
                                        "sample": f"{per_worker * cores} clips x {args.seconds:g} s, {cores} worker processes "
                                                  f"(unmodified reference package, numba), {rdt:.2f} s",
                                        "c_port_same_box": port}
        print(json.dumps(line))
    for p in plans:
        p.close()
    if world > 1:
        dist.destroy_process_group()


# dram__bytes_read.sum + dram__bytes_write.sum of the chain kernel, per launch, from the committed
# `ncu --set full` capture of exactly this workload: (preset, clips per GPU, frames per clip) -> bytes
# (the capture ran 1184 clips per launch; stored per clip so that any shard size can quote it)
NCU_DRAM_TRAFFIC = {
    ("Rain Delay", 480000): ((4.657284e9 + 5.812085e9) / 1184, "profiles/r2i_ncu_rv_kernel.csv (1184 clips per launch, scaled to this shard)"),
}


def main():
    # exactly ONE line on stdout: libraries that print there (NCCL's version banner) go to stderr instead
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = real_stdout
    args = parse()
    if args.impl == "reference":
        reference_arm(args)
    else:
        b200_arm(args)


if __name__ == "__main__":
    main()
