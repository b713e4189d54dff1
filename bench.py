#!/usr/bin/env python
"""bench.py -- throughput of the offline effect-chain hot path.

    python bench.py --gpus N --steps K --warmup W            # B200 arm (this repo)
    python bench.py --impl reference --gpus N --steps K ...  # reference arm: the CPU algorithm on host cores

A "step" is one pass of the fused preset chain over one batch of synthetic 48 kHz
stereo clips that is already resident in HBM (`value`), and the same batch pushed
through the reference-facing host-buffer call with H2D/D2H inside the timed region
(`e2e`).  1 sample = one float32 value of one channel of output;
algorithmic traffic is 8 bytes per sample (read once, write once).

Under torchrun (N>1) every rank processes its own shard of clips (no data-path
collective: clips are independent), times it with CUDA events, and rank 0 reports
total samples / max-over-ranks time ("scaling": "weak").
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
    ap.add_argument("--clips", type=int, default=1184, help="clips per GPU (default 8 per SM: whole clips for every resident CTA)")
    ap.add_argument("--seconds", type=float, default=10.0, help="clip length")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
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
    mac_flops = B * nblk * info["fft_size"] * info["partitions"] * 16.0
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
                     "kernel": "aesc_fft_blocks + aesc_mac + aesc_ifft_mix", "mac_fp32_tflops": mac_flops / (ms * 1e-3) / 1e12},
        "gpu_launches": int(L.aes_launch_count() - l0), "parity": {"max_abs_err": mx, "snr_db": snr, "frames": n_chk},
    }))
    plan.close()


def chain_config(name):
    import synth
    if name in synth.PRESETS:
        return synth.PRESETS[name]
    return EXTRA_CHAINS[name]


def workload(args):
    return {"workload": f"'{args.preset}' preset chain (app.py:41-71) on {args.clips} synthetic "
                        f"{args.seconds:g} s 48 kHz stereo float32 clips per GPU (BASELINE configs[4]-style shard)",
            "preset": args.preset, "clips_per_gpu": args.clips, "frames_per_clip": int(args.seconds * FS),
            "sample_rate": FS, "l2": "inputs larger than L2 (no flush needed)", "parallelism": "clip-sharded"}


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


def reference_arm(args):
    """The reference's own algorithm on the box's host cores.  The reference is pure
    Python+numba and cannot travel to the GPU box, so this times the C port under
    oracle/ (kind "port"), one clip per thread like the reference's single-threaded
    kernels, on all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_frames = int(args.seconds * FS)
    n_clips = args.cpu_clips or max(cores, min(20 * cores, 512))      # same sample as the B200 arm's cpu_baseline
    cpu_run(args.preset, cores, n_frames, cores)
    times = []
    for _ in range(args.warmup):
        cpu_run(args.preset, n_clips, n_frames, cores)
    for _ in range(args.steps):
        _, dt = cpu_run(args.preset, n_clips, n_frames, cores)
        times.append(dt)
    total = sum(times)
    value = n_clips * n_frames * 2 * len(times) / total / 1e6
    sample = f"{n_clips} clips x {args.seconds:g} s per step on {cores} threads (one clip per thread)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / max(1, len(times)),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload(args),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
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
        return conv_arm(args, torch, dist, _native, rank, world, local, dev)
    n_frames = int(args.seconds * FS)
    B = args.clips
    cfg = chain_config(args.preset)
    chain = file_chain(cfg, FS, channels_in=2)            # build@1024 + warm-up, as engine.py:86-99
    pipe, plans = chain.device_pipeline(n_frames)         # re-prepared at the file's frame count
    fused = pipe.n_segments == 1 and isinstance(plans[0], _native.ChainPlan)
    plan = plans[0] if fused else None
    info = plan.info() if fused else {"tile_frames": None, "smem_bytes": None, "ctas_per_sm": None,
                                      "kernel": f"{pipe.n_segments} segments (whole-clip FFT block between fused runs)"}

    x = synth_device(torch, B, n_frames, rank * B, dev)
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
    # (9 GB per step) is the former; a side workload that fits the 126 MB L2 (one 60 s clip: 46 MB)
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
        launches = L.aes_launch_count() - launches0
        per_launch_ms = [eva[k].elapsed_time(evb[k]) for k in range(args.steps)]
        total_ms = sum(per_launch_ms)
        del flush
    else:
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
        barrier()
        ev[0].record(stream)
        for k in range(args.steps):
            step()
            ev[k + 1].record(stream)
        barrier()
        launches = L.aes_launch_count() - launches0
        total_ms = ev[0].elapsed_time(ev[-1])
        per_launch_ms = [ev[k].elapsed_time(ev[k + 1]) for k in range(args.steps)]
    clk = clocks.stop(t_load0, time.time()) if rank == 0 else None

    # parity spot check of the timed buffers (first clip of this rank) against the oracle
    parity = None
    if rank == 0 and not args.no_cpu:
        from oracle import oracle as orc
        n_chk = min(n_frames, 96000)
        xs = x[0, :n_chk].cpu().numpy()
        want = orc.run_file_path(cfg, np.ascontiguousarray(xs), FS)
        got = y[0, :n_chk].cpu().numpy()
        mx, snr = synth.err_stats(got, want)
        parity = {"max_abs_err": mx, "snr_db": snr if np.isfinite(snr) else None, "frames": n_chk}

    # ---- end to end through the host-buffer call (pinned host memory, H2D + D2H timed)
    e2e = None
    if not args.no_e2e and not fused:
        xh = _native.pinned_empty((B, n_frames, 2), np.float32)
        yh = np.empty((B, n_frames, 2), np.float32)
        torch.from_numpy(xh).copy_(x)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            file_chain(cfg, FS, channels_in=2).process_batch(xh, yh)
        e2e_s = time.perf_counter() - t0
        e2e = {"value": world * B * n_frames * 2 * args.e2e_steps / e2e_s / 1e6, "unit": UNIT,
               "h2d_bytes_per_step": int(xh.nbytes), "d2h_bytes_per_step": int(yh.nbytes), "steps": args.e2e_steps,
               "api": "EffectsChain.process_batch (host buffers, one H2D/D2H per chain segment)"}
    elif not args.no_e2e:
        Be = B
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
        e2e_s = time.perf_counter() - t0
        e2e_ok = bool(np.array_equal(yh[0, :4096], y[0, :4096].cpu().numpy()))
        t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e = {"value": world * Be * n_frames * 2 * args.e2e_steps / float(t.item()) / 1e6, "unit": UNIT,
               "h2d_bytes_per_step": int(xh.nbytes), "d2h_bytes_per_step": int(yh.nbytes),
               "steps": args.e2e_steps, "matches_device_path": e2e_ok,
               "host_cpus_bound_to_gpu_node": len(numa_cpus) if numa_cpus else None,
               "api": "EffectsChain.prepare_batch(...).run_host -> aes_chain_process_host (pinned host buffers)"}
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
        t = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        pcm_s = float(t.item())
        e2e["pcm16_file_route"] = {"value": world * Be * n_frames * 2 * args.e2e_steps / pcm_s / 1e6, "unit": UNIT,
                                   "h2d_bytes_per_step": int(xq.nbytes), "d2h_bytes_per_step": int(yq.nbytes)}
        del xh, yh, xq, yq

    t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())

    if rank == 0:
        samples_per_step = world * B * n_frames * 2
        value = samples_per_step * args.steps / (total_ms * 1e-3) / 1e6
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        k_ms = sum(per_launch_ms) / len(per_launch_ms)
        achieved = B * n_frames * 2 * 8 / (k_ms * 1e-3) / 1e9
        traffic, traffic_src = NCU_DRAM_TRAFFIC.get((args.preset, B, n_frames), (None, None))
        kernel = info["kernel"]
        if args.preset == "c2-biquad-cascade" and info["ctas_per_sm"] and not os.environ.get("AES_NO_SCAN") and \
                B < info["ctas_per_sm"] * torch.cuda.get_device_properties(0).multi_processor_count:
            # fewer clips than resident CTAs: launch_chain (aes_chain.cu) takes the time-parallel scan
            kernel = "aes_biquad_scan_kernel (one CTA per 1024-frame tile, %s look-back)" % (
                "chained" if os.environ.get("AES_SCAN_CHAINED") else "truncated")
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {**workload(args), "tile_frames": info["tile_frames"], "smem_bytes_per_cta": info["smem_bytes"],
                       "ctas_per_sm": info["ctas_per_sm"],
                       **({"l2": "fits L2: a 256 MB buffer is written before every timed step, steps timed one by one"}
                          if l2_fits else {})},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": peak_src,
                         "kernel": kernel, "algorithmic_bytes_per_launch": B * n_frames * 16,
                         "launch_ms": k_ms},
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clk, "parity": parity,
        }
        if not args.no_cpu and world == 1:          # the CPU baseline is reported at N = 1 only
            cores = os.cpu_count() or 1
            n_cpu = args.cpu_clips or max(cores, min(20 * cores, 512))     # ~20 s of CPU work on all threads
            v, dt = cpu_run(args.preset, n_cpu, n_frames, cores, reps=2)
            numpy_fft = any(c["type"] in ("spectral", "convreverb") for c in cfg)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": 1 if numpy_fft else cores, "kind": "port",
                                    "sample": (f"2 clips x {args.seconds:g} s, 1 thread (numpy FFT block), {dt:.2f} s" if numpy_fft
                                               else f"{n_cpu} clips x {args.seconds:g} s, {cores} threads, {dt:.2f} s")}
        print(json.dumps(line))
    for p in plans:
        p.close()
    if world > 1:
        dist.destroy_process_group()


# dram__bytes_read.sum + dram__bytes_write.sum of the chain kernel, per launch, from the committed
# `ncu --set full` capture of exactly this workload: (preset, clips per GPU, frames per clip) -> bytes
NCU_DRAM_TRAFFIC = {
    ("Rain Delay", 1184, 480000): (4.613200e9 + 5.971754e9, "profiles/r1o_ncu_chain_kernel.csv"),
}


def main():
    args = parse()
    if args.impl == "reference":
        reference_arm(args)
    else:
        b200_arm(args)


if __name__ == "__main__":
    main()
