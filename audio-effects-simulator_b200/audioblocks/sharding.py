"""Batch sharding across the GPUs of one box (one process per GPU, torch.distributed).

Clips and (clip, preset) pairs are independent: rank r processes a
contiguous slice of the batch with no exchange during compute.  The only
collective is the optional gather of results over NCCL/NVLink (gloo on CPU in the
tests); it is never inside the timed compute region -- gathering full float32
audio onto one GPU costs ~25x the compute at roofline, so callers gather int16 or
per-clip statistics when they can."""
from __future__ import annotations


def shard_range(total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced [start, stop) of `total` items for `rank` of `world`:
    the first total % world ranks take one extra item."""
    if world < 1 or not 0 <= rank < world or total < 0:
        raise ValueError("bad shard request")
    base, extra = divmod(total, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_sizes(total: int, world: int) -> list[int]:
    return [shard_range(total, r, world)[1] - shard_range(total, r, world)[0] for r in range(world)]


def gather_clips(y_local, total: int, dst: int | None = 0, group=None):
    """Gather per-rank results (n_local, frames, 2) into (total, frames, 2).

    dst=None -> every rank gets the full batch (all_gather); otherwise only `dst`
    does (others return None).  Works for CUDA tensors over NCCL and CPU tensors
    over gloo; ragged shards are padded to the largest one for the collective."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = shard_sizes(total, world)
    if y_local.shape[0] != sizes[rank]:
        raise ValueError(f"rank {rank}: expected {sizes[rank]} clips, got {y_local.shape[0]}")
    # NCCL has no 16-bit integer type: ship int16 PCM as bytes and view it back
    as_bytes = y_local.dtype == torch.int16
    orig_dtype = y_local.dtype
    if as_bytes:
        y_local = y_local.contiguous().view(torch.uint8)
    cap = max(sizes)
    pad = y_local
    if y_local.shape[0] < cap:
        pad = torch.zeros((cap,) + tuple(y_local.shape[1:]), dtype=y_local.dtype, device=y_local.device)
        pad[: y_local.shape[0]] = y_local
    pad = pad.contiguous()
    if dst is None:
        bufs = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(bufs, pad, group=group)
    else:
        bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
        dist.gather(pad, bufs, dst=dst, group=group)
        if rank != dst:
            return None
    full = torch.cat([b[:n] for b, n in zip(bufs, sizes)], dim=0)
    return full.view(orig_dtype) if as_bytes else full


def all_gather_pcm(y_local, out=None, group=None):
    """Gather equal-sized int16 PCM shards (n_local, frames, 2) of every rank into one preallocated
    (world * n_local, frames, 2) tensor with a single all_gather_into_tensor: no padding, no list of
    per-rank tensors, no concatenation.  NCCL has no 16-bit integer type, so the PCM travels as bytes.
    `out` may be passed to reuse the destination (bench.py double-buffers it to overlap the gather of
    one chunk with the compute of the next).  Shards must have the same shape on every rank
    (`shard_range` of a batch that the world size divides, or a fixed chunk size)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if y_local.dtype != torch.int16:
        raise TypeError("all_gather_pcm gathers int16 PCM (quantise on the device first: out_fmt I16)")
    y_local = y_local.contiguous()
    if out is None:
        out = torch.empty((world * y_local.shape[0],) + tuple(y_local.shape[1:]), dtype=torch.int16, device=y_local.device)
    elif out.dtype != torch.int16 or out.numel() != world * y_local.numel() or not out.is_contiguous():
        raise ValueError("out must be a contiguous int16 tensor of world x the local shard")
    dist.all_gather_into_tensor(out.view(torch.uint8).reshape(-1), y_local.view(torch.uint8).reshape(-1), group=group)
    return out


def max_over_ranks(seconds: float, device=None, group=None) -> float:
    """Timing reduction used by bench.py: the slowest rank defines the step."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([seconds], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def _parse_cpulist(text: str) -> set[int]:
    cpus: set[int] = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_host_to_gpu(pci_bus_id: str) -> list[int] | None:
    """Pin this process to the CPUs that sit on the GPU's own PCIe root / NUMA node
    (`/sys/bus/pci/devices/<id>/local_cpulist`), BEFORE any pinned host buffer is allocated, so
    staging memory is node-local and the H2D/D2H copies of different ranks do not cross the
    inter-socket link.  With one process per GPU and unbound ranks the end-to-end rate of 8 ranks
    was 1.4x that of one.  Returns the CPU list, or None when the topology cannot be read (then
    nothing is changed)."""
    import os
    bus = pci_bus_id.lower()
    if len(bus.split(":")[0]) == 8:               # nvml style 00000000:1B:00.0 -> sysfs 0000:1b:00.0
        bus = bus[4:]
    try:
        with open(f"/sys/bus/pci/devices/{bus}/local_cpulist") as fh:
            cpus = _parse_cpulist(fh.read())
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if not cpus or cpus == allowed:
            return None
        os.sched_setaffinity(0, cpus)
        return sorted(cpus)
    except (OSError, ValueError, AttributeError):
        return None
