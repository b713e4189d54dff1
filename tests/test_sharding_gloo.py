"""N>1 host logic on CPU: two gloo ranks shard a batch, run their clips (the kernel
source on the CPU emulator stands in for the GPU), gather, and rank 0 checks the
assembled batch against the oracle."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from audioblocks import sharding


def test_shard_range_covers_batch_exactly():
    for total in (0, 1, 7, 8, 1024, 8191):
        for world in (1, 2, 3, 4, 8):
            spans = [sharding.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(4, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, n):
    here = os.path.dirname(os.path.abspath(__file__))
    for p in (os.path.dirname(here), here, os.path.join(os.path.dirname(here), "audio-effects-simulator_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import emu
    import synth
    from oracle import oracle as orc
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cfg = synth.PRESETS["Robot Voice"]
        lo, hi = sharding.shard_range(total, rank, world)
        x = synth.batch(lo, hi - lo, n)                       # each rank synthesises its own clips
        y = emu.run(emu.resolved_descs(cfg, 48000, n, 2), 48000, x, grid=2)
        full = sharding.gather_clips(torch.from_numpy(y), total, dst=0)
        every = sharding.gather_clips(torch.from_numpy(y), total, dst=None)
        q = (np.clip(y, -1, 1) * 32767).astype(np.int16)
        qfull = sharding.gather_clips(torch.from_numpy(q), total, dst=None)      # int16 PCM travels as bytes
        assert qfull.dtype == torch.int16 and torch.equal(qfull[lo:hi], torch.from_numpy(q))
        if total % world == 0:                                # equal shards: the single-collective path bench.py times
            qall = sharding.all_gather_pcm(torch.from_numpy(q))
            assert qall.shape == (total, n, 2) and torch.equal(qall, qfull)
            again = sharding.all_gather_pcm(torch.from_numpy(q), out=qall)
            assert again is qall and torch.equal(again, qfull)
        slow = sharding.max_over_ranks(1.0 + rank)
        assert slow == float(world)
        assert every.shape == (total, n, 2)
        if rank == 0:
            assert torch.equal(full, every)
            for b in range(total):
                want = orc.run_file_path(cfg, synth.clip(b, n, 2), 48000)
                mx, snr = synth.err_stats(full[b].numpy(), want)
                assert mx <= 1e-5 and snr >= 100.0, (b, mx, snr)
        else:
            assert full is None
    finally:
        dist.destroy_process_group()


def test_two_rank_shard_process_gather_gloo():
    mp.spawn(_worker, args=(2, _free_port(), 5, 3000), nprocs=2, join=True)


def test_two_rank_equal_shards_all_gather_into_tensor_gloo():
    mp.spawn(_worker, args=(2, _free_port(), 4, 2500), nprocs=2, join=True)
