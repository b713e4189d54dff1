"""Multi-GPU check (run under torchrun on a GPU box, not a pytest test):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tests/multi_gpu_check.py
Every rank processes its shard of a preset sweep on its own GPU; NCCL gathers the int16 results
on rank 0 (the only collective: results, never inside the compute), rank 0 checks them against the
CPU oracle."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE, os.path.join(ROOT, "audio-effects-simulator_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np
import torch
import torch.distributed as dist

import synth
from audioblocks import _native, sharding
from audioblocks.engine import file_chain


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    _native.check(_native.lib().aes_set_device(local))
    total, n = 13, 40000                                   # ragged over 2/4/8 ranks on purpose
    lo, hi = sharding.shard_range(total, rank, world)
    ok = True
    for name in ("Rain Delay", "Robot Voice", "Slapback Echo"):
        cfg = synth.PRESETS[name]
        x = synth.batch(lo, hi - lo, n)
        out = np.empty((hi - lo, n, 2), np.int16)
        file_chain(cfg, 48000, channels_in=2).process_batch(x, out)          # int16 halves the gather
        full = sharding.gather_clips(torch.from_numpy(out).cuda(), total, dst=0)
        if rank == 0:
            from oracle import oracle as orc
            got = full.cpu().numpy()
            for b in range(total):
                want = orc.quantize_i16(np.clip(orc.run_file_path(cfg, synth.clip(b, n, 2), 48000), -1, 1))
                worst = int(np.max(np.abs(got[b].astype(np.int32) - want.astype(np.int32))))
                if worst > 1:
                    ok = False
                    print(f"MISMATCH {name} clip {b}: {worst} LSB")
            print(f"{name}: {total} clips over {world} ranks gathered, max int16 deviation <= 1 LSB: {ok}")
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("MULTI_GPU_CHECK", "OK" if ok else "FAILED")
        sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
