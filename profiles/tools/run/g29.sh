# r2an: two ranks at the final revision: the default bench (strong scaling, gather leg) and the multi-GPU parity check
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_default_2gpu_r2an.json 2> gpurun_out/bench_default_2gpu_r2an.err
tail -2 gpurun_out/bench_default_2gpu_r2an.err
python -c "
import json;d=json.loads(open('gpurun_out/bench_default_2gpu_r2an.json').read().strip().splitlines()[-1]);print(d['value'],d['n_gpus'],d['roofline']['frac'],d['parity'],d['e2e']['value'],d.get('gather'))"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 tests/multi_gpu_check.py 2>&1 | tail -4
