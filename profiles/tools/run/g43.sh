# r2ay: eight ranks at the round's final revision: the default bench (strong scaling of the 8192-clip job, gather leg)
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_default_8gpu_r2ay.json 2> gpurun_out/bench_default_8gpu_r2ay.err
tail -2 gpurun_out/bench_default_8gpu_r2ay.err
python -c "
import json;d=json.loads(open('gpurun_out/bench_default_8gpu_r2ay.json').read().strip().splitlines()[-1]);print(d['value'],d['n_gpus'],d['roofline']['frac'],d['parity'],d['e2e']['value'],d.get('gather'))"
