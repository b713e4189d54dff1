( time timeout 600 python bench.py > gpurun_out/bench_default_r2ac.json 2> gpurun_out/bench_default_r2ac.err ) 2>&1 | grep real
tail -3 gpurun_out/bench_default_r2ac.err
python -c "
import json;d=json.loads(open('gpurun_out/bench_default_r2ac.json').read().strip().splitlines()[-1]);print(d['value'],d['roofline']['frac'],d['parity'],d['pipelined_steps'])"
