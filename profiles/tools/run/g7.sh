set -x
( time python -m pytest tests -m gpu -x -q 2>&1 | tail -6 ) 2>&1
( time python bench.py > gpurun_out/bench_default_r2z.json 2> gpurun_out/bench_default_r2z.err ) 2>&1
tail -3 gpurun_out/bench_default_r2z.err
python -c "
import json;d=json.loads(open('gpurun_out/bench_default_r2z.json').read().strip().splitlines()[-1]);print(d['value'],d['roofline']['frac'],d['e2e']['value'],d['cpu_baseline'])"
