set -x
( time python bench.py > gpurun_out/bench_default_r2z2.json 2> gpurun_out/bench_default_r2z2.err ) 2>&1
tail -5 gpurun_out/bench_default_r2z2.err
python -c "
import json;d=json.loads(open('gpurun_out/bench_default_r2z2.json').read().strip().splitlines()[-1]);print(json.dumps(d.get('baseline_configs'),indent=1))"
