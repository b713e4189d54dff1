timeout 300 python -m pytest tests -m gpu -x -q -k "not conv and not analysis" 2>&1 | tail -3
for args in "--preset c2-biquad-cascade --clips 1184" "--preset c2-biquad-cascade --total-clips 8192" "--preset c2-biquad-cascade --clips 30000 --seconds 2" "--preset c2-biquad-cascade --clips 2368"; do
  eval timeout 200 python bench.py $args --no-e2e --no-sweep --no-gather --steps 10 --warmup 3 2>> gpurun_out/bqseq.err | python -c "import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print(d['config'].get('preset'),d['config'].get('clips_this_rank'),d['value'],d['roofline']['frac'],d['parity']['max_abs_err'])"
done
