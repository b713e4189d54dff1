set -x
python -m pytest tests -m gpu -x -q -k "not conv and not analysis" 2>&1 | tail -4
for args in "--preset c2-biquad-cascade --clips 1184" "--preset c2-biquad-cascade --total-clips 8192" "--preset c2-biquad-cascade --clips 296" "--preset Guitar\ Filter --clips 1184"; do
  eval python bench.py $args --no-e2e --no-sweep --no-gather --steps 10 --warmup 3 2>> gpurun_out/bqseq.err | python -c "import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print(d['config'].get('preset'),d['config'].get('clips_this_rank'),d['value'],d['roofline']['frac'],d['parity'])"
done
tail -3 gpurun_out/bqseq.err
AES_NO_BQSEQ=1 python bench.py --preset c2-biquad-cascade --clips 1184 --no-e2e --no-cpu --no-sweep --no-gather --steps 10 --warmup 3 2>/dev/null | python -c "import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('old path',d['value'],d['roofline']['frac'])"
