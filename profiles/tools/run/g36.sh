# r2as: biquad batch kernel: step loop unrolled 4x against 2x (both at launch bounds 64 x 4)
for w in "" 1; do
  echo "AES_BQSEQ_UNROLL4=$w"
  for args in "--total-clips 8192" "--clips 2368" "--clips 1184"; do
    if [ -n "$w" ]; then export AES_BQSEQ_UNROLL4=1; else unset AES_BQSEQ_UNROLL4; fi
    timeout 200 python bench.py --preset c2-biquad-cascade $args --no-e2e --no-sweep --no-gather --steps 10 --warmup 3 2>> gpurun_out/bqseq.err | python -c "import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print(d['config'].get('clips_this_rank'),d['value'],d['roofline']['frac'],d['parity']['max_abs_err'])"
  done
  timeout 300 python profiles/tools/time_chains.py 8192 2>&1 | grep "default"
done
