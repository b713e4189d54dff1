set -x
python -m pytest tests -m gpu -x -q -k "not conv and not analysis" 2>&1 | tail -4
python bench.py --no-e2e --no-cpu --no-sweep --no-gather --steps 20 --warmup 3 > gpurun_out/rv_r2aa_8192.json 2> gpurun_out/rv_r2aa.err
python bench.py --no-e2e --no-cpu --no-sweep --no-gather --steps 20 --warmup 3 --clips 1184 > gpurun_out/rv_r2aa_1184.json 2>> gpurun_out/rv_r2aa.err
for p in "Cathedral" "Guitar Filter"; do python bench.py --preset "$p" --no-e2e --no-cpu --no-sweep --no-gather --steps 10 --warmup 3 --clips 1184 2>> gpurun_out/rv_r2aa.err | python -c "import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print(d['config']['preset'],d['value'],d['roofline']['frac'],d['roofline']['kernel'])"; done
tail -3 gpurun_out/rv_r2aa.err
python -c "
import json
for f in ('8192','1184'):
    d=json.loads(open('gpurun_out/rv_r2aa_%s.json'%f).read().strip().splitlines()[-1]);print(f,d['value'],d['roofline']['frac'],d['roofline']['kernel'])"
ncu --set full --import-source on --clock-control none -k regex:aes_rv_kernel -s 4 -c 1 -o gpurun_out/rv_r2aa python bench.py --no-e2e --no-cpu --no-sweep --no-gather --steps 3 --warmup 3 --clips 1184 > gpurun_out/rv_ncu.log 2>&1
