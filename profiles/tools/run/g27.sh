# r2al: the shipped rows10 variant (<2 pairs, 2 CTAs, 1 buffer, bulk-copy prefetch>), new full-size spectral test, full suite, bench
( time timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 ) 2>&1 | grep -v "^$\|user\|sys"
AES_SPECTRAL_ROWS_SMEM=1 timeout 300 python -m pytest tests/test_gpu_full_size.py -m gpu -x -q -k spectral 2>&1 | tail -1
( time timeout 600 python bench.py > gpurun_out/bench_default_r2al.json 2> gpurun_out/bench_default_r2al.err ) 2>&1 | grep real
tail -3 gpurun_out/bench_default_r2al.err
python -c "
import json;d=json.loads(open('gpurun_out/bench_default_r2al.json').read().strip().splitlines()[-1]);print(d['value'],d['roofline']['frac'],d['parity'],d['e2e']['value'], d['clocks']);print({k:(round(v['ms'],2),round(v['frac_of_hbm_roofline'],3)) for k,v in d['sweep']['presets'].items()}, d['sweep']['one_at_a_time'], d['sweep']['six_streams'])"
