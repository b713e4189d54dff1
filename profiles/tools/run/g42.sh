# r2ax: the round's final revision once more: full GPU suite, smoke, default bench, reference arm
( time timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 ) 2>&1 | grep -v "^$\|user\|sys"
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
( time timeout 600 python bench.py > gpurun_out/bench_default_r2ax.json 2> gpurun_out/bench_default_r2ax.err ) 2>&1 | grep real
tail -3 gpurun_out/bench_default_r2ax.err
( time timeout 600 python bench.py --impl reference > gpurun_out/bench_reference_r2ax.json 2> gpurun_out/bench_reference_r2ax.err ) 2>&1 | grep real
python -c "
import json;d=json.loads(open('gpurun_out/bench_default_r2ax.json').read().strip().splitlines()[-1]);print(d['value'],d['roofline']['frac'],d['parity'],d['e2e']['value'], d['clocks']);print({k:(round(v['ms'],2),round(v['frac_of_hbm_roofline'],3)) for k,v in d['sweep']['presets'].items()}, d['sweep']['one_at_a_time'], d['sweep']['six_streams'])
r=json.loads(open('gpurun_out/bench_reference_r2ax.json').read().strip().splitlines()[-1]);print('reference', r['value'], r['cpu_baseline']['kind'], r['cpu_baseline']['cores'])"
timeout 300 python profiles/tools/time_presets.py 8192 Robot Cathedral Guitar 2>&1 | tail -3
