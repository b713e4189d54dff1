# r2az: the clips of a partial last round go to one CTA per SM first (probe of the CTA -> SM placement)
timeout 300 python bench.py --no-e2e --no-cpu --no-sweep --no-gather --steps 10 --warmup 3 --clips 1024 2>/dev/null | python -c "import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('rain 1024 clips, tail order',d['value'],d['ms_per_step'],d['roofline']['frac'])"
AES_NO_TAIL_ORDER=1 timeout 300 python bench.py --no-e2e --no-cpu --no-sweep --no-gather --steps 10 --warmup 3 --clips 1024 2>/dev/null | python -c "import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('rain 1024 clips, CTAs 0..R-1 ',d['value'],d['ms_per_step'],d['roofline']['frac'])"
( time timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 ) 2>&1 | grep -v "^$\|user\|sys"
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
( time timeout 600 python bench.py > gpurun_out/bench_default_r2az.json 2> gpurun_out/bench_default_r2az.err ) 2>&1 | grep real
tail -2 gpurun_out/bench_default_r2az.err
python -c "
import json;d=json.loads(open('gpurun_out/bench_default_r2az.json').read().strip().splitlines()[-1]);print(d['value'],d['roofline']['frac'],d['parity'],d['e2e']['value']);print({k:(round(v['ms'],2),round(v['frac_of_hbm_roofline'],3)) for k,v in d['sweep']['presets'].items()}, d['sweep']['one_at_a_time']['ms'], d['sweep']['six_streams']['ms']);print({k[:40]:(v.get('value'),v.get('parity',{}).get('max_abs_err')) for k,v in d['baseline_configs'].items()})"
