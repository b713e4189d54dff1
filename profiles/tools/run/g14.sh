timeout 300 python -m pytest tests -m gpu -x -q -k "not conv and not analysis" 2>&1 | tail -3
timeout 300 python profiles/tools/time_chains.py 2368 2>&1 | grep -i "guitar"
timeout 300 python profiles/tools/time_chains.py 8192 2>&1 | grep -i "guitar"
timeout 200 python bench.py --preset "Guitar Filter" --no-e2e --no-sweep --no-gather --steps 10 --warmup 3 2>> gpurun_out/gf.err | python -c "import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print(d['config'].get('preset'),d['config'].get('clips_this_rank'),d['value'],d['roofline']['frac'],d['parity'])"
