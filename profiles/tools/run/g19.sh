# r2ae: idle walker lanes shadow their own column, swizzled staged-input read in aes_fast_kernel, aligned vectors + select
# for unaligned staged delay lines
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python profiles/tools/time_presets.py 1184 Rain Robot Cathedral Guitar Slapback c3 delay octaver reverb 2>&1 | tail -14
timeout 300 python bench.py --no-e2e --no-cpu --no-sweep --no-gather --steps 10 --warmup 3 --clips 1184 2>/dev/null | python -c "import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('rain 1184',d['value'],d['roofline']['frac'])"
