set -x
python -m pytest tests -m gpu -x -q -k "conv" 2>&1 | tail -5
python bench.py --preset c4-convreverb --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/c4_r2y.json 2> gpurun_out/c4_r2y.err
tail -3 gpurun_out/c4_r2y.err
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:aesc_ --csv --log-file gpurun_out/c4_r2y_launches.csv python bench.py --preset c4-convreverb --steps 2 --warmup 1 --no-e2e --no-cpu > /dev/null 2>&1
