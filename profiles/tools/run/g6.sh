set -x
python bench.py --preset "Clean Noise Removal" --no-e2e --no-cpu --no-sweep --no-gather --steps 3 --warmup 2 > gpurun_out/cnr.json 2> gpurun_out/cnr.err
tail -2 gpurun_out/cnr.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/cnr_launches.csv python bench.py --preset "Clean Noise Removal" --no-e2e --no-cpu --no-sweep --no-gather --steps 1 --warmup 1 > /dev/null 2>&1
