set -x
python bench.py --preset c4-convreverb --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/c4_before.json 2> gpurun_out/c4_before.err
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:aesc_ --csv --log-file gpurun_out/c4_before_launches.csv python bench.py --preset c4-convreverb --steps 2 --warmup 1 --no-e2e --no-cpu > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:aes_rv_kernel -s 4 -c 1 -o gpurun_out/rv_r2v python bench.py --no-e2e --no-cpu --no-sweep --no-gather --steps 3 --warmup 3 --clips 1184 > gpurun_out/rv_ncu.log 2>&1
ls -la gpurun_out
