# r2aw: inverse column kernel with launch bounds (256, 3)
CHUNKS_MB=4096 timeout 200 python profiles/tools/time_spectral.py 2048 2>&1 | grep "smooth chunk"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:aesm -c 40 --csv --log-file gpurun_out/k3_launches.csv python bench.py --preset "Clean Noise Removal" --total-clips 2048 --no-e2e --no-cpu --no-sweep --no-gather --steps 2 --warmup 3 > /dev/null 2>&1
python profiles/tools/launch_summary.py gpurun_out/k3_launches.csv
