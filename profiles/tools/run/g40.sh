# r2av: the gate shape at 3 CTAs per SM (Clean Noise Removal's second stage)
CHUNKS_MB=4096 timeout 200 python profiles/tools/time_spectral.py 2048 2>&1 | grep "smooth chunk"
CHUNKS_MB=4096 timeout 200 python profiles/tools/time_spectral.py 8192 2>&1 | grep "smooth chunk"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:aes_fast -c 10 --csv --log-file gpurun_out/gate_launches.csv python bench.py --preset "Clean Noise Removal" --total-clips 2048 --no-e2e --no-cpu --no-sweep --no-gather --steps 2 --warmup 3 > /dev/null 2>&1
python profiles/tools/launch_summary.py gpurun_out/gate_launches.csv
timeout 600 python -m pytest tests -m gpu -x -q -k "gate or Clean or noise or golden or preset" 2>&1 | tail -1
