ncu --set full --import-source on --clock-control none -k regex:aesm_rows10 -s 6 -c 1 -o gpurun_out/rows10 python bench.py --preset "Clean Noise Removal" --total-clips 2048 --no-e2e --no-cpu --no-sweep --no-gather --steps 2 --warmup 3 > gpurun_out/rows10_ncu.log 2>&1
tail -1 gpurun_out/rows10_ncu.log | cut -c1-300
