# the headline kernel once more with the source page, for per-instruction shared-memory wavefronts
ncu --set full --import-source on --clock-control none -k regex:aes_rv_kernel -s 4 -c 1 -o gpurun_out/rv_head python bench.py --no-e2e --no-cpu --no-sweep --no-gather --steps 3 --warmup 3 --clips 1184 > gpurun_out/rv_head_ncu.log 2>&1
tail -1 gpurun_out/rv_head_ncu.log | cut -c1-300
ncu --set full --import-source on --clock-control none -k regex:aes_fast_kernel -s 4 -c 1 -o gpurun_out/fast_robot python bench.py --preset "Robot Voice" --no-e2e --no-cpu --no-sweep --no-gather --steps 3 --warmup 3 --clips 1184 > gpurun_out/fast_robot_ncu.log 2>&1
tail -1 gpurun_out/fast_robot_ncu.log | cut -c1-300
