set -x
ncu --set full --import-source on --clock-control none -k regex:aes_fast_kernel -s 4 -c 1 -o gpurun_out/c2batch python bench.py --preset c2-biquad-cascade --clips 1184 --no-e2e --no-cpu --no-sweep --no-gather --steps 3 --warmup 3 > gpurun_out/c2batch_ncu.log 2>&1
tail -2 gpurun_out/c2batch_ncu.log
