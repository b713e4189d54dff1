ncu --set full --import-source on --clock-control none -k regex:aes_biquad_seq -s 4 -c 1 -o gpurun_out/bqseq python bench.py --preset c2-biquad-cascade --total-clips 8192 --no-e2e --no-cpu --no-sweep --no-gather --steps 3 --warmup 3 > gpurun_out/bqseq_ncu.log 2>&1
tail -2 gpurun_out/bqseq_ncu.log
