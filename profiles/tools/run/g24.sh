# r2ai: rows10 variants (G pairs per CTA, CTAs per SM, exchange buffers per row); K1 / K3 back to their r2ad loops
timeout 600 python -m pytest tests -m gpu -x -q -k "spectral or Clean or noise or golden or preset" 2>&1 | tail -3
for v in 0 1 2 5 6 7; do
  echo "variant $v"; AES_ROWS10_VARIANT=$v CHUNKS_MB=4096 timeout 200 python profiles/tools/time_spectral.py 2048 2>&1 | grep "smooth chunk"
done
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:aes -c 60 --csv --log-file gpurun_out/spectral_launches_r2ai.csv python bench.py --preset "Clean Noise Removal" --total-clips 2048 --no-e2e --no-cpu --no-sweep --no-gather --steps 2 --warmup 3 > /dev/null 2>&1
python profiles/tools/launch_summary.py gpurun_out/spectral_launches_r2ai.csv
