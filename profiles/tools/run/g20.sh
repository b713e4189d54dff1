# r2af: register-resident row-pair kernel of the four-step SpectralFilter (aesm_rows10_body): parity, variants, launch list
timeout 600 python -m pytest tests -m gpu -x -q -k "spectral or Clean or noise or golden or preset" 2>&1 | tail -3
for v in 0 1 2 3 4; do
  echo "variant $v"; AES_ROWS10_VARIANT=$v CHUNKS_MB=4096 timeout 200 python profiles/tools/time_spectral.py 2048 2>&1 | grep "smooth chunk"
done
echo "shared-memory rows kernel"; AES_SPECTRAL_ROWS_SMEM=1 CHUNKS_MB=4096 timeout 200 python profiles/tools/time_spectral.py 2048 2>&1 | grep "smooth chunk"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:aes -c 60 --csv --log-file gpurun_out/spectral_launches_rows10.csv python bench.py --preset "Clean Noise Removal" --total-clips 2048 --no-e2e --no-cpu --no-sweep --no-gather --steps 2 --warmup 3 > /dev/null 2>&1
python profiles/tools/launch_summary.py gpurun_out/spectral_launches_rows10.csv
timeout 300 python bench.py --preset "Clean Noise Removal" --no-e2e --no-sweep --no-gather --steps 3 --warmup 3 2>gpurun_out/cnr.err | python -c "import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('CNR 8192',d['value'],d['roofline']['frac'],d['parity'])"
tail -2 gpurun_out/cnr.err
