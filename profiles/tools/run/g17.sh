# HEAD validation: the full GPU suite, ncu of the biquad batch kernel's final revision and of the SpectralFilter
# four-step kernels, then the default bench line
( time timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 ) 2>&1 | grep -v "^$\|user\|sys"
ncu --set full --import-source on --clock-control none -k regex:aes_biquad_seq -s 4 -c 1 -o gpurun_out/bqseq_final python bench.py --preset c2-biquad-cascade --total-clips 8192 --no-e2e --no-cpu --no-sweep --no-gather --steps 3 --warmup 3 > gpurun_out/bqseq_ncu.log 2>&1
tail -1 gpurun_out/bqseq_ncu.log | cut -c1-300
ncu --set full --import-source on --clock-control none -k regex:aesm_ -s 6 -c 3 -o gpurun_out/spectral_smooth python bench.py --preset "Clean Noise Removal" --total-clips 2048 --no-e2e --no-cpu --no-sweep --no-gather --steps 2 --warmup 3 > gpurun_out/spectral_ncu.log 2>&1
tail -1 gpurun_out/spectral_ncu.log | cut -c1-300
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:aes -c 60 --csv --log-file gpurun_out/spectral_launches.csv python bench.py --preset "Clean Noise Removal" --total-clips 2048 --no-e2e --no-cpu --no-sweep --no-gather --steps 2 --warmup 3 > /dev/null 2>&1
( time timeout 600 python bench.py > gpurun_out/bench_default_r2ad.json 2> gpurun_out/bench_default_r2ad.err ) 2>&1 | grep real
tail -3 gpurun_out/bench_default_r2ad.err
python -c "
import json;d=json.loads(open('gpurun_out/bench_default_r2ad.json').read().strip().splitlines()[-1]);print(d['value'],d['roofline']['frac'],d['parity'],d['e2e']['value']);print(json.dumps(d.get('baseline_configs'))[:1500]);print(json.dumps(d.get('sweep'))[:1500])"
