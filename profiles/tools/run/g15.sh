( time timeout 600 python bench.py > gpurun_out/bench_default_r2ab.json 2> gpurun_out/bench_default_r2ab.err ) 2>&1 | grep real
tail -3 gpurun_out/bench_default_r2ab.err
( time timeout 600 python bench.py --impl reference > gpurun_out/bench_reference_r2ab.json 2> gpurun_out/bench_reference_r2ab.err ) 2>&1 | grep real
cat gpurun_out/bench_reference_r2ab.json | head -c 600
