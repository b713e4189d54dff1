set -x
python -m pytest tests/test_gpu_analysis.py -m gpu -x -q 2>&1 | tail -8
python profiles/tools/time_analysis.py 4096 > gpurun_out/analysis_timing.json 2>&1
cat gpurun_out/analysis_timing.json
