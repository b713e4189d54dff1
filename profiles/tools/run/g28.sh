# r2am: the file route and live blocks at the round's final revision (1 GPU)
timeout 300 python profiles/tools/file_route_timing.py > gpurun_out/file_route_r2am.json 2> gpurun_out/file_route_r2am.err; tail -2 gpurun_out/file_route_r2am.err; cat gpurun_out/file_route_r2am.json | tail -1 | cut -c1-600
timeout 300 python profiles/tools/stream_block_timing.py > gpurun_out/stream_block_r2am.txt 2>&1; tail -8 gpurun_out/stream_block_r2am.txt
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:aes -c 40 --csv --log-file gpurun_out/file_route_launches_r2am.csv python profiles/tools/file_route_timing.py > /dev/null 2>&1
python profiles/tools/launch_summary.py gpurun_out/file_route_launches_r2am.csv
