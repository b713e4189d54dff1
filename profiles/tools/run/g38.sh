timeout 600 python profiles/tools/time_presets.py 1184 Robot gate 2>&1 | tail -3
AESIM_LIB=$PWD/audio-effects-simulator_b200/lib/libaesim_ctas3.so timeout 600 python profiles/tools/time_presets.py 1184 Robot gate 2>&1 | tail -3
timeout 600 python profiles/tools/time_presets.py 8192 Robot 2>&1 | tail -1
AESIM_LIB=$PWD/audio-effects-simulator_b200/lib/libaesim_ctas3.so timeout 600 python profiles/tools/time_presets.py 8192 Robot 2>&1 | tail -1
