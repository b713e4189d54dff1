# r2at: aes_fast_kernel at 3 CTAs per SM (80 registers, -DAESF_MIN_CTAS=3 build beside the shipped one)
echo "shipped (2 CTAs per SM, <= 128 registers)"
timeout 600 python profiles/tools/time_presets.py 1184 Robot Slapback c3 delay octaver 2>&1 | tail -7
echo "3 CTAs per SM"
AESIM_LIB=$PWD/audio-effects-simulator_b200/lib/libaesim_ctas3.so timeout 600 python profiles/tools/time_presets.py 1184 Robot Slapback c3 delay octaver 2>&1 | tail -7
