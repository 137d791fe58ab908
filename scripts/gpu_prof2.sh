python scripts/prof_rgb.py 32 2>&1 | tail -4
ncu --set full --clock-control none --import-source on -k regex:tc_conv_rgb -s 2 -c 1 -o gpurun_out/prof_rgb_r01 python scripts/prof_rgb.py 32 > /dev/null 2>&1
T=./tests/cuda/tc_selftest
ncu --set full --clock-control none --import-source on -k regex:tc_conv_halo -s 3 -c 1 -o gpurun_out/prof_halo384_late_r01 $T prof2 32 64 64 384 384 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:tc_conv_halo -s 3 -c 1 -o gpurun_out/prof_halo96_late_r01 $T prof2 8 256 256 96 96 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep | tail -4
