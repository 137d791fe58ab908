# one full-set capture per tensor-core kernel (launch #2 of each = warm), source-level stalls included
ncu --set full --clock-control none --import-source on -k regex:tc_wgrad_kernel -s 1 -c 1 -o gpurun_out/prof_wgrad_r01 ./tests/cuda/tc_selftest prof > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:tc_conv_kernel -s 1 -c 1 -o gpurun_out/prof_conv_r01 ./tests/cuda/tc_selftest prof > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:tc_wgrad_kernel -s 4 -c 1 -o gpurun_out/prof_wgrad96_r01 ./tests/cuda/tc_selftest prof > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:tc_conv_kernel -s 4 -c 1 -o gpurun_out/prof_conv96_r01 ./tests/cuda/tc_selftest prof > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
