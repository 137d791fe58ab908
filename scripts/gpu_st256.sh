L=gpurun_out/st256.log; : > $L
timeout 300 ./tests/cuda/tc_selftest convperf 2>&1 | grep -A1 "^\[p[0-9]\|TC_SELFTEST\|FAIL" >> $L
cat $L
python scripts/bench_elementwise.py > gpurun_out/ew_bench4.log 2>&1; head -15 gpurun_out/ew_bench4.log
python -m pytest tests/test_tc_selftest.py tests/test_kernels_gpu.py tests/test_biggan_gpu.py -m gpu -x -q 2>&1 | tail -4
ICGAN_BENCH_SHAPES=1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_st256.json 2> gpurun_out/bench_st256.err; tail -1 gpurun_out/bench_st256.json | cut -c1-300
