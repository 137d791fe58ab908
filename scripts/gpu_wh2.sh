L=gpurun_out/wh2.log; : > $L
timeout 300 ./tests/cuda/tc_selftest perf 2>&1 | grep -A1 "^\[wp\|TC_SELFTEST" >> $L
echo "== old split rule" >> $L
ICGAN_TC_WGRAD_OLD_SPLITS=1 ICGAN_TC_WGRAD_HALO=0 timeout 300 ./tests/cuda/tc_selftest perf 2>&1 | grep -A1 "^\[wp" >> $L
cat $L
python scripts/bench_elementwise.py > gpurun_out/ew_bench3.log 2>&1; head -28 gpurun_out/ew_bench3.log
python -m pytest tests/test_tc_selftest.py tests/test_kernels_gpu.py tests/test_biggan_gpu.py -m gpu -x -q 2>&1 | tail -4
ICGAN_BENCH_SHAPES=1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_wh2.json 2> gpurun_out/bench_wh2.err; tail -1 gpurun_out/bench_wh2.json | cut -c1-300
