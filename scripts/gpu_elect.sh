L=gpurun_out/elect1.log; : > $L
timeout 300 ./tests/cuda/tc_selftest perf >> $L 2>&1; echo "exit=$?" >> $L
grep -v "^\[[chw][0-9].*ok" $L | tail -40
if grep -q "FAILED\|exit=[1-9]" $L; then echo "SELFTEST PROBLEM - skipping rest"; else
python -m pytest tests/test_kernels_gpu.py tests/test_biggan_gpu.py tests/test_tc_selftest.py -m gpu -x -q 2>&1 | tail -4
timeout 600 python -m pytest tests/test_knn.py -m gpu -x -q 2>&1 | tail -3
ICGAN_BENCH_SHAPES=1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_elect.json 2> gpurun_out/bench_elect.err; tail -1 gpurun_out/bench_elect.json | cut -c1-300
fi
