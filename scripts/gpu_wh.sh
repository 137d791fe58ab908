L=gpurun_out/wh1.log; : > $L
echo "== wgrad halo ON" >> $L
timeout 300 ./tests/cuda/tc_selftest perf >> $L 2>&1; echo "exit=$?" >> $L
echo "== wgrad halo OFF (timings only)" >> $L
ICGAN_TC_WGRAD_HALO=0 timeout 300 ./tests/cuda/tc_selftest perf 2>&1 | grep -A1 "^\[wp" >> $L
grep -v "^\[[chw][0-9]*_.*ok\|^\[p[0-9]" $L | grep -v "^    time.*TFLOP" | tail -30
grep -A1 "^\[wp\|== " $L | tail -40
python scripts/bench_elementwise.py > gpurun_out/ew_bench2.log 2>&1; head -28 gpurun_out/ew_bench2.log
if grep -q "FAILED\|exit=[1-9]" $L; then echo "SELFTEST PROBLEM - skipping rest"; else
python -m pytest tests/test_kernels_gpu.py tests/test_biggan_gpu.py -m gpu -x -q 2>&1 | tail -4
ICGAN_BENCH_SHAPES=1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_wh.json 2> gpurun_out/bench_wh.err; tail -1 gpurun_out/bench_wh.json | cut -c1-300
fi
