L=gpurun_out/rgb1.log; : > $L
echo "== PAD64=1" >> $L
ICGAN_TC_HALO_PAD64=1 timeout 300 ./tests/cuda/tc_selftest convperf 2>&1 | grep -A1 "^\[p[0-9]\|TC_SELFTEST\|FAIL\|^\[c3\|^\[c7\|^\[h2" >> $L
cat $L
python -m pytest tests/test_kernels_gpu.py tests/test_biggan_gpu.py -m gpu -x -q 2>&1 | tail -6
ICGAN_BENCH_SHAPES=1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_rgb.json 2> gpurun_out/bench_rgb.err; tail -1 gpurun_out/bench_rgb.json | cut -c1-300
