python scripts/prof_rgb.py 32 2>&1 | tail -2
python -m pytest tests/test_tc_selftest.py tests/test_kernels_gpu.py tests/test_biggan_gpu.py -m gpu -x -q 2>&1 | tail -4
ICGAN_BENCH_SHAPES=1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_rgb2.json 2> gpurun_out/bench_rgb2.err; tail -1 gpurun_out/bench_rgb2.json | cut -c1-300
