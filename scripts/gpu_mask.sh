python -m pytest tests/test_kernels_gpu.py tests/test_biggan_gpu.py -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_mask.json 2> gpurun_out/bench_mask.err; tail -1 gpurun_out/bench_mask.json | cut -c1-260
