# conv selftest + timings: old per-tap kernel vs halo-reuse kernel variants
L=gpurun_out/halo1.log; : > $L
run() { echo "== $*" >> $L; env "$@" timeout 150 ./tests/cuda/tc_selftest convperf >> $L 2>&1; echo "exit=$?" >> $L; }
run ICGAN_TC_HALO=1
run ICGAN_TC_HALO=0
run ICGAN_TC_HALO=1 ICGAN_TC_HALO_CW=32
run ICGAN_TC_HALO=1 ICGAN_TC_HALO_CW=32 ICGAN_TC_HALO_STAGE_KB=64
grep -v "^\[c[0-9]\|^\[h[0-9].*ok" $L | tail -80
if grep -q "FAILED\|exit=[1-9]" $L; then echo "SELFTEST PROBLEM - skipping bench"; else
python -m pytest tests/test_stylegan_ops.py tests/test_stylegan_conv.py tests/test_biggan_gpu.py -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_halo.json 2> gpurun_out/bench_halo.err; tail -3 gpurun_out/bench_halo.json | cut -c1-900
fi
