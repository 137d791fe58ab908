T=./tests/cuda/tc_selftest
ncu --set full --clock-control none --import-source on -k regex:tc_wgrad_halo -s 1 -c 1 -o gpurun_out/prof_wgrad_halo_r01 $T prof > /dev/null 2>&1
OUT=gpurun_out/launches_r01_final.csv
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT \
    python bench.py --ncu --steps 1 --warmup 1 --per-gpu-batch 128 --micro-batch 128 > gpurun_out/ncu_bench_final.log 2>&1
tail -1 gpurun_out/ncu_bench_final.log | cut -c1-200
python scripts/summarize_launches.py $OUT > gpurun_out/launches_r01_final_summary.txt
head -12 gpurun_out/launches_r01_final_summary.txt
ls -la gpurun_out/prof_wgrad_halo_r01.ncu-rep
