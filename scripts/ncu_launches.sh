# every launch of one short step with its device time (cold-cache, serialised: compare SHARES, not absolutes)
OUT=gpurun_out/launches_r01c.csv
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT \
    python bench.py --ncu --steps 1 --warmup 1 --per-gpu-batch 32 --micro-batch 32 > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log
python scripts/summarize_launches.py $OUT > gpurun_out/launches_r01c_summary.txt
head -40 gpurun_out/launches_r01c_summary.txt
