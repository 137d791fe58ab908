# Profiling pass on one B200 (run through gpurun; results land in gpurun_out/, summaries are then written to profiles/):
#   1. `ncu --set full` of one warm launch of each tensor-core kernel family (source-level stalls included)
#   2. launch list (gpu__time_duration) of two half-size bench steps -> kernel shares of the step
T=./tests/cuda/tc_selftest
FULL="ncu --set full --clock-control none --import-source on"
$FULL -k regex:tc_conv_halo -s 3 -c 1 -o gpurun_out/prof_halo384 $T prof2 32 64 64 384 384 > /dev/null 2>&1
$FULL -k regex:tc_conv_halo -s 3 -c 1 -o gpurun_out/prof_halo96 $T prof2 8 256 256 96 96 > /dev/null 2>&1
$FULL -k regex:tc_wgrad_halo -s 1 -c 1 -o gpurun_out/prof_wgrad_halo $T prof > /dev/null 2>&1
$FULL -k regex:tc_conv_rgb -s 2 -c 1 -o gpurun_out/prof_rgb python scripts/prof_rgb.py 32 > /dev/null 2>&1
OUT=gpurun_out/launches.csv
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT \
    python bench.py --ncu --steps 1 --warmup 1 --per-gpu-batch 128 --micro-batch 128 > gpurun_out/ncu_bench.log 2>&1
python scripts/summarize_launches.py $OUT > gpurun_out/launches_summary.txt
head -20 gpurun_out/launches_summary.txt
ls -la gpurun_out/*.ncu-rep
# then, here: python scripts/ncu_summary.py gpurun_out/prof_*.ncu-rep; python scripts/ncu_stalls.py <rep> [N]
