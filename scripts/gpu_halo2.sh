L=gpurun_out/halo2.log; : > $L
T=./tests/cuda/tc_selftest
for pad in 0 1; do
  echo "== PAD64=$pad" >> $L
  ICGAN_TC_HALO_PAD64=$pad timeout 60 $T prof2 8 256 256 96 96 >> $L 2>&1
  ICGAN_TC_HALO_PAD64=$pad timeout 60 $T prof2 8 128 128 96 192 >> $L 2>&1
done
cat $L
# full-set captures of the halo kernel: 384@64 (BN=128) and 96@256 (BN=96, 64B rows)
ncu --set full --clock-control none --import-source on -k regex:tc_conv_halo -s 3 -c 1 -o gpurun_out/prof_halo384_r01 $T prof2 32 64 64 384 384 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:tc_conv_halo -s 3 -c 1 -o gpurun_out/prof_halo96_r01 $T prof2 8 256 256 96 96 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:tc_wgrad -s 3 -c 1 -o gpurun_out/prof_wgrad_late_r01 $T prof > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
# launch list of one step at the bench micro-batch
OUT=gpurun_out/launches_r01d.csv
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT \
    python bench.py --ncu --steps 1 --warmup 1 --per-gpu-batch 128 --micro-batch 128 > gpurun_out/ncu_bench_d.log 2>&1
tail -2 gpurun_out/ncu_bench_d.log | cut -c1-300
python scripts/summarize_launches.py $OUT > gpurun_out/launches_r01d_summary.txt
head -30 gpurun_out/launches_r01d_summary.txt
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_halo_m128.json 2> gpurun_out/bench_halo_m128.err; tail -1 gpurun_out/bench_halo_m128.json | cut -c1-400
