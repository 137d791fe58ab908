L=gpurun_out/selftest3.log; : > $L
for i in 6 0 1 2 3 4 5 7 8 9; do echo "== wgrad $i" >> $L; timeout 60 ./tests/cuda/tc_selftest wgrad $i >> $L 2>&1; echo "exit=$?" >> $L; done
echo "== perf" >> $L
timeout 120 ./tests/cuda/tc_selftest perf 2>&1 | grep -v "^\[c[0-9]" >> $L
tail -60 $L
