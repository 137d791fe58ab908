python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -2
python -m pytest tests/ -x -q -m gpu 2>&1 | tail -4
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -1 gpurun_out/bench_final.json | cut -c1-250; wc -l gpurun_out/bench_final.json
