# Round-end style validation on one B200 (run through gpurun): smoke, the whole GPU suite, both bench arms.
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -2
python -m pytest tests/ -q -m gpu 2>&1 | tail -6
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -1 gpurun_out/bench.json | cut -c1-250; wc -l gpurun_out/bench.json
python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -1 gpurun_out/bench_ref.json | cut -c1-400
