#!/bin/bash
# last check of a session: every GPU test, smoke(), the default bench line, the ncu launch list of the bench command
S=${1:-r1_s5b}
python -m pytest tests -m gpu -x -q --timeout 900 > gpurun_out/t_all.log 2>&1; tail -2 gpurun_out/t_all.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log | cut -c1-200
python bench.py > gpurun_out/bench_full.log 2>gpurun_out/bench_full.err; tail -1 gpurun_out/bench_full.log | cut -c1-1700
python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/plain_l.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${S}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_l.log 2>&1
tail -1 gpurun_out/ncu_l.log | cut -c1-120
