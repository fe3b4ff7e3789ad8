#!/bin/bash
# end-of-session measurement: full GPU test suite, default bench, launch list under ncu
python -m pytest tests -m gpu -x -q --timeout 1800 > gpurun_out/t_all.log 2>&1; tail -3 gpurun_out/t_all.log
python bench.py > gpurun_out/bench_full.log 2>gpurun_out/bench_full.err; tail -1 gpurun_out/bench_full.log | cut -c1-600
LE_BENCH_STREAMS=1 python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/plain_l.log 2>&1 && \
LE_BENCH_STREAMS=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_s2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_l.log 2>&1
tail -2 gpurun_out/ncu_l.log | cut -c1-200
