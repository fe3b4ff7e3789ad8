#!/bin/bash
python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/plain_n.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"$1" -s ${2:-3} -c ${3:-1} -o gpurun_out/$4 -f python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_n.log 2>&1; tail -1 gpurun_out/ncu_n.log
