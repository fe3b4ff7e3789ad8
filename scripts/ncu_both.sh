#!/bin/bash
# one --set full capture of each dominant kernel (single stream: whole batch per launch)
export LE_BENCH_STREAMS=1
python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/plain_n.log 2>&1 || exit 1
for k in k_front_rows k_hough_vote; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -o gpurun_out/r1_final_$k -f python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_$k.log 2>&1
done
ls -la gpurun_out/r1_final_*
