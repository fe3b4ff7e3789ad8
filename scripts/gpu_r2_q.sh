#!/bin/bash
# round 2, call Q: e2e with two batches in flight (run(wait=False)), all four workloads
mkdir -p gpurun_out
for w in canny_hough_1080p_b256 canny_houghp_4k_b128 lsd_1080p_b512 lines_1080p_cave_b128; do
  timeout 400 python bench.py --steps 5 --warmup 3 --workload $w --no-cpu 2>gpurun_out/q_$w.err | tee gpurun_out/q_$w.json | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']; print(sys.argv[1], 'fps', round(d['value']), 'e2e', round(e['value']), 'ceiling', round(e['copy_ceiling_fps']), round(e['value']/e['copy_ceiling_fps'],3), 'parity', d['parity_sample']['equal'])" $w
  tail -c 300 gpurun_out/q_$w.err
done
