#!/bin/bash
# how much of the step is per-launch event overhead?
for ev in 1 0; do
  LE_BENCH_EVENTS=$ev python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('events=$ev fps', round(d['value']), 'ms/step', round(d['ms_per_step'],4), 'step_frac', round(d['roofline']['step_frac_of_hbm_roofline'],3))"
done
