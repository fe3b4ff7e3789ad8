#!/bin/bash
# stripe height of the Hough vote kernel: kernel split per LE_VOTE_A
for a in 16 32 8; do
  LE_VOTE_A=$a python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('A=$a fps', round(d['value']), 'ms/step', round(d['ms_per_step'],4), {k: round(v,3) for k,v in d['roofline']['kernel_ms_per_launch'].items()})"
done
