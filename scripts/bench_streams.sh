#!/bin/bash
for s in 1 2 4; do python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --streams $s 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('streams', d['config']['streams'], 'fps', round(d['value']), 'ms/step', round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['roofline']['kernel_ms_per_launch'].items()}, d['clocks'])"; done
