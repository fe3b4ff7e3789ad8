#!/bin/bash
run() { echo "$@"; env "$@" python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --streams 1 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('  ms/step', round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['roofline']['kernel_ms_per_launch'].items()})"; }
python -m pytest tests/test_gpu_front.py -q -x 2>&1 | tail -2
run LE_FRONT_ROWS=120
run LE_FRONT_ROWS=160
run LE_FRONT_ROWS=200
run LE_FRONT_ROWS=120 LE_FRONT_LEVELS=3
run LE_FRONT_ROWS=160 LE_FRONT_LEVELS=3
run LE_FRONT_ROWS=90 LE_FRONT_LEVELS=3
run LE_FRONT_ROWS=90 LE_FRONT_LEVELS=2
