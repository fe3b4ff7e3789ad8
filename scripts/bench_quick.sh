#!/bin/bash
# quick GPU check: front+hough tests, then the bench kernel split
python -m pytest tests/test_gpu_front.py tests/test_gpu_hough.py -m gpu -q --timeout 900 -x > gpurun_out/tq.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed|^E  " gpurun_out/tq.log | head -20
python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('fps', round(d['value']), 'ms/step', round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['roofline']['kernel_ms_per_launch'].items()}, 'frac', round(d['roofline']['frac'],3), 'step_frac', round(d['roofline']['step_frac_of_hbm_roofline'],3))"
