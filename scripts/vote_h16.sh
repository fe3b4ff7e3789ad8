#!/bin/bash
# half2-staged vote loop (default for frames up to 2048 x 2048) against float2 staging (LE_VOTE_F32=1)
python -m pytest tests/test_gpu_hough.py tests/test_gpu_fuzz.py tests/test_gpu_mixed.py tests/test_gpu_front.py -m gpu -q -x 2>&1 | tail -2
run() {
  python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('fps', round(d['value']), 'ms/step', round(d['ms_per_step'],4), {k: round(v,4) for k,v in d['roofline']['kernel_ms_per_launch'].items()})"
}
echo "== h16"; run
echo "== f32"; LE_VOTE_F32=1 run
echo "== h16"; run
