#!/bin/bash
# front kernel occupancy variants (gpurun_variants/front_<blocks per SM>_<TMA stages>.so, built with
# -DFR_MINBLK=.. -DFR_NS=..): parity tests of the front end, then the default bench's kernel split
cp graphics-project_b200/liblineengine.so /tmp/lib_default.so
run() { python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(sys.argv[1], 'fps', round(d['value']), 'ms/step', round(d['ms_per_step'],4), {k: round(v,4) for k,v in d['roofline']['kernel_ms_per_launch'].items()}, 'parity', d['parity_sample']['equal'])" "$1"; }
run default
for v in "$@"; do
  cp gpurun_variants/front_$v.so graphics-project_b200/liblineengine.so
  timeout 200 python -m pytest tests/test_gpu_front.py tests/test_gpu_fuzz.py -m gpu -q -x 2>&1 | tail -1
  run "front_$v"
done
cp /tmp/lib_default.so graphics-project_b200/liblineengine.so
