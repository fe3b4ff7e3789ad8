#!/bin/bash
# Hough vote CTA shape comparison: variant libraries built with -DHV_THREADS / -DHV_MINB / -DHV_ZERO_WORDS into
# gpurun_variants/vote_*.so (see the build lines in profiles/r1_ncu_summary.md); each one runs the Hough parity tests
# and the bench kernel split
cp graphics-project_b200/liblineengine.so /tmp/lib_default.so
run() {
  python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('fps', round(d['value']), 'ms/step', round(d['ms_per_step'],4), {k: round(v,4) for k,v in d['roofline']['kernel_ms_per_launch'].items()})"
}
echo "== default"; run
for v in "$@"; do
  cp gpurun_variants/vote_$v.so graphics-project_b200/liblineengine.so
  echo "== $v"; python -m pytest tests/test_gpu_hough.py -m gpu -q -x 2>&1 | tail -1; run
done
cp /tmp/lib_default.so graphics-project_b200/liblineengine.so
