#!/bin/bash
# dev helper: quick tests + kernel split for the in-tree library, then the kernel split of every variant
# library under gpurun_variants/ (built here with other -D settings)
bash scripts/bench_quick.sh
cp graphics-project_b200/liblineengine.so /tmp/lib_default.so
for v in gpurun_variants/*.so; do
  [ -f "$v" ] || continue
  echo "== $v"
  cp "$v" graphics-project_b200/liblineengine.so
  python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('fps', round(d['value']), 'ms/step', round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['roofline']['kernel_ms_per_launch'].items()})"
done
cp /tmp/lib_default.so graphics-project_b200/liblineengine.so
