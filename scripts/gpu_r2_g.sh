#!/bin/bash
# round 2, call G: sort_frame fix (short timeouts first), 3-channel Canny, full GPU suite, content benches
mkdir -p gpurun_out
( time timeout 180 python -m pytest tests/test_gpu_hough.py tests/test_gpu_front.py -m gpu -x -q ) > gpurun_out/g_pytest1.log 2>&1
echo "pytest1 rc=$?" >> gpurun_out/g_pytest1.log; tail -4 gpurun_out/g_pytest1.log
if grep -q "rc=0" gpurun_out/g_pytest1.log; then
  timeout 240 python scripts/bench_front_content.py 2>gpurun_out/g_front.err | tee gpurun_out/g_front_content.jsonl | cut -c1-330
  ( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/g_pytest2.log 2>&1
  echo "pytest2 rc=$?" >> gpurun_out/g_pytest2.log; tail -4 gpurun_out/g_pytest2.log
fi
