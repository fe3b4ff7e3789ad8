#!/bin/bash
# round 2, call A: the whole GPU suite, then the four bench workloads at N = 1 (short)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/a_smi.txt
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/a_pytest.log
for w in canny_hough_1080p_b256 canny_houghp_4k_b128 lsd_1080p_b512 lines_1080p_cave_b128; do
  ( time timeout 600 python bench.py --steps 5 --warmup 3 --workload $w ) > gpurun_out/a_bench_$w.log 2> gpurun_out/a_bench_$w.err
  echo "rc=$?" >> gpurun_out/a_bench_$w.err
done
tail -3 gpurun_out/a_pytest.log
for w in canny_hough_1080p_b256 canny_houghp_4k_b128 lsd_1080p_b512 lines_1080p_cave_b128; do tail -c 600 gpurun_out/a_bench_$w.log; tail -5 gpurun_out/a_bench_$w.err; done
