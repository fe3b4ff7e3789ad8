#!/bin/bash
# round 2, call H: fresh-context test, then the four workloads at N = 1 with every leg (the lines profiles/ keeps)
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests/test_gpu_batch_parity.py -m gpu -x -q ) > gpurun_out/h_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/h_pytest.log; tail -4 gpurun_out/h_pytest.log
for w in canny_hough_1080p_b256 canny_houghp_4k_b128 lsd_1080p_b512 lines_1080p_cave_b128; do
  timeout 600 python bench.py --steps 10 --warmup 3 --workload $w > gpurun_out/h_bench_$w.json 2> gpurun_out/h_bench_$w.err
  echo "$w rc=$?"; tail -c 200 gpurun_out/h_bench_$w.err
done
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/h_ref_canny_hough.json 2>/dev/null; echo "ref rc=$?"
