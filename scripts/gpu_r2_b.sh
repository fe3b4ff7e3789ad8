#!/bin/bash
# round 2, call B: the new PPHT (rounds over a precomputed order) -- parity first, then timings against the round-1 kernel
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_hough.py tests/test_gpu_batch_parity.py tests/test_gpu_mixed.py tests/test_gpu_fuzz.py -m gpu -x -q ) > gpurun_out/b_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/b_pytest.log
tail -5 gpurun_out/b_pytest.log
for w in canny_hough_1080p_b256 canny_houghp_4k_b128 lsd_1080p_b512 lines_1080p_cave_b128; do
  ( time timeout 900 python bench.py --steps 5 --warmup 3 --workload $w ) > gpurun_out/b_bench_$w.log 2> gpurun_out/b_bench_$w.err
  echo "rc=$?" >> gpurun_out/b_bench_$w.err
  tail -c 300 gpurun_out/b_bench_$w.err
done
q() { python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(sys.argv[1], 'fps', round(d['value']), 'ms/step', round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['roofline']['kernel_ms_per_launch'].items()}, d['parity_sample']['equal'] if d.get('parity_sample') else None)" "$1"; }
for v in "LE_PPHT_ROUNDS=0" "LE_PPHT_B=32" "LE_PPHT_B=64" "LE_PPHT_B=128"; do
  env $v timeout 600 python bench.py --steps 5 --warmup 3 --workload canny_houghp_4k_b128 --no-cpu --no-e2e 2>>gpurun_out/b_var.err | q "4k $v" | tee -a gpurun_out/b_var.log
done
for v in "LE_PPHT_ROUNDS=0" "LE_PPHT_B=32" "LE_PPHT_B=64" "LE_PPHT_B=128"; do
  env $v timeout 600 python scripts/bench_prob_1080p.py 2>>gpurun_out/b_var.err | tee -a gpurun_out/b_var.log
done
