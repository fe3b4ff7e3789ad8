#!/bin/bash
# round 2, call C: PPHT (un-vote combining, staged RNG, round size sweep), LSD (multi-entry BFS trips, tag epochs)
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_hough.py tests/test_gpu_batch_parity.py tests/test_gpu_mixed.py tests/test_gpu_fuzz.py tests/test_gpu_lsd_vp.py tests/test_gpu_lines_batch.py -m gpu -x -q ) > gpurun_out/c_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c_pytest.log
tail -5 gpurun_out/c_pytest.log
q() { python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(sys.argv[1], 'fps', round(d['value']), 'ms/step', round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['roofline']['kernel_ms_per_launch'].items()}, d['parity_sample']['equal'] if d.get('parity_sample') else None, 'e2e', round(d['e2e']['value']) if d.get('e2e') else None)" "$1"; }
for v in "LE_PPHT_B=8" "LE_PPHT_B=16" "LE_PPHT_B=24" "LE_PPHT_B=32" "LE_PPHT_B=48"; do
  env $v timeout 600 python bench.py --steps 5 --warmup 3 --workload canny_houghp_4k_b128 --no-cpu --no-e2e 2>>gpurun_out/c_var.err | q "4k $v" | tee -a gpurun_out/c_var.log
done
for v in "LE_PPHT_B=8" "LE_PPHT_B=16" "LE_PPHT_B=32"; do
  env $v timeout 600 python scripts/bench_prob_1080p.py 2>>gpurun_out/c_var.err | tee -a gpurun_out/c_var.log
done
for w in lsd_1080p_b512 lines_1080p_cave_b128 canny_houghp_4k_b128; do
  timeout 900 python bench.py --steps 5 --warmup 3 --workload $w --no-cpu 2>>gpurun_out/c_var.err | tee gpurun_out/c_bench_$w.log | q "$w" | tee -a gpurun_out/c_var.log
done
timeout 600 python scripts/bench_front_content.py 2>>gpurun_out/c_var.err | tee gpurun_out/c_front_content.jsonl
