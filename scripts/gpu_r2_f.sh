#!/bin/bash
# round 2, call F: LSD after the prep changes and the parallel pre-commit; the frames of rank 6 of the 8-GPU run
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_lsd_vp.py tests/test_gpu_lines_batch.py tests/test_gpu_batch_parity.py tests/test_gpu_hough.py tests/test_gpu_mixed.py -m gpu -x -q ) > gpurun_out/f_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/f_pytest.log
tail -15 gpurun_out/f_pytest.log
q() { python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(sys.argv[1], 'fps', round(d['value']), 'ms/step', round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['roofline']['kernel_ms_per_launch'].items()}, d['parity_sample']['equal'] if d.get('parity_sample') else None, 'e2e', round(d['e2e']['value']) if d.get('e2e') else None)" "$1"; }
for sr in 0 6 3; do
  timeout 600 python bench.py --steps 5 --warmup 3 --workload lsd_1080p_b512 --no-cpu --seed-rank $sr 2>gpurun_out/f_lsd_sr$sr.err | tee gpurun_out/f_lsd_sr$sr.json | q "lsd seed-rank $sr"
  tail -3 gpurun_out/f_lsd_sr$sr.err
done
timeout 600 python bench.py --steps 5 --warmup 3 --workload lines_1080p_cave_b128 --no-cpu 2>gpurun_out/f_lines.err | tee gpurun_out/f_lines.json | q "lines"
timeout 600 python scripts/bench_front_content.py 2>gpurun_out/f_front.err | tee gpurun_out/f_front_content.jsonl | cut -c1-330
