#!/bin/bash
# round 2, call L: LSD growth with the parallel alignment test
mkdir -p gpurun_out
q() { python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(sys.argv[1], 'fps', round(d['value']), 'ms/step', round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['roofline']['kernel_ms_per_launch'].items()}, d['parity_sample']['equal'] if d.get('parity_sample') else None)" "$1"; }
( time timeout 400 python -m pytest tests/test_gpu_lsd_vp.py tests/test_gpu_batch_parity.py tests/test_gpu_lines_batch.py tests/test_gpu_mixed.py -m gpu -x -q ) > gpurun_out/l_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/l_pytest.log; tail -4 gpurun_out/l_pytest.log
timeout 300 python bench.py --steps 5 --warmup 3 --workload lsd_1080p_b512 --no-cpu --no-e2e 2>gpurun_out/l_lsd.err | q "lsd" | tee -a gpurun_out/l_var.log
timeout 300 python bench.py --steps 5 --warmup 3 --workload lines_1080p_cave_b128 --no-cpu --no-e2e 2>>gpurun_out/l_lsd.err | q "lines" | tee -a gpurun_out/l_var.log
