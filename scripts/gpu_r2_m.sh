#!/bin/bash
# round 2, call M: PPHT mask prefetch, robustness tests
mkdir -p gpurun_out
q() { python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(sys.argv[1], 'fps', round(d['value']), 'ms/step', round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['roofline']['kernel_ms_per_launch'].items()}, d['parity_sample']['equal'] if d.get('parity_sample') else None)" "$1"; }
( time timeout 300 python -m pytest tests/test_gpu_hough.py tests/test_gpu_batch_parity.py tests/test_gpu_mixed.py tests/test_gpu_fuzz.py -m gpu -x -q ) > gpurun_out/m_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/m_pytest.log; tail -12 gpurun_out/m_pytest.log
timeout 300 python bench.py --steps 5 --warmup 3 --workload canny_houghp_4k_b128 --no-cpu --no-e2e 2>gpurun_out/m.err | q "4k" | tee -a gpurun_out/m_var.log
timeout 300 python scripts/bench_prob_1080p.py 2>>gpurun_out/m.err | tee -a gpurun_out/m_var.log
