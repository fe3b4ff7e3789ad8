#!/bin/bash
# round 2, call N: A/B of the PPHT round-size adaptation and mask prefetch
mkdir -p gpurun_out
q() { python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(sys.argv[1], 'fps', round(d['value']), 'ms/step', round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['roofline']['kernel_ms_per_launch'].items()}, d['parity_sample']['equal'] if d.get('parity_sample') else None)" "$1"; }
timeout 200 python -m pytest tests/test_gpu_batch_parity.py -m gpu -x -q -k "refused or dense" 2>&1 | tail -3
for v in "A=0" "LE_PPHT_NOPF=1" "LE_PPHT_NOADAPT=1" "LE_PPHT_NOPF=1 LE_PPHT_NOADAPT=1"; do
  env $v timeout 300 python bench.py --steps 5 --warmup 3 --workload canny_houghp_4k_b128 --no-cpu --no-e2e 2>>gpurun_out/n.err | q "4k $v" | tee -a gpurun_out/n_var.log
  env $v timeout 300 python scripts/bench_prob_1080p.py 2>>gpurun_out/n.err | cut -c1-260 | tee -a gpurun_out/n_var.log
done
