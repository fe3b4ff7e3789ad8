#!/bin/bash
# round 2, call J: LSD pixel-set size variant (more L1 for four frames per SM), final ncu captures
mkdir -p gpurun_out
q() { python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(sys.argv[1], 'fps', round(d['value']), 'ms/step', round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['roofline']['kernel_ms_per_launch'].items()}, d['parity_sample']['equal'] if d.get('parity_sample') else None)" "$1"; }
( time timeout 400 python -m pytest tests/test_gpu_lsd_vp.py tests/test_gpu_batch_parity.py -m gpu -x -q ) > gpurun_out/j_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/j_pytest.log; tail -4 gpurun_out/j_pytest.log
timeout 300 python bench.py --steps 5 --warmup 3 --workload lsd_1080p_b512 --no-cpu --no-e2e 2>gpurun_out/j_lsd.err | q "lsd default" | tee -a gpurun_out/j_var.log
cp graphics-project_b200/liblineengine.so /tmp/lib_default.so
cp gpurun_variants/lsd_h512.so graphics-project_b200/liblineengine.so
( timeout 300 python -m pytest tests/test_gpu_lsd_vp.py -m gpu -x -q 2>&1 | tail -2 ) | tee -a gpurun_out/j_var.log
timeout 300 python bench.py --steps 5 --warmup 3 --workload lsd_1080p_b512 --no-cpu --no-e2e 2>>gpurun_out/j_lsd.err | q "lsd SG_HASH=512" | tee -a gpurun_out/j_var.log
timeout 300 python bench.py --steps 5 --warmup 3 --workload lines_1080p_cave_b128 --no-cpu --no-e2e 2>>gpurun_out/j_lsd.err | q "lines SG_HASH=512" | tee -a gpurun_out/j_var.log
cp /tmp/lib_default.so graphics-project_b200/liblineengine.so
cap() {  # workload batch kernel-regex skip name
  python scripts/ncu_one.py $1 $2 > gpurun_out/j_plain.log 2>&1 &&
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$3 -s $4 -c 1 -f -o gpurun_out/r2f_$5 python scripts/ncu_one.py $1 $2 > gpurun_out/j_ncufull_$5.log 2>&1
  echo "ncu full $5 rc=$?"
}
cap canny_houghp_4k_b128 128 k_ppht_rounds 1 k_ppht_rounds_b128
cap lsd_1080p_b512 512 k_lsd_grow_spec 1 k_lsd_grow_spec_b512
cap lsd_1080p_b512 64 k_blur_resize_w 1 k_blur_resize_w
cap lsd_1080p_b512 64 k_lsd_hist 1 k_lsd_hist
cap lsd_1080p_b512 64 k_lsd_scatter 1 k_lsd_scatter
ls -la gpurun_out/r2f_*.ncu-rep
