#!/bin/bash
# round 2, call E: parity subset, the four workloads at N = 1 (full lines incl. CPU legs), ncu launch lists and
# --set full captures of the kernels that had none in round 1
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_hough.py tests/test_gpu_batch_parity.py tests/test_gpu_mixed.py tests/test_gpu_fuzz.py -m gpu -x -q ) > gpurun_out/e_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/e_pytest.log
tail -3 gpurun_out/e_pytest.log
for w in canny_hough_1080p_b256 canny_houghp_4k_b128 lsd_1080p_b512 lines_1080p_cave_b128; do
  timeout 900 python bench.py --steps 10 --warmup 3 --workload $w > gpurun_out/e_bench_$w.json 2> gpurun_out/e_bench_$w.err
  echo "$w rc=$?"; tail -c 200 gpurun_out/e_bench_$w.err
done
python scripts/bench_prob_1080p.py 2>&1 | tee gpurun_out/e_prob_1080p.jsonl
# ---- ncu: launch lists (cold-cache, serialised: shares only) and --set full of one launch per kernel
for w in canny_houghp_4k_b128:16 lsd_1080p_b512:64 lines_1080p_cave_b128:32; do
  name=${w%%:*}; b=${w##*:}
  python scripts/ncu_one.py $name $b > gpurun_out/e_plain_$name.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/e_launches_$name.csv python scripts/ncu_one.py $name $b > gpurun_out/e_ncu_$name.log 2>&1
  echo "ncu launches $name rc=$?"
done
cap() {  # workload batch kernel-regex skip
  python scripts/ncu_one.py $1 $2 > gpurun_out/e_plain2.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$3 -s $4 -c 1 -f -o gpurun_out/r2_$3 python scripts/ncu_one.py $1 $2 > gpurun_out/e_ncufull_$3.log 2>&1
  echo "ncu full $3 rc=$?"
}
cap canny_houghp_4k_b128 16 k_ppht_rounds 1
cap canny_houghp_4k_b128 16 k_ppht_order 1
cap lsd_1080p_b512 64 k_lsd_grow_spec 1
cap lsd_1080p_b512 64 k_blur_resize_w 1
cap lsd_1080p_b512 64 k_lsd_scatter 1
cap lsd_1080p_b512 64 k_lsd_hist 1
cap lines_1080p_cave_b128 32 k_vp_fit 1
cap lines_1080p_cave_b128 32 k_sphere_hist_batch 1
ls -la gpurun_out/*.ncu-rep | tail -12
