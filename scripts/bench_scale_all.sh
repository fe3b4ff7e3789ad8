#!/bin/bash
# the four bench workloads at N GPUs of one box, launched the way the driver launches bench.py (torchrun for N > 1);
# one JSON line per workload into gpurun_out/scale_<workload>_n<N>.json
#   gpurun --gpus N -- 'bash scripts/bench_scale_all.sh N [steps]'
N=${1:-2}; STEPS=${2:-10}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/scale_topo_n$N.txt 2>&1
for w in ${WORKLOADS:-canny_hough_1080p_b256 canny_houghp_4k_b128 lsd_1080p_b512 lines_1080p_cave_b128}; do
  if [ "$N" = "1" ]; then
    timeout 900 python bench.py --gpus 1 --steps $STEPS --warmup 3 --workload $w --no-cpu > gpurun_out/scale_${w}_n$N.json 2> gpurun_out/scale_${w}_n$N.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) \
      bench.py --gpus $N --steps $STEPS --warmup 3 --workload $w --no-cpu > gpurun_out/scale_${w}_n$N.json 2> gpurun_out/scale_${w}_n$N.err
  fi
  echo "$w N=$N rc=$?"
  tail -1 gpurun_out/scale_${w}_n$N.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('  fps', round(d['value']), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'ceiling', round(d['e2e']['copy_ceiling_fps']), 'parity', d['parity_sample']['equal'], d['e2e']['numa'])" 2>/dev/null || tail -3 gpurun_out/scale_${w}_n$N.err
done
