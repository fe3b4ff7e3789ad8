#!/bin/bash
# weak-scaling loss of the default workload: does the NCCL gather cost SM slots?  Same bench at N GPUs with fewer /
# smaller NCCL CTAs (NCCL_MAX_NCHANNELS, NCCL_NTHREADS) and without any gather (LE_BENCH_NO_GATHER=1)
N=${1:-2}
run() { env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) \
  bench.py --gpus $N --steps 20 --warmup 5 --no-cpu --no-e2e --no-parity 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(sys.argv[1], 'fps', round(d['value']), 'ms/step', round(d['ms_per_step'],4))" "$*"; }
run A=0
run NCCL_MAX_NCHANNELS=1
run NCCL_MAX_NCHANNELS=2 NCCL_NTHREADS=128
run LE_BENCH_NO_GATHER=1
run LE_BENCH_GATHER_EVERY=4
