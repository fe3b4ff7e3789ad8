#!/bin/bash
# does the nvidia-smi clock sampler cost step time at N GPUs?  default workload, with and without it
N=${1:-4}
run() { env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) \
  bench.py --gpus $N --steps 20 --warmup 5 --no-cpu --no-e2e --no-parity 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(sys.argv[1], 'fps', round(d['value']), 'ms/step', round(d['ms_per_step'],4), d['clocks'])" "$*"; }
run A=0
run LE_BENCH_NO_SAMPLER=1
run LE_BENCH_NO_SAMPLER=1 LE_BENCH_NO_GATHER=1
run A=0
