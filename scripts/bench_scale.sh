#!/bin/bash
# N-GPU bench exactly as the driver launches it; usage (under gpurun --gpus N): bash scripts/bench_scale.sh N
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err
tail -1 gpurun_out/bench_n$N.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('N', d['n_gpus'], 'fps', round(d['value']), 'ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), d['scaling'])"
