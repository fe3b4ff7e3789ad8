#!/bin/bash
# round 2, call O: what the driver runs at round end -- the whole GPU suite, smoke(), both bench arms
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/o_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/o_pytest.log; tail -5 gpurun_out/o_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/o_ref.json 2>gpurun_out/o_ref.err; echo "ref rc=$?"
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/o_bench.json 2>gpurun_out/o_bench.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/o_bench.json').read().strip().splitlines()[-1]); r=json.loads(open('gpurun_out/o_ref.json').read().strip().splitlines()[-1])
print('value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'ref', round(r['value']), 'e2e/ref', round(d['e2e']['value']/r['value'],2), 'roofline', d['roofline']['kernel'], round(d['roofline']['frac'],4), 'traffic', d['roofline']['traffic'], 'parity', d['parity_sample']['equal'], 'same config', d['config']==r['config'], d['clocks'])
"
