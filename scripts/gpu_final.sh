#!/bin/bash
# end-of-session battery: every GPU test, smoke(), both bench arms, other chains, lines pipeline, ncu launch list + one
# --set full capture of each dominant kernel (every ncu run preceded by the same command exiting 0 without ncu)
S=${1:-r1_s5}
python -m pytest tests -m gpu -x -q --timeout 900 > gpurun_out/t_all.log 2>&1; tail -2 gpurun_out/t_all.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log | cut -c1-300
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>gpurun_out/bench_ref.err; tail -1 gpurun_out/bench_ref.log | cut -c1-400
python bench.py > gpurun_out/bench_full.log 2>gpurun_out/bench_full.err; tail -1 gpurun_out/bench_full.log | cut -c1-1800
python scripts/bench_others.py > gpurun_out/others.log 2>&1; grep -c "^{" gpurun_out/others.log
python scripts/bench_lines.py 128 512 > gpurun_out/lines.log 2>&1; tail -1 gpurun_out/lines.log | cut -c1-400
python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/plain_l.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${S}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_l.log 2>&1
tail -1 gpurun_out/ncu_l.log | cut -c1-200
for k in k_front_rows k_hough_vote; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -o gpurun_out/${S}_$k -f python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_$k.log 2>&1
done
ls -la gpurun_out/${S}_*
