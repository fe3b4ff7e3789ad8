#!/bin/bash
# launch list of one LSD call (128 voxel frames)
python scripts/lsd_launches.py 128 > gpurun_out/lsd_plain.log 2>&1 || { tail -5 gpurun_out/lsd_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r1_s4_lsd_launches.csv python scripts/lsd_launches.py 128 > gpurun_out/lsd_ncu.log 2>&1
tail -2 gpurun_out/lsd_ncu.log
