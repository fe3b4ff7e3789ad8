#!/bin/bash
# LSD parity tests (incl. fused blur+resize vs LE_LSD_UNFUSED), the batched lines pipeline, LSD per-kernel times
timeout 500 python -m pytest tests/test_gpu_lsd_vp.py tests/test_gpu_lines_batch.py tests/test_gpu_mixed.py tests/test_gpu_hough.py tests/test_gpu_fuzz.py -m gpu -x -q 2>&1 | tail -8
timeout 200 python scripts/bench_lines.py 128 512 > gpurun_out/lines.log 2>&1; tail -3 gpurun_out/lines.log | cut -c1-900
timeout 100 python scripts/lsd_launches.py > gpurun_out/lsd_plain2.log 2>&1; tail -2 gpurun_out/lsd_plain2.log
timeout 200 python scripts/bench_others.py > gpurun_out/others2.log 2>&1; head -3 gpurun_out/others2.log | cut -c1-400
