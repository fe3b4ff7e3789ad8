#!/bin/bash
# phase cycle counters of the rounds PPHT kernel (frame 0 of a batch): needs gpurun_variants/ppht_profile.so, built
# like the Makefile does plus -DLE_PPHT_PROFILE
cp graphics-project_b200/liblineengine.so /tmp/lib_default.so
cp gpurun_variants/ppht_profile.so graphics-project_b200/liblineengine.so
for B in ${PPHT_PROFILE_B:-32}; do
LE_PPHT_B=$B python - <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import graphics_project_b200 as gp
eng = gp.Engine(0)
for n, h, w, kind in ((128, 2160, 3840, "cubes"), (256, 1080, 1920, "cubes"), (128, 1080, 1920, "cave")):
    u = gp.synth.make_batch(8, h, w, kind)
    t = torch.from_numpy(np.stack([u[i % 8] for i in range(n)])).cuda()
    edges = eng.front_canny(t, 5, 150)
    torch.cuda.synchronize()
    print(f"== {n} x {w}x{h} {kind}, LE_PPHT_B={os.environ.get('LE_PPHT_B')}", flush=True)
    for _ in range(2):
        eng.hough_lines_p(edges, 1, np.pi / 180, 30, 50, 10, max_segs=2048)
        torch.cuda.synchronize()
    del t, edges
PY
done
cp /tmp/lib_default.so graphics-project_b200/liblineengine.so
