#!/bin/bash
cp graphics-project_b200/liblineengine.so /tmp/lib_default.so
cp gpurun_variants/lsdstats.so graphics-project_b200/liblineengine.so
python - <<'PY'
import sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import graphics_project_b200 as gp
eng = gp.Engine(0)
for kind in ("voxel", "cave"):
    u = gp.synth.make_batch(1, 1080, 1920, kind)
    gray = eng.bgr2gray(torch.from_numpy(np.stack(u)).cuda())
    segs, _, _, counts = eng.lsd(gray, refine=0); torch.cuda.synchronize()
    print(kind, int(counts[0]), flush=True)
PY
cp /tmp/lib_default.so graphics-project_b200/liblineengine.so
