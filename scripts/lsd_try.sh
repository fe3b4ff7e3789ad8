#!/bin/bash
# LSD: parity tests, then grow-kernel timing of the speculative kernel against the one-warp-per-frame kernel
timeout 900 python -m pytest tests/test_gpu_lsd_vp.py tests/test_gpu_fuzz.py -m gpu -q -x --timeout 600 2>&1 | tail -5
cat > /tmp/lsd_t.py <<'PY'
import sys, json
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import graphics_project_b200 as gp
eng = gp.Engine(0)
for kind, n in (("voxel", 128), ("voxel", 512), ("cave", 512)):
    u = gp.synth.make_batch(16, 1080, 1920, kind)
    t = torch.from_numpy(np.stack([u[i % 16] for i in range(n)])).cuda()
    gray = eng.bgr2gray(t); del t
    for refine in (0, 1):
        eng.lsd(gray, refine=refine); torch.cuda.synchronize()
        eng.timing(True); eng.timing_read()
        segs, _, _, counts = eng.lsd(gray, refine=refine); torch.cuda.synchronize()
        kt = {k: round(v[0] / v[1], 3) for k, v in eng.timing_read().items()}; eng.timing(False)
        print(json.dumps({"kind": kind, "n": n, "refine": refine, "kernels_ms": kt, "segs_mean": float(counts.float().mean()),
                          "sum": int(counts.sum())}), flush=True)
    del gray
PY
echo "== default"; timeout 600 python /tmp/lsd_t.py
