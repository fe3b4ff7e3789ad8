"""Two passes of one workload's chain on a small batch (for ncu: launch lists and --set full captures).
    python scripts/ncu_one.py <workload> [batch]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.argv = [sys.argv[0]] + sys.argv[1:]
import numpy as np, torch
import bench
import graphics_project_b200 as gp

wl = bench.WORKLOADS[sys.argv[1]]()
B = int(sys.argv[2]) if len(sys.argv) > 2 else max(8, wl.batch // 8)
synth = bench.load_synth()
_, uniq = bench.make_frames(synth, wl, min(wl.unique, B, 8), 0)
frames = torch.from_numpy(np.stack([uniq[i % len(uniq)] for i in range(B)])).cuda()
run = bench.EngineRun(wl, gp, torch, torch.device("cuda", 0), frames, B)
for _ in range(2):
    run.step()
torch.cuda.synchronize()
run.eng.check()
print("ok", wl.name, B, float(run.counts.float().mean()))
