"""dev helper: one LSD call on 128 voxel frames (for an ncu launch list)."""
import sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import graphics_project_b200 as gp
eng = gp.Engine(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
u = gp.synth.make_batch(16, 1080, 1920, "voxel")
t = torch.from_numpy(np.stack([u[i % 16] for i in range(n)])).cuda()
gray = eng.bgr2gray(t); del t
for _ in range(2):
    segs, _, _, counts = eng.lsd(gray)
torch.cuda.synchronize()
print(float(counts.float().mean()))
