import sys; sys.path.insert(0, "/root/repo")
import numpy as np, torch
import graphics_project_b200 as gp
from oracle import restate as R
eng = gp.Engine(0)
g_out = dict(np.load("/root/repo/tests/golden/golden_outputs.npz"))
tot = same = wsame = 0
for name in ["demo_scarce", "sc_white_2", "sc_minecraft_cave", "old_sc_inf", "sc_cube", "sc_scarce_3_flat"]:
    g = torch.from_numpy(g_out[f"{name}/gray"]).cuda()[None]
    for refine, scale, key in ((0, .8, "lsd0"), (1, .8, "lsd1")):
        segs, width, prec, counts = eng.lsd(g, refine=refine, scale=scale)
        n = int(counts[0]); ref = g_out[f"{name}/{key}"][:, 0]; rw = g_out[f"{name}/{key}_w"][:, 0]
        k = min(n, len(ref))
        tot += len(ref); same += int((segs[0, :k].cpu().numpy() == ref[:k]).all(axis=1).sum()); wsame += int((width[0, :k].cpu().numpy() == rw[:k]).sum())
print("segments", same, "/", tot, "widths", wsame, "/", tot)
# synthetic 1080p
for kind in ("voxel", "cave", "cubes"):
    u = gp.synth.make_batch(3, 1080, 1920, kind)
    gray = eng.bgr2gray(torch.from_numpy(np.stack(u)).cuda())
    for refine in (0, 1, 2):
        segs, width, prec, counts, nfa = eng.lsd(gray, refine=refine, want_nfa=True)
        t = s = w = f = 0
        for i in range(3):
            ref = R.lsd(gray[i].cpu().numpy(), refine, return_nfa=True)
            n = int(counts[i]); k = min(n, len(ref[0]))
            t += len(ref[0]); s += int((segs[i, :k].cpu().numpy() == ref[0][:k, 0]).all(axis=1).sum()); w += int((width[i, :k].cpu().numpy() == ref[1][:k, 0]).sum())
            f += int((np.abs(nfa[i, :k].cpu().numpy() - ref[3][:k, 0]) <= 1e-9).sum())
        print(kind, refine, "segments", s, "/", t, "widths", w, "nfa", f)
