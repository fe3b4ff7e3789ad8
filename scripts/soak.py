"""dev helper: seeded soak of the paths with intentional races or data-dependent control flow (union-find hysteresis,
BFS hysteresis, speculative LSD growing, NFA) against the oracle on random sizes.  Prints the first mismatch."""
import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import graphics_project_b200 as gp
from oracle import restate as R
eng = gp.Engine(0)
n_seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 120
bad = 0
t0 = time.time()
for seed in range(n_seeds):
    rng = np.random.default_rng(1000 + seed)
    h, w = int(rng.integers(24, 420)), int(rng.integers(24, 640))
    kind = ("cubes", "voxel", "cave")[seed % 3]
    img = gp.synth.make_frame(seed, h, w, kind)
    if seed % 4 == 3:  # noise: many weak candidates, ragged regions
        img = np.clip(img.astype(np.int16) + rng.integers(-30, 31, img.shape), 0, 255).astype(np.uint8)
    t = torch.from_numpy(img).cuda()[None]
    g = R.bgr2gray(img)
    lo, hi = sorted(rng.integers(1, 120, 2).tolist())
    blur = bool(seed & 1)
    e = eng.front_canny(t, lo, hi, blur=blur)[0].cpu().numpy()
    ref = R.canny(R.gauss5(g) if blur else g, lo, hi)
    if not np.array_equal(e, ref):
        bad += 1; print("CANNY mismatch seed", seed, (h, w), lo, hi, blur, int((e != ref).sum())); 
    refine = seed % 3; scale = (.8, .5, 1.0)[(seed // 3) % 3]
    segs, width, prec, counts, nfa = eng.lsd(torch.from_numpy(g).cuda()[None], refine=refine, scale=scale, want_nfa=True)
    r = R.lsd(g, refine, scale, return_nfa=True)
    n = int(counts[0]); nr = 0 if r[0] is None else len(r[0])
    ok = n == nr and (n == 0 or np.abs(segs[0, :n].cpu().numpy() - r[0][:, 0]).max() <= 0.5)
    if not ok:
        bad += 1; print("LSD mismatch seed", seed, (h, w), kind, refine, scale, n, nr)
eng.check()
print("soak:", n_seeds, "seeds,", bad, "mismatches,", round(time.time() - t0, 1), "s")
# one big batch: more frames than SMs, so the 8-warp instantiation and several CTAs per SM are exercised
u = np.stack([gp.synth.make_frame(s, 270, 480, ("cubes", "voxel", "cave")[s % 3]) for s in range(40)])
gray = eng.bgr2gray(torch.from_numpy(np.concatenate([u] * 5)).cuda())  # 200 frames
for refine in (0, 1, 2):
    segs, width, prec, counts, nfa = eng.lsd(gray, refine=refine, want_nfa=True)
    c = counts.cpu().numpy()
    badb = 0
    for i in range(40):
        r = R.lsd(gray[i].cpu().numpy(), refine, return_nfa=True)
        nr = 0 if r[0] is None else len(r[0])
        for rep in range(5):
            j = i + 40 * rep
            if c[j] != nr or (nr and not np.array_equal(segs[j, :nr].cpu().numpy(), segs[i, :nr].cpu().numpy())): badb += 1
        if nr and np.abs(segs[i, :nr].cpu().numpy() - r[0][:, 0]).max() > 0.5: badb += 1
    print("batch of 200, refine", refine, "mismatches", badb)
eng.check()
