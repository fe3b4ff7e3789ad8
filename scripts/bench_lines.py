"""BASELINE config 5 (--pipeline lines): BGR frames -> gray -> LSD -> polar lines -> Gaussian-sphere histogram ->
(phi, theta) fit for a batch of 1080p cave-style frames, all on the device; prints JSON lines (per-stage CUDA-event
times, frames/s, fraction of the HBM roofline per SURVEY 8d: 3WH + 16 N_segs + 4 n_az n_el bytes per frame).
Usage: python scripts/bench_lines.py [n_frames ...]"""
import json, sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import graphics_project_b200 as gp
from graphics_project_b200 import vp

eng = gp.Engine(0)
H, W = 1080, 1920
peak = 6549.8
try:
    peak = json.load(open("/root/repo/MEASURED_PEAKS.json"))["hbm_gbs"]
except Exception:
    pass

def ev():
    return torch.cuda.Event(enable_timing=True)

for n in [int(a) for a in sys.argv[1:]] or [128, 512]:
    u = gp.synth.make_batch(min(16, n), H, W, "cave")
    t = torch.from_numpy(np.stack([u[i % len(u)] for i in range(n)])).cuda()
    def step(marks=None):
        m = (lambda: marks.append(ev()) or marks[-1].record()) if marks is not None else (lambda: None)
        m(); gray = eng.bgr2gray(t)
        m(); segs, _, _, counts = eng.lsd(gray, max_segs=2048)
        m(); polar, pcounts = eng.polar_batch(segs, counts, 10.0)
        m(); fit, hist = vp.regress_lines_batch(polar, pcounts, 1000, 500, 0.3, float(W), float(H), engine=eng)
        m()
        return fit, hist, pcounts, counts
    step(); torch.cuda.synchronize()
    e0, e1 = ev(), ev()
    reps = 3
    e0.record()
    for _ in range(reps): step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    marks = []
    fit, hist, pcounts, counts = step(marks); torch.cuda.synchronize()
    names = ["bgr2gray", "lsd", "polar", "hist+fit"]
    stages = {k: round(marks[i].elapsed_time(marks[i + 1]), 3) for i, k in enumerate(names)}
    alg = 3 * W * H + 16 * float(counts.float().mean()) + 4 * 360 * 90
    fit_h = fit.cpu().numpy()
    print(json.dumps({"chain": "lines pipeline (cfg 5): gray+LSD+polar+sphere hist+VP fit", "kind": "cave", "n": n, "h": H, "w": W,
                      "ms": round(ms, 3), "fps": round(n / ms * 1e3), "stages_ms": stages,
                      "algorithmic_bytes_per_frame": round(alg), "frac_of_measured_hbm": round(alg * n / (ms * 1e-3) / 1e9 / peak, 4),
                      "segs_mean": float(counts.float().mean()), "lines_mean": float(pcounts.float().mean()),
                      "loss_median": float(np.median(fit_h[:, 2])), "hist_votes_mean": float(hist.sum((1, 2)).float().mean())}), flush=True)
    del t
