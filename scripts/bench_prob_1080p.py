"""prob() chain (Canny(5,150) + HoughLinesP) on 256 and 1024 x 1080p cube frames and 128 cave frames: ms, frames/s,
per-kernel split.  Environment knobs (LE_PPHT_ROUNDS, LE_PPHT_B) select the PPHT kernel."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import graphics_project_b200 as gp

eng = gp.Engine(0)
tag = " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("LE_PPHT"))
for n, kind in ((256, "cubes"), (1024, "cubes"), (128, "cave")):
    u = gp.synth.make_batch(16, 1080, 1920, kind)
    t = torch.from_numpy(np.stack([u[i % 16] for i in range(n)])).cuda()
    edges = eng.front_canny(t, 5, 150)
    run = lambda: (eng.front_canny(t, 5, 150, out=edges), eng.hough_lines_p(edges, 1, np.pi / 180, 30, 50, 10, max_segs=2048))
    run(); run(); torch.cuda.synchronize()
    eng.timing(True); eng.timing_read()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    kt = {k: round(v[0] / v[1], 3) for k, v in eng.timing_read().items()}; eng.timing(False)
    _, counts = eng.hough_lines_p(edges, 1, np.pi / 180, 30, 50, 10, max_segs=2048)
    print(json.dumps({"tag": tag, "chain": "prob 1080p", "kind": kind, "n": n, "ms": round(ms, 3), "fps": round(n / ms * 1e3),
                      "kernels_ms": kt, "segs_mean": float(counts.float().mean()), "edge_px_mean": float((edges != 0).sum() / n)}), flush=True)
    del t, edges
eng.check()
