"""Content dependence of the Canny + HoughLines chain: the default bench frames (cube edges) are
> 90 % constant rows, which the front kernel's flat mode skips.  Same chain, same batch shape (256 x 1080p), on
voxel, cave-style and uniform-noise frames: per-kernel ms and fraction of the measured HBM peak."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import graphics_project_b200 as gp

H, W, B = 1080, 1920, 256
peak = 6549.8
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
eng = gp.Engine(0, H, W, B)
na, nr = eng.hough_dims(H, W)
edges = torch.empty((B, H, W), dtype=torch.uint8, device="cuda")
outs = (torch.empty((B, 1024, 2), dtype=torch.float32, device="cuda"), None, torch.empty((B,), dtype=torch.int32, device="cuda"),
        torch.empty((B, na + 2, nr + 2), dtype=torch.int32, device="cuda"))
for kind in ("cubes", "voxel", "cave", "noise"):
    if kind == "noise":
        rng = np.random.default_rng(0)
        u = rng.integers(0, 256, (8, H, W, 3), dtype=np.uint8)
    else:
        u = gp.synth.make_batch(8, H, W, kind)
    t = torch.from_numpy(np.stack([u[i % 8] for i in range(B)])).cuda()
    run = lambda: (eng.front_canny(t, 5, 10, blur=True, out=edges), eng.hough_lines(None, 1.0, np.pi / 180, 50, out=outs))
    for _ in range(3): run()
    torch.cuda.synchronize()
    eng.timing(True); eng.timing_read()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    kt = {k: v[0] / v[1] for k, v in eng.timing_read().items()}; eng.timing(False)
    overflow = None
    try:
        eng.check()
    except gp.EngineError as e:
        overflow = str(e)
    front_b, acc_b = 4 * W * H * B, 4 * (na + 2) * (nr + 2) * B
    print(json.dumps({"frames": kind, "n": B, "ms_per_step": round(ms, 4), "fps": round(B / ms * 1e3),
                      "kernel_ms": {k: round(v, 4) for k, v in kt.items()},
                      "front_frac_of_hbm_peak": round(front_b / (kt["front_nms"] * 1e-3) / 1e9 / peak, 4),
                      "vote_frac_of_hbm_peak": round(acc_b / (kt["hough_vote"] * 1e-3) / 1e9 / peak, 4),
                      "step_frac_of_hbm_peak": round((front_b + acc_b) / (ms * 1e-3) / 1e9 / peak, 4),
                      "edge_px_per_frame": float((edges != 0).sum() / B), "lines_per_frame": float(outs[2].float().mean()),
                      "peak_list_overflow": overflow}), flush=True)
    del t
