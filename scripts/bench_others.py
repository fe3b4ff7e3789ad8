"""Throughput of the other hot-path rows (HoughLinesP, LSD, VP) on synthetic batches; prints JSON lines."""
import json, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import graphics_project_b200 as gp

def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

eng = gp.Engine(0)
def frames(n, h, w, kind, uniq=16):
    u = gp.synth.make_batch(min(uniq, n), h, w, kind)
    return torch.from_numpy(np.stack([u[i % len(u)] for i in range(n)])).cuda()

out = []
for (n, h, w, kind) in ((256, 1080, 1920, "cubes"), (1024, 1080, 1920, "cubes"), (128, 2160, 3840, "cubes")):
    t = frames(n, h, w, kind)
    edges = eng.front_canny(t, 5, 150)
    eng.timing(True); eng.timing_read()
    ms = timeit(lambda: (eng.front_canny(t, 5, 150, out=edges), eng.hough_lines_p(edges, 1, np.pi/180, 30, 50, 10)))
    kt = {k: round(v[0]/v[1], 3) for k, v in eng.timing_read().items()}; eng.timing(False)
    segs, counts = eng.hough_lines_p(edges, 1, np.pi/180, 30, 50, 10)
    out.append({"chain": "prob: Canny(5,150)+HoughLinesP(30,50,10)", "n": n, "h": h, "w": w, "ms": round(ms, 3), "fps": round(n / ms * 1e3), "kernels_ms": kt, "segs_mean": float(counts.float().mean())})
    print(json.dumps(out[-1]), flush=True)
    del t, edges
# Canny + standard HoughLines (the bench.py chain) at 4K, with the accumulator written
for (n, h, w) in ((64, 2160, 3840),):
    t = frames(n, h, w, "cubes")
    na, nr = eng.hough_dims(h, w)
    edges = torch.empty((n, h, w), dtype=torch.uint8, device="cuda")
    outs = (torch.empty((n, 1024, 2), dtype=torch.float32, device="cuda"), None, torch.empty((n,), dtype=torch.int32, device="cuda"),
            torch.empty((n, na + 2, nr + 2), dtype=torch.int32, device="cuda"))
    run = lambda: (eng.front_canny(t, 5, 10, blur=True, out=edges), eng.hough_lines(None, 1.0, np.pi / 180, 50, out=outs))
    run(); eng.timing(True); eng.timing_read()
    ms = timeit(run, reps=5)
    kt = {k: round(v[0]/v[1], 3) for k, v in eng.timing_read().items()}; eng.timing(False)
    alg = 4 * w * h + 4 * (na + 2) * (nr + 2)
    out.append({"chain": "Canny(5,10)+blur+HoughLines(50), accumulator written", "n": n, "h": h, "w": w, "ms": round(ms, 3),
                "fps": round(n / ms * 1e3), "kernels_ms": kt, "algorithmic_bytes_per_frame": alg,
                "frac_of_measured_hbm_6549.8": round(alg * n / (ms * 1e-3) / 1e9 / 6549.8, 3), "lines_mean": float(outs[2].float().mean())})
    print(json.dumps(out[-1]), flush=True)
    del t, edges, outs
# LSD_REFINE_ADV at the simulator's parameters (opencv.py:142-145)
t = frames(128, 1080, 1920, "voxel"); gray = eng.bgr2gray(t); del t
eng.lsd(gray, refine=2, scale=.5); eng.timing(True); eng.timing_read()
ms = timeit(lambda: eng.lsd(gray, refine=2, scale=.5), reps=2)
kt = {k: round(v[0]/v[1], 3) for k, v in eng.timing_read().items()}; eng.timing(False)
segs, _, _, counts = eng.lsd(gray, refine=2, scale=.5)
out.append({"chain": "lsd(refine 2, scale .5)", "kind": "voxel", "n": 128, "h": 1080, "w": 1920, "ms": round(ms, 3), "fps": round(128 / ms * 1e3), "kernels_ms": kt, "segs_mean": float(counts.float().mean())})
print(json.dumps(out[-1]), flush=True)
del gray
for (n, h, w, kind) in ((128, 1080, 1920, "voxel"), (512, 1080, 1920, "voxel"), (512, 1080, 1920, "cave")):
    t = frames(n, h, w, kind)
    gray = eng.bgr2gray(t); del t
    eng.timing(True); eng.timing_read()
    ms = timeit(lambda: eng.lsd(gray), reps=2)
    kt = {k: round(v[0]/v[1], 3) for k, v in eng.timing_read().items()}; eng.timing(False)
    segs, _, _, counts = eng.lsd(gray)
    out.append({"chain": "lsd(refine 0, scale .8)", "kind": kind, "n": n, "h": h, "w": w, "ms": round(ms, 3), "fps": round(n / ms * 1e3), "kernels_ms": kt, "segs_mean": float(counts.float().mean())})
    print(json.dumps(out[-1]), flush=True)
    del gray
# VP: regress_lines on LSD segments of one frame
img = gp.synth.make_frame(1, 400, 600, "voxel")
pairs = gp.detectors.lsd(img)
lines = gp.vp.edges_to_polar_lines(pairs)
t0 = time.perf_counter(); (phi, theta), loss = gp.vp.regress_lines(lines, 1000, 500, np.deg2rad(15), verbose=False); dt = time.perf_counter() - t0
print(json.dumps({"chain": "regress_lines(1000,500,15deg)", "lines": len(lines), "seconds": round(dt, 4), "loss": loss}))
# CPU side for comparison (cv2, single frame, single thread)
import cv2 as cv
cv.setNumThreads(1)
f = gp.synth.make_frame(0, 1080, 1920, "cubes"); g = cv.cvtColor(f, cv.COLOR_BGR2GRAY)
t0 = time.perf_counter()
for _ in range(5): e = cv.Canny(cv.cvtColor(f, cv.COLOR_BGR2GRAY), 5, 150); cv.HoughLinesP(e, 1, np.pi/180, 30, minLineLength=50, maxLineGap=10)
print(json.dumps({"cv2_1thread_prob_1080p_ms": round((time.perf_counter() - t0) / 5 * 1e3, 2)}))
f = gp.synth.make_frame(0, 1080, 1920, "voxel"); g = cv.cvtColor(f, cv.COLOR_BGR2GRAY)
d = cv.createLineSegmentDetector(0, scale=.8, sigma_scale=.6, quant=2.0, ang_th=22.5, log_eps=0, density_th=.7, n_bins=1024)
t0 = time.perf_counter()
for _ in range(3): d.detect(g)
print(json.dumps({"cv2_1thread_lsd_1080p_voxel_ms": round((time.perf_counter() - t0) / 3 * 1e3, 2)}))
# segment merging: one CTA per list; the reference's recursive Python scan on the same list
pairs = gp.detectors.lsd(gp.synth.make_frame(3, 400, 600, "cave"))
segs = torch.from_numpy(np.array([[a[0], a[1], b[0], b[1]] for a, b in pairs], dtype=np.float64))[None].repeat(64, 1, 1).cuda()
ms = timeit(lambda: eng.combine_parallel_lines(segs))
_, _, oc = eng.combine_parallel_lines(segs)
print(json.dumps({"chain": "combineParallelLines", "lists": 64, "segments_in": segs.shape[1], "segments_out": int(oc[0]), "ms_per_64_lists": round(ms, 3)}))
