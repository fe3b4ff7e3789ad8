#!/usr/bin/env python
"""Benchmark of the line-detection hot path (BASELINE.json: frames/s of Canny + Hough / HoughLinesP / LSD chains).

    python bench.py --gpus N --steps K --warmup W [--workload NAME]      # this engine
    python bench.py --impl reference --gpus N ... [--workload NAME]        # the reference's cv2 CPU path

Workloads (BASELINE.json configs 2-5; one step = one pass of the chain over one batch per GPU):
  canny_hough_1080p_b256  (default) 256 cube-edge 1080p frames: gray -> GaussianBlur(5,5) -> Canny(5,10) ->
                          HoughLines(1, pi/180, 50), full int32 accumulator written (opencv_fit_color.py:74-83)
  canny_houghp_4k_b128    128 cube-edge 4K frames: prob() = gray -> Canny(5,150) -> HoughLinesP(1, pi/180, 30, 50, 10)
                          in cv2's exact pixel visiting order (opencv.py:211-219); 1024 frames over 8 GPUs
  lsd_1080p_b512          512 voxel-world 1080p frames: lsd() = gray -> LSD(refine 0, scale .8) (opencv.py:204-209)
  lines_1080p_cave_b128   128 cave-style 1080p frames: get_camera_angles(method='lsd') = lsd -> polar lines ->
                          vanishing-point fit with Gaussian-sphere seeds (opencv_fit_color.py:85-93)
Prints ONE JSON line (see the task contract): value = whole-job frames/s with inputs resident in HBM; e2e = the same
through the host-buffer API (pinned host frames, H2D + D2H inside the timed region); roofline = the dominant
kernel's algorithmic bytes / its live CUDA-event duration vs the measured HBM peak; cpu_baseline = the reference's
cv2 chain timed on this box's cores; parity_sample = frames of the TIMED batch compared with cv2 after timing.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


# ----------------------------------------------------------------------------------------------- workloads
class Workload:
    """Static description + the reference's cv2 chain.  Engine-side state lives in ``EngineRun``."""
    name = metric = kind = ""
    h = w = batch = unique = 0
    ref_step_frames_per_core = 2  # frames per core in one step of the --impl reference arm

    def config(self):
        """The ``config`` object of the JSON line: identical in both arms."""
        return {"workload": self.name, "height": self.h, "width": self.w, "batch_per_gpu": self.batch,
                "scene": self.kind, "unique_frames": self.unique, "chain": self.chain,
                "l2": f"inputs larger than L2 ({self.batch * self.h * self.w * 3 / 1e6:.0f} MB BGR per step)"}

    def chain_bytes(self):
        """compulsory HBM bytes per frame of the whole chain."""
        raise NotImplementedError

    def cv2_chain(self, cv, frame):
        raise NotImplementedError

    def cv2_stages(self, cv, frame):
        """[(stage name, seconds)] for one frame, the reference's calls one by one."""
        raise NotImplementedError

    needs_processes = False  # the reference chain holds the GIL (pure-Python search): pool of processes


class CannyHough(Workload):
    name = "canny_hough_1080p_b256"
    metric = "1080p frames/s Canny+HoughLines (batch of 256 synthetic cube-edge frames per GPU)"
    h, w, batch, unique, kind = 1080, 1920, 256, 32, "cubes"
    chain = {"canny": [5, 10], "blur": "5x5", "hough": [1, "pi/180", 50], "accumulator_written": True}
    ref_step_frames_per_core = 16
    LO, HI, THR = 5, 10, 50

    def chain_bytes(self):
        numrho = 2 * (self.w + self.h) + 1
        return 3 * self.w * self.h + self.w * self.h + 4 * 182 * (numrho + 2)

    def kernel_bytes(self):
        numrho = 2 * (self.w + self.h) + 1
        return {"front_nms": 4 * self.w * self.h, "hough_vote": 4 * 182 * (numrho + 2)}

    def cv2_chain(self, cv, frame):  # opencv_fit_color.py:74-83, verbatim
        gray = cv.cvtColor(frame, cv.COLOR_BGR2GRAY)
        blurred = cv.GaussianBlur(gray, (5, 5), 0)
        canny = cv.Canny(blurred, self.LO, self.HI)
        return cv.HoughLines(canny, 1, np.pi / 180, self.THR, None, 0, 0)

    def cv2_stages(self, cv, frame):
        t = [time.perf_counter()]
        gray = cv.cvtColor(frame, cv.COLOR_BGR2GRAY); t.append(time.perf_counter())
        blurred = cv.GaussianBlur(gray, (5, 5), 0); t.append(time.perf_counter())
        canny = cv.Canny(blurred, self.LO, self.HI); t.append(time.perf_counter())
        cv.HoughLines(canny, 1, np.pi / 180, self.THR, None, 0, 0); t.append(time.perf_counter())
        return list(zip(("cvtColor", "GaussianBlur", "Canny", "HoughLines"), np.diff(t)))


class CannyHoughP(Workload):
    name = "canny_houghp_4k_b128"
    metric = "4K frames/s Canny+HoughLinesP, cv2's exact visiting order (batch of 128 synthetic cube-edge frames per GPU)"
    h, w, batch, unique, kind = 2160, 3840, 128, 16, "cubes"
    chain = {"canny": [5, 150], "houghp": [1, "pi/180", 30, 50, 10], "order": "exact (cv::RNG replay)"}

    def chain_bytes(self):  # 3WH + WH (+ 16 B per segment)
        return 4 * self.w * self.h

    def kernel_bytes(self):
        # front: BGR in + edge map out; edge list: edge map in; PPHT proper has no compulsory HBM traffic beyond
        # the edge map it is handed (its accumulator and mask are scratch): charged the edge map once
        return {"front_nms": 4 * self.w * self.h, "ppht_prep": self.w * self.h, "ppht": self.w * self.h}

    def cv2_chain(self, cv, frame):  # opencv.py:213-217
        gray = cv.cvtColor(frame, cv.COLOR_BGR2GRAY)
        edges = cv.Canny(gray, 5, 150, apertureSize=3)
        return cv.HoughLinesP(edges, 1, np.pi / 180, threshold=30, minLineLength=50, maxLineGap=10)

    def cv2_stages(self, cv, frame):
        t = [time.perf_counter()]
        gray = cv.cvtColor(frame, cv.COLOR_BGR2GRAY); t.append(time.perf_counter())
        edges = cv.Canny(gray, 5, 150, apertureSize=3); t.append(time.perf_counter())
        cv.HoughLinesP(edges, 1, np.pi / 180, threshold=30, minLineLength=50, maxLineGap=10); t.append(time.perf_counter())
        return list(zip(("cvtColor", "Canny", "HoughLinesP"), np.diff(t)))


def _cv_lsd(cv, gray):  # opencv.py:206-207 with lsd()'s defaults
    det = cv.createLineSegmentDetector(0, scale=0.8, sigma_scale=0.6, quant=2.0, ang_th=22.5, log_eps=0.0,
                                       density_th=0.7, n_bins=1024)
    return det.detect(gray)[0]


class Lsd(Workload):
    name = "lsd_1080p_b512"
    metric = "1080p frames/s LSD, refine 0 / scale .8 (batch of 512 synthetic voxel-world frames per GPU)"
    h, w, batch, unique, kind = 1080, 1920, 512, 32, "voxel"
    chain = {"lsd": {"refine": 0, "scale": 0.8, "sigma_scale": 0.6, "quant": 2.0, "ang_th": 22.5, "log_eps": 0.0,
                     "density_th": 0.7, "n_bins": 1024}}

    def chain_bytes(self):  # 3WH (+ 16 B per segment)
        return 3 * self.w * self.h

    def kernel_bytes(self):
        return {"lsd_prep": self.w * self.h, "lsd_grow": 0, "misc": 4 * self.w * self.h}

    def cv2_chain(self, cv, frame):
        return _cv_lsd(cv, cv.cvtColor(frame, cv.COLOR_BGR2GRAY))

    def cv2_stages(self, cv, frame):
        t = [time.perf_counter()]
        gray = cv.cvtColor(frame, cv.COLOR_BGR2GRAY); t.append(time.perf_counter())
        _cv_lsd(cv, gray); t.append(time.perf_counter())
        return list(zip(("cvtColor", "LSD.detect"), np.diff(t)))


class LinesVp(Workload):
    name = "lines_1080p_cave_b128"
    metric = "1080p frames/s --pipeline lines: LSD + vanishing-point fit (batch of 128 synthetic cave-style frames per GPU)"
    h, w, batch, unique, kind = 1080, 1920, 128, 16, "cave"
    chain = {"lsd": Lsd.chain["lsd"], "min_segment_px": 10, "regress_lines": [1000, 500, 0.3],
             "focal_frame": "frame size (the reference hard-codes its 600x400 fixture size)"}
    ref_step_frames_per_core = 1
    needs_processes = True
    n_az, n_el = 360, 90

    def chain_bytes(self):  # LSD bytes + 4 n_az n_el histogram out
        return 3 * self.w * self.h + 4 * self.n_az * self.n_el

    def kernel_bytes(self):
        return {"lsd_prep": self.w * self.h, "lsd_grow": 0, "misc": 4 * self.w * self.h}

    def cv2_chain(self, cv, frame, seed=None, fast=False):  # opencv_fit_color.py:85-93 via oracle/vp_ref.py
        from oracle import vp_ref
        segs = _cv_lsd(cv, cv.cvtColor(frame, cv.COLOR_BGR2GRAY))
        pairs = [] if segs is None else [(s[0, :2], s[0, 2:]) for s in segs]
        lines = vp_ref.edges_to_polar_lines(pairs, min_len=10)
        if len(lines) == 0:
            return segs, lines, ((0, 0), 100000)
        fit = vp_ref.regress_lines(lines, 1000, 500, 0.3, seed=seed, width=self.w, height=self.h,
                                   loss=vp_ref.sum_loss_vec if fast else vp_ref.sum_loss)
        return segs, lines, fit

    def cv2_stages(self, cv, frame):
        from oracle import vp_ref
        t = [time.perf_counter()]
        gray = cv.cvtColor(frame, cv.COLOR_BGR2GRAY); t.append(time.perf_counter())
        segs = _cv_lsd(cv, gray); t.append(time.perf_counter())
        pairs = [] if segs is None else [(s[0, :2], s[0, 2:]) for s in segs]
        lines = vp_ref.edges_to_polar_lines(pairs, min_len=10); t.append(time.perf_counter())
        if len(lines):
            vp_ref.regress_lines(lines, 1000, 500, 0.3, seed=0, width=self.w, height=self.h)
        t.append(time.perf_counter())
        return list(zip(("cvtColor", "LSD.detect", "polar_lines", "regress_lines (Python)"), np.diff(t)))


WORKLOADS = {c.name: c for c in (CannyHough, CannyHoughP, Lsd, LinesVp)}


# ----------------------------------------------------------------------------------------------- helpers
def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples nvidia-smi SM clocks and throttle reasons while the timed region runs.  ONE sampler per job (rank 0)
    queries every GPU of the job in one call every 200 ms: a sampler per rank put N times as many nvidia-smi
    processes on the driver, which showed up as a step time that grew with N."""

    def __init__(self, indices):
        super().__init__(daemon=True)
        self.ids, self.rows, self._stop_evt = ",".join(str(i) for i in indices), [], threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.ids}", f"--query-gpu={q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
                if out.returncode == 0:
                    for ln in out.stdout.strip().splitlines():
                        if ln.strip():
                            self.rows.append([c.strip() for c in ln.split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def finish(self):
        self._stop_evt.set()
        self.join(6)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_min_mhz": sm[0] if sm else None,
                "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons, "samples": len(self.rows),
                "gpus_sampled": self.ids}


def load_synth():
    """The synthetic generator is importable without loading the CUDA library."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("graphics_project_b200_synth",
                                                  os.path.join(ROOT, "graphics-project_b200", "synth.py"))
    synth = importlib.util.module_from_spec(spec)
    sys.modules["graphics_project_b200_synth"] = synth
    spec.loader.exec_module(synth)
    return synth


# process-pool worker of the `lines` reference chain (pure-Python search: threads would serialise on the GIL)
_WL = None


def _pool_init(name):
    global _WL
    import cv2 as cv
    cv.setNumThreads(1)
    _WL = WORKLOADS[name]()


def _pool_chain(frame):
    import cv2 as cv
    r = _WL.cv2_chain(cv, frame)
    return r[2] if isinstance(r, tuple) else (0 if r is None else len(r))


class CpuPool:
    """Frame-parallel pool of os.cpu_count() workers with cv2.setNumThreads(1) each: HoughLines, HoughLinesP and
    LSD are single-threaded inside cv2, so this is the strongest honest CPU figure.  Threads (cv2 releases the
    GIL) unless the chain runs Python code, then processes."""

    def __init__(self, wl):
        import cv2 as cv
        from concurrent.futures import ThreadPoolExecutor, ProcessPoolExecutor
        self.cv, self.wl = cv, wl
        self.cores = os.cpu_count() or 1
        cv.setNumThreads(1)
        if wl.needs_processes:
            self.ex = ProcessPoolExecutor(self.cores, initializer=_pool_init, initargs=(wl.name,))
            self.fn = _pool_chain
        else:
            self.ex = ThreadPoolExecutor(self.cores)
            self.fn = lambda f: wl.cv2_chain(cv, f)
        self.kind = "processes" if wl.needs_processes else "threads"

    def run(self, frames):
        return list(self.ex.map(self.fn, frames))

    def close(self):
        self.ex.shutdown()


def time_cpu_reference(wl, frames, min_seconds=6.0):
    """Repeats the sample until ``min_seconds`` of wall time.  Returns (frames/s, cores, seconds, frames, kind)."""
    pool = CpuPool(wl)
    pool.run(frames[:pool.cores])  # warm
    done, t0 = 0, time.perf_counter()
    while True:
        pool.run(frames)
        done += len(frames)
        dt = time.perf_counter() - t0
        if dt >= min_seconds:
            break
    pool.close()
    return done / dt, pool.cores, dt, done, pool.kind


def cpu_detail(wl, frames, budget_s=8.0):
    """BASELINE.md section 3 / SURVEY 8(d): per-stage times (median over the sample, one frame at a time) and the
    end-to-end chain frame by frame with cv2.setNumThreads(os.cpu_count()) -- best of 5 passes and median."""
    import cv2 as cv
    cores = os.cpu_count() or 1
    cv.setNumThreads(cores)
    stages = {}
    t_end = time.perf_counter() + budget_s / 2
    for f in frames:
        for name, s in wl.cv2_stages(cv, f):
            stages.setdefault(name, []).append(s * 1e3)
        if time.perf_counter() > t_end:
            break
    passes = []
    n = min(len(frames), max(1, len(next(iter(stages.values())))))
    t_end = time.perf_counter() + budget_s / 2
    for _ in range(5):
        t0 = time.perf_counter()
        for f in frames[:n]:
            wl.cv2_chain(cv, f)
        passes.append(n / (time.perf_counter() - t0))
        if time.perf_counter() > t_end:
            break
    cv.setNumThreads(1)
    return {"stages_ms_median": {k: round(statistics.median(v), 3) for k, v in stages.items()},
            "stage_sample_frames": len(next(iter(stages.values()))),
            "setNumThreads_all": {"threads": cores, "frames_per_pass": n, "passes": len(passes),
                                  "fps_best": max(passes), "fps_median": statistics.median(passes)}}


def make_frames(synth, wl, n, rank):
    uniq = synth.make_batch(min(wl.unique, n), wl.h, wl.w, wl.kind, first_seed=rank * 1000)
    return [uniq[i % len(uniq)] for i in range(n)], uniq


def run_reference(args, wl, rank):
    """--impl reference: the reference's own CPU implementation (the cv2 wheel of this image; the `lines` workload
    adds oracle/vp_ref.py, the restatement of the reference's pure-Python search)."""
    if rank != 0:
        return
    synth = load_synth()
    pool = CpuPool(wl)
    sample = max(1, wl.ref_step_frames_per_core * pool.cores)
    frames, _ = make_frames(synth, wl, sample, 0)
    import cv2 as cv
    for _ in range(args.warmup):
        pool.run(frames)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        pool.run(frames)
    dt = time.perf_counter() - t0
    pool.close()
    fps = sample * args.steps / dt
    line = {"impl": "reference", "metric": wl.metric, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": wl.config(),
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": pool.cores, "kind": "reference",
                             "sample": f"{sample} frames of the workload per step, cv2 {cv.__version__}, "
                                       f"frame-parallel pool of {pool.cores} {pool.kind}, setNumThreads(1)"},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------- engine arm
def seg_match(a, b, tol):
    """LSD criterion: same count, endpoints within ``tol`` px in order (either orientation)."""
    if len(a) != len(b):
        return False
    if len(a) == 0:
        return True
    d0 = np.abs(a - b).max(1)
    d1 = np.abs(a - b[:, [2, 3, 0, 1]]).max(1)
    return bool((np.minimum(d0, d1) <= tol).all())


class EngineRun:
    """Device-resident state of one workload on one GPU: buffers, the step, results, the host pipeline, parity."""

    def __init__(self, wl, gp, torch, dev, frames, B):
        self.wl, self.gp, self.torch, self.dev, self.frames, self.B = wl, gp, torch, dev, frames, B
        self.eng = gp.Engine(dev.index, wl.h, wl.w, B)
        self.i = 0
        name = wl.name
        mk = lambda *s, dt: torch.empty(s, dtype=dt, device=dev)  # noqa: E731
        if name == CannyHough.name:
            self.max_k = 1024
            na, nr = self.eng.hough_dims(wl.h, wl.w)
            self.edges = mk(B, wl.h, wl.w, dt=torch.uint8)
            self.accum = mk(B, na + 2, nr + 2, dt=torch.int32)
            # two result sets: with N > 1 the gather of step i runs beside the kernels of step i + 1
            self.payload2 = [mk(B, self.max_k, 2, dt=torch.float32) for _ in range(2)]
            self.counts2 = [mk(B, dt=torch.int32) for _ in range(2)]
        elif name == CannyHoughP.name:
            self.max_k = 1024
            self.edges = mk(B, wl.h, wl.w, dt=torch.uint8)
        elif name == Lsd.name:
            self.max_k = 2048
            self.gray = mk(B, wl.h, wl.w, dt=torch.uint8)
        else:
            self.max_k = 2048
        self.payload = self.counts = self.extra = None

    # one pass of the chain over the resident batch, on the current stream
    def step(self):
        wl, eng, name = self.wl, self.eng, self.wl.name
        if name == CannyHough.name:
            self.payload, self.counts = self.payload2[self.i & 1], self.counts2[self.i & 1]
            eng.front_canny(self.frames, wl.LO, wl.HI, blur=True, out=self.edges)
            eng.hough_lines(None, 1.0, np.pi / 180, wl.THR, out=(self.payload, None, self.counts, self.accum))
        elif name == CannyHoughP.name:
            eng.front_canny(self.frames, 5, 150, out=self.edges)
            self.payload, self.counts = eng.hough_lines_p(self.edges, 1, np.pi / 180, 30, 50, 10, max_segs=self.max_k)
        elif name == Lsd.name:
            gray = eng.bgr2gray(self.frames)
            self.payload, _, _, self.counts = eng.lsd(gray, max_segs=self.max_k)
        else:
            fit, hist, polar, pcounts, segs, counts = self.gp.vp.get_camera_angles_batch(
                self.frames, 1000, 500, max_segs=self.max_k, engine=eng)
            self.payload, self.counts, self.extra = fit, pcounts, (hist, polar, segs, counts)
        self.i += 1

    def gather_args(self):
        return self.counts, self.payload

    MAX_K = {CannyHough.name: 1024, CannyHoughP.name: 1024, Lsd.name: 2048, LinesVp.name: 2048}

    @staticmethod
    def pipeline(wl, gp, torch, dev, B, chunk=None):
        """(host pipeline, pinned output tensors, class name) of the workload's chain."""
        P, d = gp.pipeline, dev.index
        pin = lambda *s, dt: torch.empty(s, dtype=dt, pin_memory=True)  # noqa: E731
        k = EngineRun.MAX_K[wl.name]
        if wl.name == CannyHough.name:
            return (P.CannyHoughHost(wl.h, wl.w, d, chunk=chunk or 32, max_lines=k, lo=wl.LO, hi=wl.HI, blur=True,
                                     threshold=wl.THR),
                    [pin(B, k, 2, dt=torch.float32), pin(B, dt=torch.int32)], "CannyHoughHost")
        if wl.name == CannyHoughP.name:
            return (P.ProbHost(wl.h, wl.w, d, chunk=chunk or 16, max_segs=k),
                    [pin(B, k, 4, dt=torch.int32), pin(B, dt=torch.int32)], "ProbHost")
        if wl.name == Lsd.name:
            return (P.LsdHost(wl.h, wl.w, d, chunk=chunk or 64, max_segs=k),
                    [pin(B, k, 4, dt=torch.float32), pin(B, dt=torch.int32)], "LsdHost")
        return (P.LinesVpHost(wl.h, wl.w, d, chunk=chunk or 32, max_segs=k),
                [pin(B, 3, dt=torch.float64), pin(B, dt=torch.int32)], "LinesVpHost")

    def parity_sample(self, host_frames, idxs):
        """Frames of the TIMED batch (results of the last timed step still in the buffers) against cv2 on the same
        host frames: bit-exact for Canny / HoughLines / HoughLinesP, 0.5 px for LSD; the VP fit is graded by the
        reference's own loss against three seeded runs of the reference's search."""
        import cv2 as cv
        wl, name = self.wl, self.wl.name
        cv.setNumThreads(os.cpu_count() or 1)
        ok, detail = True, []
        counts = self.counts.cpu().numpy()
        for i in idxs:
            f = host_frames[i].numpy()
            if name == CannyHough.name:
                gray = cv.cvtColor(f, cv.COLOR_BGR2GRAY)
                canny = cv.Canny(cv.GaussianBlur(gray, (5, 5), 0), wl.LO, wl.HI)
                ref = cv.HoughLines(canny, 1, np.pi / 180, wl.THR, None, 0, 0)
                n = 0 if ref is None else len(ref)
                e = np.array_equal(self.edges[i].cpu().numpy(), canny) and int(counts[i]) == n and (
                    n == 0 or np.array_equal(self.payload[i, :min(n, self.max_k)].cpu().numpy(), ref[:self.max_k, 0]))
            elif name == CannyHoughP.name:
                edges = cv.Canny(cv.cvtColor(f, cv.COLOR_BGR2GRAY), 5, 150, apertureSize=3)
                ref = cv.HoughLinesP(edges, 1, np.pi / 180, threshold=30, minLineLength=50, maxLineGap=10)
                n = 0 if ref is None else len(ref)
                e = np.array_equal(self.edges[i].cpu().numpy(), edges) and int(counts[i]) == n and (
                    n == 0 or np.array_equal(self.payload[i, :min(n, self.max_k)].cpu().numpy(), ref[:self.max_k, 0]))
            elif name == Lsd.name:
                ref = _cv_lsd(cv, cv.cvtColor(f, cv.COLOR_BGR2GRAY))
                n = 0 if ref is None else len(ref)
                e = int(counts[i]) == n and (n == 0 or seg_match(self.payload[i, :n].cpu().numpy(), ref[:, 0], 0.5))
            else:
                from oracle import vp_ref
                hist, polar, segs, scounts = self.extra
                ref = _cv_lsd(cv, cv.cvtColor(f, cv.COLOR_BGR2GRAY))
                n = 0 if ref is None else len(ref)
                e = int(scounts[i]) == n and (n == 0 or seg_match(segs[i, :n].cpu().numpy(), ref[:, 0], 0.5))
                lines = polar[i, :int(counts[i])].cpu().numpy()
                if e and len(lines):
                    phi, theta, loss = self.payload[i].cpu().numpy()
                    rescored = vp_ref.sum_loss_vec(phi, theta, lines, wl.w, wl.h)
                    runs = sorted(vp_ref.regress_lines(lines, 1000, 500, 0.3, seed=s, width=wl.w, height=wl.h,
                                                       loss=vp_ref.sum_loss_vec)[1] for s in range(3))
                    e = abs(rescored - loss) <= 1e-9 * max(1.0, abs(loss)) and loss <= runs[1]
                    detail.append({"frame": int(i), "lines": len(lines), "loss": float(loss),
                                   "reference_search_losses": [float(r) for r in runs]})
            ok = ok and bool(e)
        cv.setNumThreads(1)
        crit = {CannyHough.name: "edge map, line count, (rho, theta) and order bit-exact vs cv2",
                CannyHoughP.name: "edge map, segment count, endpoints and order bit-exact vs cv2",
                Lsd.name: "segment count equal, endpoints within 0.5 px in cv2's order",
                LinesVp.name: "LSD as lsd_1080p_b512; fit loss re-scored by the reference's sum_loss and <= the "
                              "median of 3 seeded runs of the reference's search"}[name]
        out = {"frames": len(idxs), "frame_indices": [int(i) for i in idxs], "equal": ok, "criterion": crit,
               "cv2": cv.__version__}
        if detail:
            out["fits"] = detail
        return out


def run_engine(args, wl, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import graphics_project_b200 as gp

    assert torch.cuda.is_available(), "bench.py needs a GPU (the engine has no CPU fallback)"
    synth = load_synth()
    torch.cuda.set_device(local_rank)
    affinity0 = os.sched_getaffinity(0)  # the CPU leg below gets every core back
    numa = gp.pipeline.bind_to_gpu_numa_node(local_rank) if not args.no_numa else {"bound": False, "skipped": True}
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    B = args.batch or wl.batch

    # ---- synthetic input: frames are generated on the host once and uploaded before timing
    # The job's batch is the workload's `unique` distinct frames tiled cyclically over all B x N slots (generation
    # cost only), so every GPU holds the same frames and the per-GPU work is exactly fixed as N grows (weak scaling).
    # --rank-seeds gives every rank its own frames instead (seed + 1000 * rank): step times then differ by up to
    # 5 % with the content and the job runs at the slowest rank's pace (profiles/r2_seed_rank_step_times.txt).
    seed_rank = args.seed_rank if args.seed_rank >= 0 else (rank if args.rank_seeds else 0)
    _, uniq = make_frames(synth, wl, min(wl.unique, B), seed_rank)
    host_frames = torch.empty((B, wl.h, wl.w, 3), dtype=torch.uint8, pin_memory=True)
    for i in range(B):
        host_frames[i] = torch.from_numpy(uniq[i % len(uniq)])
    frames = host_frames.to(dev)
    run = EngineRun(wl, gp, torch, dev, frames, B)
    eng = run.eng
    state = {"pending": None, "gathered": None}

    no_gather = os.environ.get("LE_BENCH_NO_GATHER") == "1"  # experiments only (scripts/nccl_channels.sh)

    def step():
        run.step()
        if world > 1 and not no_gather:  # gather the per-frame results of the whole job, one step behind
            if state["pending"] is not None:
                state["gathered"] = state["pending"].wait()
            c, p = run.gather_args()
            state["pending"] = gp.dist.gather_results(c, p, B * world, async_op=True, slot=run.i & 1)

    def drain():
        if state["pending"] is not None:
            state["gathered"] = state["pending"].wait()
            state["pending"] = None

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    drain()
    barrier()
    eng.check()
    # one process samples every GPU of the job (LE_BENCH_NO_SAMPLER=1: experiments only)
    sampler = ClockSampler(range(world)) if rank == 0 and os.environ.get("LE_BENCH_NO_SAMPLER") != "1" else None
    if sampler:
        sampler.start()
    t_load = time.perf_counter()
    while time.perf_counter() - t_load < 1.0:  # keep the GPU under load while the clock sampler collects
        step()
        drain()
        torch.cuda.synchronize()
    launches0 = eng.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        step()
    drain()  # the last step's gather is inside the timed region
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = eng.launches - launches0
    clocks = sampler.finish() if sampler else None
    eng.check()
    ms_ranks = [ms]
    if world > 1:
        # every rank's own time (ranks draw different frames, so they differ by a few per cent); the step time of
        # the job is the slowest rank's
        tl = torch.zeros((world,), dtype=torch.float64, device=dev)
        tl[rank] = ms
        dist.all_reduce(tl)
        ms_ranks = [float(v) for v in tl.tolist()]
        ms = max(ms_ranks)
        lt = torch.tensor([launches], dtype=torch.int64, device=dev)
        dist.all_reduce(lt)
        launches = int(lt.item())
    fps = B * world * args.steps / (ms * 1e-3)
    want_counts = run.counts.cpu()
    results_mean = float(want_counts.float().mean().item())

    # ---- parity: frames of the timed batch against cv2 (rank 0; before anything overwrites the result buffers)
    parity = None
    if rank == 0 and not args.no_parity:
        idxs = sorted({0, 1, B // 2, B - 1} if B >= 4 else set(range(B)))
        parity = run.parity_sample(host_frames, idxs)

    # ---- roofline of the dominant kernel: per-launch durations from CUDA events the library records around every
    # launch (on its launch stream), in a separate pass of K more steps of the same workload on the same stream
    # (eight event records per step cost ~22 us, which is 2.5 % of the default workload's step)
    peak, peak_src = hbm_peak()
    eng.timing(True)
    eng.timing_read()
    for _ in range(max(1, min(args.steps, 10))):
        run.step()
    torch.cuda.synchronize()
    kt = eng.timing_read()
    eng.timing(False)
    kernel_ms = {k: v[0] / v[1] for k, v in kt.items()}
    kbytes = wl.kernel_bytes()
    dominant = max(kt, key=lambda k: kt[k][0])
    alg = kbytes.get(dominant, 0) * B
    achieved = alg / (kernel_ms[dominant] * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get(wl.name, {}).get(dominant)
    chain_frac = wl.chain_bytes() * B / (ms / args.steps * 1e-3) / 1e9 / peak
    roofline = {"bound": "hbm", "kernel": dominant, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg,
                "kernel_ms_per_launch": kernel_ms,
                "kernel_share_of_step": {k: v[0] / sum(x[0] for x in kt.values()) for k, v in kt.items()},
                "chain_algorithmic_bytes_per_frame": wl.chain_bytes(),
                "step_frac_of_hbm_roofline": chain_frac}

    # ---- e2e: host-buffer API, H2D + D2H inside the timed region
    e2e = None
    if not args.no_e2e:
        del run
        torch.cuda.empty_cache()
        pipe, outs, api = EngineRun.pipeline(wl, gp, torch, dev, B, args.chunk)
        outs2 = [torch.empty_like(o).pin_memory() for o in outs]  # the batches alternate two sets of host outputs
        for _ in range(2):
            pipe.run(host_frames, *outs)
        barrier()
        l0 = pipe.launches
        e2e_steps = max(2, min(args.steps, 5))
        # every step: H2D of the whole batch from pinned memory, the chain, D2H of its results.  Step i + 1 is
        # enqueued before step i is waited for (two output sets), as an application that streams batches would
        t0 = time.perf_counter()
        prev = None
        for si in range(e2e_steps):
            cur = pipe.run(host_frames, *(outs if si % 2 == 0 else outs2), wait=False)
            if prev is not None:
                prev.wait()
            prev = cur
        prev.wait()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        e2e_launches = pipe.launches - l0
        pipe.run(host_frames, *outs)  # (the comparison below reads the first set)
        # copy-only ceiling: the same pinned batch through the same chunks and streams, no kernels; the ranks start
        # together (barrier) and copy as many steps as the e2e leg ran, so every rank sees the others' traffic
        barrier()
        ceil_steps = e2e_steps
        t0 = time.perf_counter()
        for _ in range(ceil_steps):
            for ci, c0 in enumerate(range(0, B, pipe.chunk)):
                c1 = min(B, c0 + pipe.chunk)
                si = ci % len(pipe.streams)
                with torch.cuda.stream(pipe.streams[si]):
                    pipe.bufs[si]["src"][:c1 - c0].copy_(host_frames[c0:c1], non_blocking=True)
            torch.cuda.synchronize()
        dt_copy = (time.perf_counter() - t0) / ceil_steps
        if world > 1:
            t = torch.tensor([dt, dt_copy], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt, dt_copy = float(t[0].item()), float(t[1].item())
        pipe.check()
        assert torch.equal(outs[1], want_counts), "e2e results differ from the resident-input run"
        e2e = {"value": B * world * e2e_steps / dt, "unit": "frames/s", "h2d_bytes_per_step": pipe.h2d_bytes,
               "d2h_bytes_per_step": pipe.d2h_bytes, "steps": e2e_steps, "gpu_launches": e2e_launches,
               "batches_in_flight": 2,
               "copy_ceiling_fps": B * world / dt_copy,
               "copy_ceiling_gbs_per_gpu": pipe.h2d_bytes / dt_copy / 1e9,
               "api": f"graphics_project_b200.pipeline.{api}.run (pinned host frames -> host results)",
               "numa": numa}

    # ---- CPU baseline: the reference's cv2 chain on a bounded sample, rank 0 at N == 1 only
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        import cv2 as cv
        os.sched_setaffinity(0, affinity0)
        sample = max(1, min(128, wl.ref_step_frames_per_core * (os.cpu_count() or 1) * 2))
        fl = [uniq[i % len(uniq)] for i in range(sample)]
        fps_cpu, cores, dt, done, kind = time_cpu_reference(wl, fl)
        cpu = {"value": fps_cpu, "unit": "frames/s", "cores": cores, "kind": "reference",
               "sample": f"{done} frames of the same workload ({sample}-frame sample repeated) in {dt:.1f} s wall = "
                         f"{dt * cores:.0f} core-seconds, cv2 {cv.__version__}, frame-parallel pool of {cores} "
                         f"{kind}, setNumThreads(1)" + (", + oracle/vp_ref.py (the reference's Python search)"
                                                        if wl.needs_processes else "")}
        cpu.update(cpu_detail(wl, fl[:max(2, min(len(fl), 8))]))

    if rank == 0:
        cfg = wl.config()
        cfg["batch_per_gpu"] = B
        if args.rank_seeds or args.seed_rank >= 0:
            cfg["frames"] = "per-rank seeds" if args.rank_seeds else f"seed rank {args.seed_rank}"
        line = {"metric": wl.metric, "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": cfg,
                "results_per_frame_mean": results_mean,
                "ms_per_step_ranks": [round(v / args.steps, 4) for v in ms_ranks],
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "parity_sample": parity,
                "gpu_launches": launches, "clocks": clocks}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--workload", default=CannyHough.name, choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="frames per GPU per step (default: the workload's)")
    ap.add_argument("--chunk", type=int, default=0, help="frames per chunk of the host pipeline (e2e)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-numa", action="store_true", help="do not bind the process to the GPU's NUMA node")
    ap.add_argument("--rank-seeds", action="store_true", help="every rank draws its own frames (seed + 1000 * rank)")
    ap.add_argument("--seed-rank", type=int, default=-1, help="generate the frames rank R of a --rank-seeds run gets")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    wl = WORKLOADS[args.workload]()
    if args.impl == "reference":
        return run_reference(args, wl, rank)
    return run_engine(args, wl, rank, world, local_rank)


if __name__ == "__main__":
    main()
