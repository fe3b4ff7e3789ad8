#!/usr/bin/env python
"""Benchmark of the line-detection hot path (BASELINE.json metric: 1080p frames/s, Canny + HoughLines).

    python bench.py --gpus N --steps K --warmup W            # this engine
    python bench.py --impl reference --gpus N ...              # the reference's cv2 CPU path

Workload (BASELINE.json configs[1]): a batch of 256 synthetic 1920x1080 cube-edge BGR frames per
GPU; one step = gray -> GaussianBlur(5,5) -> Canny(5,10) -> HoughLines(1, pi/180, 50) with the
full int32 accumulator written (the get_camera_angles 'hough' chain, opencv_fit_color.py:74-83).
Prints ONE JSON line (see the task contract): value = whole-job frames/s with inputs resident in
HBM; e2e = the same through the host-buffer API (pinned host frames, H2D + D2H inside the timed
region); roofline = the dominant kernel's algorithmic bytes / its live CUDA-event duration vs the
measured HBM peak; cpu_baseline = the reference's cv2 chain timed on this box's cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W = 1080, 1920
BATCH = 256
UNIQUE = 32  # distinct synthetic frames, tiled cyclically to the batch (generation cost only)
CANNY_LO, CANNY_HI, HOUGH_THR = 5, 10, 50
METRIC = "1080p frames/s Canny+HoughLines (batch of 256 synthetic cube-edge frames per GPU)"


def algorithmic_bytes(h, w):
    """compulsory HBM traffic per frame of the Canny+HoughLines chain."""
    numrho = 2 * (w + h) + 1
    front = 3 * w * h + w * h          # BGR in + edge map out        (front_nms kernel)
    accum = 4 * 182 * (numrho + 2)     # int32 accumulator out        (hough_vote kernel)
    return front, accum


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples nvidia-smi SM clocks and throttle reasons while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
                if out.returncode == 0 and out.stdout.strip():
                    self.rows.append([c.strip() for c in out.stdout.strip().split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.1)

    def finish(self):
        self._stop_evt.set()
        self.join(6)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]),
                "reasons": reasons, "samples": len(self.rows)}


def cv2_chain(cv, frame):
    """The reference's calls, verbatim (opencv_fit_color.py:74-83)."""
    gray = cv.cvtColor(frame, cv.COLOR_BGR2GRAY)
    blurred = cv.GaussianBlur(gray, (5, 5), 0)
    canny = cv.Canny(blurred, CANNY_LO, CANNY_HI)
    return cv.HoughLines(canny, 1, np.pi / 180, HOUGH_THR, None, 0, 0)


def time_cpu_reference(frames, min_seconds=6.0):
    """Frame-parallel pool of os.cpu_count() workers, cv2.setNumThreads(1) each (HoughLines is
    single-threaded inside cv2, so this is the strongest honest CPU figure).  Repeats the sample until
    ``min_seconds`` of wall time (= min_seconds * cores of CPU work).  Returns (frames/s, cores, s, n)."""
    import cv2 as cv
    from concurrent.futures import ThreadPoolExecutor
    cores = os.cpu_count() or 1
    cv.setNumThreads(1)
    done = 0
    with ThreadPoolExecutor(cores) as ex:
        list(ex.map(lambda f: cv2_chain(cv, f), frames[:cores]))  # warm
        t0 = time.perf_counter()
        while True:
            list(ex.map(lambda f: cv2_chain(cv, f), frames))
            done += len(frames)
            dt = time.perf_counter() - t0
            if dt >= min_seconds:
                break
    return done / dt, cores, dt, done


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation (cv2 4.13 wheel of this image)."""
    if rank != 0:
        return
    from graphics_project_b200_synth import make_batch
    sample = BATCH
    frames = list(make_batch(sample, H, W, unique=min(UNIQUE, sample)))
    import cv2 as cv
    from concurrent.futures import ThreadPoolExecutor
    cores = os.cpu_count() or 1
    cv.setNumThreads(1)
    with ThreadPoolExecutor(cores) as ex:
        for _ in range(args.warmup):
            list(ex.map(lambda f: cv2_chain(cv, f), frames))
        t0 = time.perf_counter()
        for _ in range(args.steps):
            list(ex.map(lambda f: cv2_chain(cv, f), frames))
        dt = time.perf_counter() - t0
    fps = sample * args.steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "canny_hough_1080p_b256", "height": H, "width": W, "batch_per_gpu": BATCH,
                       "step_sample_frames": sample, "canny": [CANNY_LO, CANNY_HI], "blur": "5x5",
                       "hough": [1, "pi/180", HOUGH_THR]},
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "reference",
                             "sample": f"{sample} of the {BATCH} frames per step, cv2 {cv.__version__}, "
                                       f"frame-parallel pool of {cores} threads, setNumThreads(1)"},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH, help="frames per GPU per step")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--streams", type=int, default=int(os.environ.get("LE_BENCH_STREAMS", "1")),
                    help="slices of the batch processed on separate CUDA streams (kernels of different slices overlap)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    # the synthetic generator is importable without loading the CUDA library
    import importlib.util
    spec = importlib.util.spec_from_file_location("graphics_project_b200_synth",
                                                  os.path.join(ROOT, "graphics-project_b200", "synth.py"))
    synth = importlib.util.module_from_spec(spec)
    sys.modules["graphics_project_b200_synth"] = synth
    spec.loader.exec_module(synth)

    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch
    import torch.distributed as dist
    import graphics_project_b200 as gp

    assert torch.cuda.is_available(), "bench.py needs a GPU (the engine has no CPU fallback)"
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    B = args.batch
    max_lines = 1024

    # ---- synthetic input: frames are generated on the host once and uploaded before timing
    uniq = synth.make_batch(min(UNIQUE, B), H, W, first_seed=rank * 1000)
    host_frames = torch.empty((B, H, W, 3), dtype=torch.uint8, pin_memory=True)
    for i in range(B):
        host_frames[i] = torch.from_numpy(uniq[i % len(uniq)])
    frames = host_frames.to(dev)

    S = max(1, min(args.streams, B))
    engs = [gp.Engine(local_rank, H, W, (B + S - 1) // S) for _ in range(S)]
    eng = engs[0]
    side = [torch.cuda.Stream(dev) for _ in range(S)]
    bounds = [(i * B // S, (i + 1) * B // S) for i in range(S)]
    na, nr = eng.hough_dims(H, W)
    edges = torch.empty((B, H, W), dtype=torch.uint8, device=dev)
    # two result sets: with N > 1 the gather of step i runs beside the kernels of step i + 1
    lines2 = [torch.empty((B, max_lines, 2), dtype=torch.float32, device=dev) for _ in range(2)]
    counts2 = [torch.empty((B,), dtype=torch.int32, device=dev) for _ in range(2)]
    lines, counts = lines2[0], counts2[0]
    state = {"i": 0, "pending": None, "gathered": None}
    accum = torch.empty((B, na + 2, nr + 2), dtype=torch.int32, device=dev)

    def step():
        lines, counts = lines2[state["i"] & 1], counts2[state["i"] & 1]
        if S == 1:
            eng.front_canny(frames, CANNY_LO, CANNY_HI, blur=True, out=edges)
            eng.hough_lines(None, 1.0, np.pi / 180, HOUGH_THR, out=(lines, None, counts, accum))
        else:
            main = torch.cuda.current_stream(dev)
            fork = torch.cuda.Event()
            fork.record(main)
            for e, st, (a, b) in zip(engs, side, bounds):
                st.wait_event(fork)
                with torch.cuda.stream(st):
                    e.front_canny(frames[a:b], CANNY_LO, CANNY_HI, blur=True, out=edges[a:b])
                    e.hough_lines(None, 1.0, np.pi / 180, HOUGH_THR, out=(lines[a:b], None, counts[a:b], accum[a:b]))
                join = torch.cuda.Event()
                join.record(st)
                main.wait_event(join)
        if world > 1:  # gather the per-frame results of the whole job, one step behind
            if state["pending"] is not None:
                state["gathered"] = state["pending"].wait()
            state["pending"] = gp.dist.gather_results(counts, lines, B * world, async_op=True, slot=state["i"] & 1)
        state["i"] += 1

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
    for e in engs:
        e.check()
    sampler = ClockSampler(local_rank)
    sampler.start()
    t_load = time.perf_counter()
    while time.perf_counter() - t_load < 1.0:  # keep the GPU under load while the clock sampler collects
        for _ in range(20):
            step()
        drain()
        torch.cuda.synchronize()
    for e in engs:
        # per-kernel events cost ~22 us per step (8 records): with one stream they are recorded in the separate
        # roofline pass below instead (same workload, same stream); LE_BENCH_EVENTS=1 puts them back here
        e.timing(S > 1 or os.environ.get("LE_BENCH_EVENTS", "0") == "1")
        e.timing_read()
    launches0 = sum(e.launches for e in engs)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        step()
    drain()  # the last step's gather is inside the timed region
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = sum(e.launches for e in engs) - launches0
    ktimes = {}
    for e in engs:
        for k, (tms, c) in e.timing_read().items():
            a = ktimes.get(k, (0.0, 0))
            ktimes[k] = (a[0] + tms, a[1] + c)
        e.timing(False)
    clocks = sampler.finish()
    for e in engs:
        e.check()
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        lt = torch.tensor([launches], dtype=torch.int64, device=dev)
        dist.all_reduce(lt)
        launches = int(lt.item())
    fps = B * world * args.steps / (ms * 1e-3)

    # ---- roofline of the dominant kernel.  With --streams > 1 the slices' kernels overlap on the GPU, so a
    # kernel's event-to-event duration in the timed region includes time it shares the SMs with the other
    # slice.  The per-launch duration used for the roofline is therefore taken live from K more steps of the
    # same workload issued on ONE stream (whole batch per launch, kernels back to back, nothing overlapping).
    front_b, accum_b = algorithmic_bytes(H, W)
    peak, peak_src = hbm_peak()
    kernel_ms_overlapped = {k: v[0] / v[1] for k, v in ktimes.items()}
    eng.timing(True)
    eng.timing_read()
    for _ in range(args.steps):
        eng.front_canny(frames, CANNY_LO, CANNY_HI, blur=True, out=edges)
        eng.hough_lines(None, 1.0, np.pi / 180, HOUGH_THR, out=(lines, None, counts, accum))
    torch.cuda.synchronize()
    kt1 = eng.timing_read()
    eng.timing(False)
    alg = {"front_nms": front_b * B, "hough_vote": accum_b * B}  # algorithmic bytes per launch (whole batch)
    kernel_ms = {k: v[0] / v[1] for k, v in kt1.items()}
    dominant = max((k for k in kernel_ms if k in alg), key=lambda k: kt1[k][0])
    achieved = alg[dominant] / (kernel_ms[dominant] * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get(dominant)
    roofline = {"bound": "hbm", "kernel": dominant, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg[dominant],
                "kernel_ms_per_launch": kernel_ms,
                "kernel_ms_per_launch_in_timed_region_overlapped_half_batches": kernel_ms_overlapped,
                "step_frac_of_hbm_roofline": (front_b + accum_b) * B / (ms / args.steps * 1e-3) / 1e9 / peak}

    # ---- e2e: host-buffer API, H2D + D2H inside the timed region
    e2e = None
    if not args.no_e2e:
        del accum, edges
        torch.cuda.empty_cache()
        pipe = gp.pipeline.CannyHoughHost(H, W, local_rank, chunk=32, max_lines=max_lines, lo=CANNY_LO, hi=CANNY_HI,
                                          blur=True, threshold=HOUGH_THR)
        out_lines = torch.empty((B, max_lines, 2), dtype=torch.float32, pin_memory=True)
        out_counts = torch.empty((B,), dtype=torch.int32, pin_memory=True)
        for _ in range(2):
            pipe.run(host_frames, out_lines, out_counts)
        barrier()
        l0 = pipe.launches
        t0 = time.perf_counter()
        e2e_steps = max(2, min(args.steps, 5))
        for _ in range(e2e_steps):
            pipe.run(host_frames, out_lines, out_counts)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        assert torch.equal(out_counts, counts.cpu()), "e2e results differ from the resident-input run"
        e2e = {"value": B * world * e2e_steps / dt, "unit": "frames/s", "h2d_bytes_per_step": pipe.h2d_bytes,
               "d2h_bytes_per_step": pipe.d2h_bytes, "steps": e2e_steps,
               "gpu_launches": pipe.launches - l0,
               "api": "graphics_project_b200.pipeline.CannyHoughHost.run (pinned host frames -> host lines)"}

    # ---- CPU baseline: the reference's cv2 chain on a bounded sample, rank 0 at N == 1 only
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        sample = 128
        fps_cpu, cores, dt, done = time_cpu_reference([uniq[i % len(uniq)] for i in range(sample)])
        import cv2 as cv
        cpu = {"value": fps_cpu, "unit": "frames/s", "cores": cores, "kind": "reference",
               "sample": f"{done} frames of the same workload ({sample}-frame sample repeated) in {dt:.1f} s wall = "
                         f"{dt * cores:.0f} core-seconds, cv2 {cv.__version__}, frame-parallel pool of {cores} "
                         f"threads, setNumThreads(1)"}

    if rank == 0:
        line = {"metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": {"workload": "canny_hough_1080p_b256", "height": H, "width": W, "batch_per_gpu": B,
                           "unique_frames": len(uniq), "canny": [CANNY_LO, CANNY_HI], "blur": "5x5",
                           "hough": [1, "pi/180", HOUGH_THR], "accumulator_written": True, "streams": S,
                           "l2": f"inputs larger than L2 ({frames.numel() / 1e6:.0f} MB BGR per step)",
                           "lines_per_frame_mean": float(counts.float().mean().item())},
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
