"""Host-buffer pipelines: the call a user with frames in host memory makes.

Frames live in pinned host memory; the batch is cut into chunks that alternate between two CUDA
streams (each with its own engine context and device buffers) so the host->device copy of chunk
i+1 overlaps the kernels of chunk i, and the small per-frame results are copied back to pinned
host memory on the same stream.  One class per reference chain:

  CannyHoughHost   gray -> blur5 -> Canny -> HoughLines        (get_camera_angles 'hough', opencv_fit_color.py:74-83)
  ProbHost         gray -> Canny(5,150) -> HoughLinesP         (prob, opencv.py:211-219)
  LsdHost          gray -> LSD                                  (lsd, opencv.py:204-209)
  LinesVpHost      gray -> LSD -> polar lines -> VP fit         (get_camera_angles 'lsd', opencv_fit_color.py:85-93)
"""
import os

import numpy as np
import torch

from .engine import Engine


def bind_to_gpu_numa_node(device_index):
    """Pin this process (and the pinned host buffers it allocates afterwards: first touch) to the CPUs of the NUMA
    node the GPU hangs off, read from sysfs.  Host->device copies from the far socket cross the inter-socket link
    and share it with every other rank.  Returns a dict describing what was done (for the bench line)."""
    info = {"numa_node": None, "cpus": None, "bound": False}
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:  # nvml prints an 8-digit domain, sysfs a 4-digit one
            bus = bus[4:]
        base = f"/sys/bus/pci/devices/{bus}"
        node = int(open(f"{base}/numa_node").read().strip())
        info["numa_node"] = node
        cpulist = open(f"{base}/local_cpulist").read().strip()
        cpus = set()
        for part in cpulist.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        info["cpus"] = len(cpus)
        if node >= 0 and cpus and cpus != allowed:
            os.sched_setaffinity(0, cpus)
            info["bound"] = True
    except Exception as e:  # no NVML / sysfs entry (containers): leave the affinity alone and say so
        info["error"] = f"{type(e).__name__}: {e}"
    return info


class Pending:
    """A batch enqueued with ``run(..., wait=False)``."""

    def __init__(self, events, outs):
        self.events, self.outs = events, outs

    def wait(self):
        for e in self.events:
            e.synchronize()
        return self.outs


class _ChunkedHost:
    """``streams`` CUDA streams, each with its own engine context and device buffers, take the chunks in turn;
    subclasses say what a chunk needs and does.  Two streams hide the copy behind the kernels of a bandwidth-bound
    chain; the chains whose kernels are bound by dependent latency inside a frame (HoughLinesP, LSD) take more, so
    that several chunks are in their kernels at once."""

    def __init__(self, h, w, device=0, chunk=32, streams=2):
        self.device = torch.device("cuda", device)
        self.h, self.w, self.chunk = h, w, chunk
        self.streams = [torch.cuda.Stream(self.device) for _ in range(streams)]
        self.engines = [Engine(device) for _ in range(streams)]
        self.bufs = [self._alloc(chunk) for _ in range(streams)]
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def _mk(self, *shape, dt=torch.uint8):
        return torch.empty(shape, dtype=dt, device=self.device)

    def _alloc(self, chunk):  # -> dict of device buffers, must hold "src" [chunk, h, w, 3] u8
        raise NotImplementedError

    def _chunk(self, eng, b, k):  # enqueue the chain on the current stream; -> list of device result tensors [k, ...]
        raise NotImplementedError

    def run(self, frames, *outs, wait=True):
        """frames [N,H,W,3] u8 and every tensor of ``outs``: pinned host tensors, filled in the order the chain's
        results are documented.  With ``wait`` (default) the call blocks until every result is in host memory.
        ``wait=False`` only enqueues the batch and returns a ``Pending`` whose ``wait()`` blocks for THIS batch: a
        caller that alternates two sets of output buffers can enqueue batch i + 1 before it waits for batch i, so
        the kernels of the last chunks of a batch run beside the first copies of the next one (device buffers are
        reused in stream order, which makes that safe)."""
        n = frames.shape[0]
        self.h2d_bytes = self.d2h_bytes = 0
        for ci, c0 in enumerate(range(0, n, self.chunk)):
            c1 = min(n, c0 + self.chunk)
            k = c1 - c0
            si = ci % len(self.streams)
            s, eng, b = self.streams[si], self.engines[si], self.bufs[si]
            with torch.cuda.stream(s):
                b["src"][:k].copy_(frames[c0:c1], non_blocking=True)
                res = self._chunk(eng, b, k)
                assert len(res) == len(outs)
                for o, r in zip(outs, res):
                    o[c0:c1].copy_(r, non_blocking=True)
                    self.d2h_bytes += r.numel() * r.element_size()
            self.h2d_bytes += frames[c0:c1].numel()
        events = []
        for s in self.streams:
            e = torch.cuda.Event()
            e.record(s)
            events.append(e)
        pending = Pending(events, outs)
        return pending.wait() if wait else pending

    @property
    def launches(self):
        return sum(e.launches for e in self.engines)

    def check(self):
        for e in self.engines:
            e.check()


class CannyHoughHost(_ChunkedHost):
    """BGR frames (pinned host, [N,H,W,3] u8) -> Canny -> HoughLines -> (lines [N,max,2] f32, counts [N])."""

    def __init__(self, h, w, device=0, chunk=32, max_lines=1024, lo=5, hi=10, blur=True, rho=1.0,
                 theta=np.pi / 180, threshold=50, keep_accum=True):
        self.max_lines = max_lines
        self.params = (lo, hi, blur, rho, theta, threshold)
        self.keep_accum = keep_accum
        super().__init__(h, w, device, chunk)

    def _alloc(self, chunk):
        na, nr = Engine.hough_dims(self.h, self.w, self.params[3], self.params[4])
        return dict(src=self._mk(chunk, self.h, self.w, 3), edges=self._mk(chunk, self.h, self.w),
                    lines=self._mk(chunk, self.max_lines, 2, dt=torch.float32), counts=self._mk(chunk, dt=torch.int32),
                    accum=self._mk(chunk, na + 2, nr + 2, dt=torch.int32) if self.keep_accum else None)

    def _chunk(self, eng, b, k):
        lo, hi, blur, rho, theta, thr = self.params
        eng.front_canny(b["src"][:k], lo, hi, blur=blur, out=b["edges"][:k])
        acc = b["accum"][:k] if b["accum"] is not None else None
        eng.hough_lines(None, rho, theta, thr, out=(b["lines"][:k], None, b["counts"][:k], acc))
        return [b["lines"][:k], b["counts"][:k]]


class ProbHost(_ChunkedHost):
    """``prob`` (opencv.py:211-219) for a batch: BGR -> gray -> Canny(5,150) -> HoughLinesP(1, pi/180, 30, 50, 10)
    -> (segs [N,max,4] i32 in cv2's order, counts [N])."""

    def __init__(self, h, w, device=0, chunk=16, max_segs=1024, lo=5, hi=150, rho=1.0, theta=np.pi / 180,
                 threshold=30, min_line_length=50, max_line_gap=10, streams=4):
        self.max_segs = max_segs
        self.params = (lo, hi, rho, theta, threshold, min_line_length, max_line_gap)
        super().__init__(h, w, device, chunk, streams)

    def _alloc(self, chunk):
        return dict(src=self._mk(chunk, self.h, self.w, 3), edges=self._mk(chunk, self.h, self.w))

    def _chunk(self, eng, b, k):
        lo, hi, rho, theta, thr, mll, mlg = self.params
        eng.front_canny(b["src"][:k], lo, hi, out=b["edges"][:k])
        segs, counts = eng.hough_lines_p(b["edges"][:k], rho, theta, thr, mll, mlg, max_segs=self.max_segs)
        return [segs, counts]


class LsdHost(_ChunkedHost):
    """``lsd`` (opencv.py:204-209) for a batch: BGR -> gray -> LSD -> (segs [N,max,4] f32, counts [N])."""

    def __init__(self, h, w, device=0, chunk=64, max_segs=2048, streams=3, **lsd_params):
        self.max_segs = max_segs
        self.lsd_params = lsd_params
        super().__init__(h, w, device, chunk, streams)

    def _alloc(self, chunk):
        return dict(src=self._mk(chunk, self.h, self.w, 3))

    def _chunk(self, eng, b, k):
        gray = eng.bgr2gray(b["src"][:k])
        segs, _, _, counts = eng.lsd(gray, max_segs=self.max_segs, **self.lsd_params)
        return [segs, counts]


class LinesVpHost(_ChunkedHost):
    """``get_camera_angles(image, method='lsd')`` (opencv_fit_color.py:85-93) for a batch: BGR -> gray -> LSD ->
    segments longer than 10 px as polar lines -> fit -> (fit [N,3] f64 = (phi, theta, loss), pcounts [N])."""

    def __init__(self, h, w, device=0, chunk=32, max_segs=2048, iterations=1000, refinement_iterations=500,
                 width=None, height=None, streams=4):
        self.max_segs = max_segs
        self.search = (iterations, refinement_iterations)
        self.size = (float(w if width is None else width), float(h if height is None else height))
        super().__init__(h, w, device, chunk, streams)

    def _alloc(self, chunk):
        return dict(src=self._mk(chunk, self.h, self.w, 3))

    def _chunk(self, eng, b, k):
        from . import vp
        fit, _, _, pcounts, _, _ = vp.get_camera_angles_batch(b["src"][:k], self.search[0], self.search[1],
                                                              width=self.size[0], height=self.size[1],
                                                              max_segs=self.max_segs, engine=eng)
        return [fit, pcounts]
