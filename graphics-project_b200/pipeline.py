"""Host-buffer pipelines: the call a user with frames in host memory makes.

Frames live in pinned host memory; the batch is cut into chunks that alternate between two CUDA
streams (each with its own engine context and device buffers) so the host->device copy of chunk
i+1 overlaps the kernels of chunk i, and the small per-frame results are copied back to pinned
host memory on the same stream.
"""
import numpy as np
import torch

from .engine import Engine


class CannyHoughHost:
    """BGR frames (pinned host, [N,H,W,3] u8) -> Canny -> HoughLines -> (lines [N,max,2] f32, counts [N])."""

    def __init__(self, h, w, device=0, chunk=32, max_lines=1024, lo=5, hi=10, blur=True, rho=1.0,
                 theta=np.pi / 180, threshold=50, keep_accum=True):
        self.device = torch.device("cuda", device)
        self.h, self.w, self.chunk, self.max_lines = h, w, chunk, max_lines
        self.params = (lo, hi, blur, rho, theta, threshold)
        self.streams = [torch.cuda.Stream(self.device) for _ in range(2)]
        self.engines = [Engine(device) for _ in range(2)]
        na, nr = Engine.hough_dims(h, w, rho, theta)
        mk = lambda *s, dt=torch.uint8: torch.empty(s, dtype=dt, device=self.device)  # noqa: E731
        self.bufs = []
        for _ in range(2):
            self.bufs.append(dict(
                src=mk(chunk, h, w, 3), edges=mk(chunk, h, w),
                lines=mk(chunk, max_lines, 2, dt=torch.float32), counts=mk(chunk, dt=torch.int32),
                accum=mk(chunk, na + 2, nr + 2, dt=torch.int32) if keep_accum else None))
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def run(self, frames, out_lines, out_counts):
        """frames / out_*: pinned host tensors.  Blocks until every result is in host memory."""
        lo, hi, blur, rho, theta, thr = self.params
        n = frames.shape[0]
        self.h2d_bytes = self.d2h_bytes = 0
        for ci, c0 in enumerate(range(0, n, self.chunk)):
            c1 = min(n, c0 + self.chunk)
            k = c1 - c0
            s, eng, b = self.streams[ci & 1], self.engines[ci & 1], self.bufs[ci & 1]
            with torch.cuda.stream(s):
                b["src"][:k].copy_(frames[c0:c1], non_blocking=True)
                eng.front_canny(b["src"][:k], lo, hi, blur=blur, out=b["edges"][:k])
                acc = b["accum"][:k] if b["accum"] is not None else None
                eng.hough_lines(None, rho, theta, thr, out=(b["lines"][:k], None, b["counts"][:k], acc))
                out_lines[c0:c1].copy_(b["lines"][:k], non_blocking=True)
                out_counts[c0:c1].copy_(b["counts"][:k], non_blocking=True)
            self.h2d_bytes += frames[c0:c1].numel()
            self.d2h_bytes += b["lines"][:k].numel() * 4 + k * 4
        for s in self.streams:
            s.synchronize()
        return out_lines, out_counts

    @property
    def launches(self):
        return sum(e.launches for e in self.engines)
