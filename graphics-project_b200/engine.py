"""Batched, stream-asynchronous front door of the line engine (CUDA tensors in, CUDA tensors out).

One ``Engine`` per (process, device).  torch supplies device memory, the current CUDA stream and
(in ``dist.py``) the process group; every pixel of arithmetic happens in liblineengine.so.
Variable-length results come back padded ``[N, max_k, ...]`` together with ``counts[N]``.
"""
import ctypes as C
import math

import numpy as np
import torch

from . import _native
from ._native import lib


class EngineError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{_native.ERRORS.get(code, code)}: {msg}")
        self.code = code


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


class Engine:
    """Owns one ``le_ctx``.  Methods enqueue work on ``torch.cuda.current_stream(device)``."""

    def __init__(self, device=None, max_h=0, max_w=0, max_batch=0):
        if not torch.cuda.is_available():
            raise RuntimeError("graphics_project_b200.Engine needs a CUDA device (no CPU fallback)")
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else
                                   (device if isinstance(device, int) else torch.device(device).index or 0))
        ctx = C.c_void_p()
        rc = lib.le_create(C.byref(ctx), self.device.index, max_h, max_w, max_batch)
        if rc != 0:
            raise EngineError(rc, lib.le_last_error(None).decode())
        self._ctx = ctx

    def close(self):
        if getattr(self, "_ctx", None):
            lib.le_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ helpers
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _ok(self, rc):
        if rc != 0:
            raise EngineError(rc, lib.le_last_error(self._ctx).decode())

    def _img(self, t, channels):
        if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.uint8):
            raise TypeError("expected a CUDA uint8 tensor")
        if channels == 3:
            if t.dim() == 3:
                t = t[None]
            if t.dim() != 4 or t.shape[-1] != 3:
                raise ValueError("expected [N,H,W,3]")
        else:
            if t.dim() == 2:
                t = t[None]
            if t.dim() != 3:
                raise ValueError("expected [N,H,W]")
        return t.contiguous()

    def _out(self, t, shape, dtype, name):
        """A caller-provided output goes to the kernels as a raw pointer: it must be exactly what they write."""
        if t is None:
            return None
        if not (isinstance(t, torch.Tensor) and t.is_cuda and t.device == self.device):
            raise TypeError(f"{name}: expected a CUDA tensor on {self.device}")
        if t.dtype != dtype:
            raise TypeError(f"{name}: expected dtype {dtype}, got {t.dtype}")
        if tuple(t.shape) != tuple(shape):
            raise ValueError(f"{name}: expected shape {tuple(shape)}, got {tuple(t.shape)}")
        if not t.is_contiguous():
            raise ValueError(f"{name}: must be contiguous")
        return t

    def check(self):
        """Synchronise and raise if a device-side queue overflowed (edge queues, Hough peak list: a frame whose
        ``counts`` exceeds the peak capacity max(4 * max_lines, 65536) is reported here, its lines are unreliable)."""
        self._ok(lib.le_check(self._ctx, self._stream()))

    @property
    def launches(self):
        return lib.le_launch_count(self._ctx)

    def timing(self, on):
        self._ok(lib.le_timing_enable(self._ctx, int(on)))

    def timing_read(self):
        """{kernel_name: (total_ms, launches)} since the last read (synchronises the recorded events)."""
        out = {}
        for k in range(10):
            ms, cnt = C.c_double(), C.c_int()
            self._ok(lib.le_timing_read(self._ctx, k, C.byref(ms), C.byref(cnt)))
            if cnt.value:
                out[lib.le_kernel_name(k).decode()] = (ms.value, cnt.value)
        return out

    @property
    def workspace_bytes(self):
        return lib.le_workspace_bytes(self._ctx)

    # ------------------------------------------------------------------ a1-a4
    def bgr2gray(self, bgr):
        bgr = self._img(bgr, 3)
        n, h, w, _ = bgr.shape
        out = torch.empty((n, h, w), dtype=torch.uint8, device=self.device)
        self._ok(lib.le_bgr2gray_u8(self._ctx, _ptr(bgr), _ptr(out), n, h, w, self._stream()))
        return out

    def gauss5(self, gray):
        gray = self._img(gray, 1)
        n, h, w = gray.shape
        out = torch.empty_like(gray)
        self._ok(lib.le_gauss5_u8(self._ctx, _ptr(gray), _ptr(out), n, h, w, self._stream()))
        return out

    def normalize_minmax(self, gray):
        gray = self._img(gray, 1)
        n, h, w = gray.shape
        out = torch.empty_like(gray)
        self._ok(lib.le_minmax_normalize_u8(self._ctx, _ptr(gray), _ptr(out), n, h, w, self._stream()))
        return out

    def canny(self, gray, lo, hi, out=None):
        return self.front_canny(gray, lo, hi, blur=False, normalize=False, out=out)

    def canny_c3(self, src, lo, hi, out=None):
        """cv.Canny on 3-channel images [N,H,W,3] (per pixel the channel with the largest |dx|+|dy|; a cv2
        convention, not the BGR->gray chain of ``front_canny``)."""
        src = self._img(src, 3)
        n, h, w = src.shape[:3]
        if out is None:
            out = torch.empty((n, h, w), dtype=torch.uint8, device=self.device)
        self._out(out, (n, h, w), torch.uint8, "canny_c3 out")
        self._ok(lib.le_canny_u8c3(self._ctx, _ptr(src), _ptr(out), n, h, w, float(lo), float(hi), self._stream()))
        self._last_edges_shape = (n, h, w)
        return out

    def front_canny(self, src, lo, hi, blur=False, normalize=False, out=None, channels=None):
        """[BGR->gray] -> [blur5] -> [normalize] -> Canny.  ``src`` is [N,H,W,3] BGR or [N,H,W] gray (or one gray
        frame [H,W]).  A 3-D tensor whose last dimension is 3 is ambiguous (one BGR frame, or H gray frames three
        pixels wide) and is refused unless ``channels`` says which."""
        if channels is None:
            if src.dim() == 3 and src.shape[-1] == 3:
                raise ValueError("front_canny: a [.,.,3] tensor is ambiguous -- pass one BGR frame as [1,H,W,3], "
                                 "or say channels=1 / channels=3")
            channels = 3 if src.dim() == 4 else 1
        if channels not in (1, 3):
            raise ValueError("channels must be 1 or 3")
        ch = channels
        src = self._img(src, ch)
        n, h, w = src.shape[:3]
        if out is None:
            out = torch.empty((n, h, w), dtype=torch.uint8, device=self.device)
        self._out(out, (n, h, w), torch.uint8, "front_canny out")
        self._ok(lib.le_front_canny_u8(self._ctx, _ptr(src), ch, int(blur), int(normalize), _ptr(out), n, h, w,
                                       float(lo), float(hi), self._stream()))
        self._last_edges_shape = (n, h, w)
        return out

    # ------------------------------------------------------------------ n4: float32 frame-buffer ingest
    @staticmethod
    def _f32_strided(src):
        """(device tensor holding the frame memory, element offset of pixel [0,0,0,0], element strides n,y,x,c)
        for a float32 [N,H,W,3] CUDA tensor (any strides) or a tuple (base, offset, strides) made by
        ``upload_f32_view`` (numpy views with negative strides, which torch cannot represent)."""
        if isinstance(src, tuple):
            return src
        if not (isinstance(src, torch.Tensor) and src.is_cuda and src.dtype == torch.float32):
            raise TypeError("expected a CUDA float32 tensor")
        if src.dim() == 3:
            src = src[None]
        if src.dim() != 4 or src.shape[-1] != 3:
            raise ValueError("expected [N,H,W,3]")
        return src, 0, tuple(src.stride()), tuple(src.shape[:3])

    def upload_f32_view(self, view):
        """Upload the memory behind a float32 numpy view [H,W,3] or [N,H,W,3] WITHOUT materialising the view
        (``_fboToImage`` returns ``raw.reshape(h,w,3)[::-1,:,::-1]``: negative row and channel strides)."""
        if view.dtype != np.float32:
            raise TypeError("expected float32")
        if view.ndim == 3:
            view = view[None]
        if view.ndim != 4 or view.shape[-1] != 3:
            raise ValueError("expected [N,H,W,3]")
        if any(s % 4 for s in view.strides):
            view = np.ascontiguousarray(view)
        lo = hi = 0  # byte extent of the view relative to its first element
        for n, s in zip(view.shape, view.strides):
            if s < 0:
                lo += (n - 1) * s
            else:
                hi += (n - 1) * s
        raw = (C.c_char * (hi - lo + 4)).from_address(view.__array_interface__["data"][0] + lo)
        base = torch.frombuffer(raw, dtype=torch.float32).to(self.device)  # `view` keeps the memory alive
        return base, -lo // 4, tuple(s // 4 for s in view.strides), tuple(view.shape[:3])

    def f32_bgr2gray(self, src):
        base, off, (sn, sy, sx, sc), (n, h, w) = self._f32_strided(src)
        out = torch.empty((n, h, w), dtype=torch.float32, device=self.device)
        self._ok(lib.le_f32_bgr2gray(self._ctx, C.c_void_p(base.data_ptr() + 4 * off), sn, sy, sx, sc, _ptr(out),
                                     n, h, w, self._stream()))
        return out

    def _f32_plane(self, t):
        if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32):
            raise TypeError("expected a CUDA float32 tensor")
        if t.dim() == 2:
            t = t[None]
        if t.dim() != 3:
            raise ValueError("expected [N,H,W]")
        return t.contiguous()

    def f32_gauss5(self, gray):
        gray = self._f32_plane(gray)
        n, h, w = gray.shape
        out = torch.empty_like(gray)
        self._ok(lib.le_f32_gauss5(self._ctx, _ptr(gray), _ptr(out), n, h, w, self._stream()))
        return out

    def f32_normalize_minmax(self, src, as_u8=False):
        """cv.normalize(f32, None, 0, 255, NORM_MINMAX); ``as_u8`` adds the reference's .astype(np.uint8)."""
        src = self._f32_plane(src)
        n, h, w = src.shape
        out = torch.empty((n, h, w), dtype=torch.uint8 if as_u8 else torch.float32, device=self.device)
        self._ok(lib.le_f32_minmax_normalize(self._ctx, _ptr(src), None if as_u8 else _ptr(out),
                                             _ptr(out) if as_u8 else None, n, h, w, self._stream()))
        return out

    def f32_front_canny(self, src, lo, hi, out=None):
        """float32 BGR -> gray -> blur5 -> normalize -> uint8 -> Canny (doCanny on a frame-buffer image)."""
        base, off, (sn, sy, sx, sc), (n, h, w) = self._f32_strided(src)
        if out is None:
            out = torch.empty((n, h, w), dtype=torch.uint8, device=self.device)
        self._out(out, (n, h, w), torch.uint8, "f32_front_canny out")
        self._ok(lib.le_f32_front_canny(self._ctx, C.c_void_p(base.data_ptr() + 4 * off), sn, sy, sx, sc, _ptr(out),
                                        n, h, w, float(lo), float(hi), self._stream()))
        self._last_edges_shape = (n, h, w)
        return out

    # ------------------------------------------------------------------ a5
    @staticmethod
    def hough_dims(h, w, rho=1.0, theta=math.pi / 180):
        na, nr = C.c_int(), C.c_int()
        if lib.le_hough_dims(h, w, rho, theta, C.byref(na), C.byref(nr)) != 0:
            raise ValueError("bad Hough parameters")
        return na.value, nr.value

    def hough_lines(self, edges, rho, theta, threshold, max_lines=4096, want_accum=False, want_votes=True,
                    shape=None, out=None):
        """``edges`` [N,H,W] u8 (non-zero = edge) or None to vote from the edge lists left by the last
        ``front_canny`` (then ``shape`` defaults to that call's).  Returns (lines [N,max,2] f32,
        votes [N,max] i32 or None, counts [N] i32, accum [N,na+2,nr+2] i32 or None)."""
        if edges is None:
            n, h, w = shape or self._last_edges_shape
        else:
            edges = self._img(edges, 1)
            n, h, w = edges.shape
        na, nr = self.hough_dims(h, w, rho, theta)
        if out is not None:
            lines, votes, counts, accum = out
            if not (isinstance(lines, torch.Tensor) and lines.dim() == 3):
                raise ValueError("hough_lines out: lines must be [N, max_lines, 2]")
            self._out(lines, (n, lines.shape[1], 2), torch.float32, "hough_lines out: lines")
            self._out(votes, (n, lines.shape[1]), torch.int32, "hough_lines out: votes")
            self._out(counts, (n,), torch.int32, "hough_lines out: counts")
            self._out(accum, (n, na + 2, nr + 2), torch.int32, "hough_lines out: accum")
        else:
            lines = torch.empty((n, max_lines, 2), dtype=torch.float32, device=self.device)
            votes = torch.empty((n, max_lines), dtype=torch.int32, device=self.device) if want_votes else None
            counts = torch.empty((n,), dtype=torch.int32, device=self.device)
            accum = torch.empty((n, na + 2, nr + 2), dtype=torch.int32, device=self.device) if want_accum else None
        self._ok(lib.le_hough_lines(self._ctx, _ptr(edges), n, h, w, float(rho), float(theta), int(threshold),
                                    _ptr(accum), _ptr(lines), _ptr(votes), _ptr(counts), lines.shape[1],
                                    self._stream()))
        return lines, votes, counts, accum

    # ------------------------------------------------------------------ a6
    def hough_lines_p(self, edges, rho, theta, threshold, min_line_length=0, max_line_gap=0, max_segs=4096,
                      exact_order=True):
        """-> (segs [N,max_segs,4] i32 in cv2's emission order, counts [N] i32).  ``exact_order=False`` raises
        LE_ERR_UNSUPPORTED: an order-free HoughLinesP cannot meet the 1 px / 99 % criterion (see the header)."""
        edges = self._img(edges, 1)
        n, h, w = edges.shape
        segs = torch.empty((n, max_segs, 4), dtype=torch.int32, device=self.device)
        counts = torch.empty((n,), dtype=torch.int32, device=self.device)
        self._ok(lib.le_hough_lines_p(self._ctx, _ptr(edges), n, h, w, float(rho), float(theta), int(threshold),
                                      int(min_line_length), int(max_line_gap), _ptr(segs), _ptr(counts), max_segs,
                                      int(bool(exact_order)), self._stream()))
        return segs, counts

    # ------------------------------------------------------------------ a7
    def lsd(self, gray, refine=0, scale=0.8, sigma_scale=0.6, quant=2.0, ang_th=22.5, log_eps=0.0, density_th=0.7,
            n_bins=1024, max_segs=4096, want_nfa=False):
        """-> (segs [N,max,4] f32, width [N,max] f64, prec [N,max] f64, counts [N] i32), plus nfa [N,max] f64 as a
        fifth element when ``want_nfa`` (cv2 fills it for refine == 2 only; -1 otherwise)."""
        gray = self._img(gray, 1)
        n, h, w = gray.shape
        segs = torch.empty((n, max_segs, 4), dtype=torch.float32, device=self.device)
        width = torch.empty((n, max_segs), dtype=torch.float64, device=self.device)
        prec = torch.empty((n, max_segs), dtype=torch.float64, device=self.device)
        nfa = torch.full((n, max_segs), -1.0, dtype=torch.float64, device=self.device) if want_nfa else None
        counts = torch.empty((n,), dtype=torch.int32, device=self.device)
        self._ok(lib.le_lsd(self._ctx, _ptr(gray), n, h, w, int(refine), float(scale), float(sigma_scale),
                            float(quant), float(ang_th), float(log_eps), float(density_th), int(n_bins), _ptr(segs),
                            _ptr(width), _ptr(prec), _ptr(nfa), _ptr(counts), max_segs, self._stream()))
        return (segs, width, prec, counts, nfa) if want_nfa else (segs, width, prec, counts)

    # ------------------------------------------------------------------ a9-a12
    def segments_to_polar(self, segs):
        segs = segs.to(self.device, torch.float64).contiguous().view(-1, 4)
        out = torch.empty((segs.shape[0], 2), dtype=torch.float64, device=self.device)
        if segs.shape[0]:
            self._ok(lib.le_segments_to_polar(self._ctx, _ptr(segs), segs.shape[0], _ptr(out), self._stream()))
        return out

    def vp_loss(self, lines, cand, width=600.0, height=400.0):
        lines = lines.to(self.device, torch.float64).contiguous().view(-1, 2)
        cand = cand.to(self.device, torch.float64).contiguous().view(-1, 2)
        out = torch.empty((cand.shape[0],), dtype=torch.float64, device=self.device)
        self._ok(lib.le_vp_loss(self._ctx, _ptr(lines), lines.shape[0], _ptr(cand), cand.shape[0], float(width),
                                float(height), _ptr(out), self._stream()))
        return out

    def vp_sphere_hist(self, segs, width=600.0, height=400.0, n_az=360, n_el=90):
        segs = segs.to(self.device, torch.float64).contiguous().view(-1, 4)
        hist = torch.empty((n_el, n_az), dtype=torch.int32, device=self.device)
        self._ok(lib.le_vp_sphere_hist(self._ctx, _ptr(segs), segs.shape[0], float(width), float(height), _ptr(hist),
                                       n_az, n_el, self._stream()))
        return hist

    def pair_intersections(self, lines, max_out=None):
        lines = lines.to(self.device, torch.float32).contiguous().view(-1, 2)
        m = lines.shape[0]
        max_out = max(1, m * (m - 1) // 2 if max_out is None else max_out)
        out = torch.empty((max_out, 2), dtype=torch.float64, device=self.device)
        count = torch.zeros((1,), dtype=torch.int32, device=self.device)
        self._ok(lib.le_pair_intersections(self._ctx, _ptr(lines), m, _ptr(out), _ptr(count), max_out,
                                           self._stream()))
        return out, count

    # ------------------------------------------------------------------ a14 'lsd' branch, batched (config 5)
    def polar_batch(self, segs, counts, min_len=10.0):
        """LSD output of a batch (segs [N, max, 4] float32, counts [N]) -> (polar [N, max, 2] float64, pcounts [N]):
        per frame the segments longer than ``min_len`` as (rho, phi) lines (opencv_fit_color.py:88)."""
        segs = segs.to(self.device, torch.float32).contiguous()
        counts = counts.to(self.device, torch.int32).contiguous()
        n, max_segs = segs.shape[0], segs.shape[1]
        polar = torch.empty((n, max_segs, 2), dtype=torch.float64, device=self.device)
        pcounts = torch.empty((n,), dtype=torch.int32, device=self.device)
        if n and max_segs:
            self._ok(lib.le_polar_batch(self._ctx, _ptr(segs), _ptr(counts), n, max_segs, float(min_len), _ptr(polar),
                                        _ptr(pcounts), self._stream()))
        else:
            pcounts.zero_()
        return polar, pcounts

    def vp_fit_batch(self, polar, pcounts, width=600.0, height=400.0, n_phi=192, n_th=96, n_ref=32, rounds=4,
                     ref_area=0.3, n_az=360, n_el=90, top=8, want_hist=True):
        """regress_lines for every frame of a batch in one CTA each (opencv_fit_color.py:28-47): polar
        [N, max, 2] float64, pcounts [N] -> (fit [N, 3] float64 = (phi, theta, loss), hist [N, n_el, n_az] int32
        or None)."""
        polar = polar.to(self.device, torch.float64).contiguous()
        pcounts = pcounts.to(self.device, torch.int32).contiguous()
        n, max_lines = polar.shape[0], polar.shape[1]
        fit = torch.empty((n, 3), dtype=torch.float64, device=self.device)
        hist = torch.empty((n, n_el, n_az), dtype=torch.int32, device=self.device) if want_hist else None
        self._ok(lib.le_vp_fit_batch(self._ctx, _ptr(polar), _ptr(pcounts), n, max_lines, float(width), float(height),
                                     int(n_phi), int(n_th), int(n_ref), int(rounds), float(ref_area),
                                     _ptr(hist) if want_hist else None, n_az, n_el, top if want_hist else 0,
                                     _ptr(fit), self._stream()))
        return fit, hist


    def combine_parallel_lines(self, segs, counts=None, max_distance=5, max_angle=3):
        """Batched combineParallelLines (util.py:141-159): segs [N, max_k, 4] float64 (x1,y1,x2,y2), counts [N]
        or None.  Returns (segs [N, max_k, 4], src [N, max_k] int32, counts [N] int32); src holds the original
        index of each surviving slot with bit 30 set when the slot is a merged segment."""
        segs = segs.to(self.device, torch.float64).contiguous()
        n, max_k = segs.shape[0], segs.shape[1]
        if counts is not None:
            counts = counts.to(self.device, torch.int32).contiguous()
        out = torch.empty_like(segs)
        src = torch.empty((n, max_k), dtype=torch.int32, device=self.device)
        out_counts = torch.empty((n,), dtype=torch.int32, device=self.device)
        cosd = lambda a: float(np.cos(np.radians(a)))  # noqa: E731  (the reference's own expression)
        self._ok(lib.le_combine_parallel_lines(self._ctx, _ptr(segs), _ptr(counts) if counts is not None else None, n,
                                               max_k, float(max_distance), cosd(max_angle), 5.0, cosd(3), _ptr(out),
                                               _ptr(src), _ptr(out_counts), self._stream()))
        return out, src, out_counts

    def face_quads(self, segs1, segs2, counts1=None, counts2=None, threshold=15, max_out=4096):
        """Batched quad search of get_faces_from_pairs (opencv_renewed.py:62-112): segs1 [N, max1, 4], segs2
        [N, max2, 4] float64 (x1,y1,x2,y2), counts [N] or None.  Returns (quads [N, max_out, 4] int32 in the
        reference's (i,j,k,l) order, faces [N, max_out, 4, 2] float64 oriented, keep [N, max_out] uint8,
        counts [N] int32 = true number of candidates)."""
        segs1 = segs1.to(self.device, torch.float64).contiguous()
        segs2 = segs2.to(self.device, torch.float64).contiguous()
        n, max1, max2 = segs1.shape[0], segs1.shape[1], segs2.shape[1]
        if segs2.shape[0] != n:
            raise ValueError("segs1 and segs2 must hold the same number of problems")
        c1 = counts1.to(self.device, torch.int32).contiguous() if counts1 is not None else None
        c2 = counts2.to(self.device, torch.int32).contiguous() if counts2 is not None else None
        quads = torch.empty((n, max_out, 4), dtype=torch.int32, device=self.device)
        faces = torch.empty((n, max_out, 4, 2), dtype=torch.float64, device=self.device)
        keep = torch.empty((n, max_out), dtype=torch.uint8, device=self.device)
        counts = torch.empty((n,), dtype=torch.int32, device=self.device)
        self._ok(lib.le_face_quads(self._ctx, _ptr(segs1), _ptr(c1), max1, _ptr(segs2), _ptr(c2), max2, n,
                                   float(threshold), _ptr(quads), _ptr(faces), _ptr(keep), _ptr(counts), max_out,
                                   self._stream()))
        return quads, faces, keep, counts


_default = {}


def default_engine(device=None):
    """Process-wide engine for the single-image numpy API (one per device)."""
    idx = torch.cuda.current_device() if device is None else int(device)
    if idx not in _default:
        _default[idx] = Engine(idx)
    return _default[idx]
