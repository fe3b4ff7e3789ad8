"""Reference-level detector wrappers with the reference's own signatures.

``doCanny`` (opencv.py:21-28; uint8 images and the simulator's float32 frame-buffer view,
``_fboToImage`` opencv.py:101-108), ``lsd`` (opencv.py:204-209), ``prob`` (opencv.py:211-219),
``lineMatrixToPairs`` (util.py:110-113), ``_getIntersections`` (opencv.py:171-189) and
``combineParallelLines`` (util.py:141-159).  Each wrapper
issues the fused GPU chain instead of one call per cv2 stage, so a BGR frame is read once.
"""
import numpy as np
import torch

from . import cv_shim as cv
from .engine import default_engine

# constants.json:5-6 (opencv.CANNY_THRESH_SOFT / CANNY_THRESH_HARD)
CANNY_THRESH_SOFT = 30
CANNY_THRESH_HARD = 50


def lineMatrixToPairs(lines):
    """util.py:110-113 -- indexes lines[0], so ``None`` (no detections) raises like the reference."""
    if np.array(lines[0]).shape != (1, 4):
        return lines
    return [(np.array(line[0][0:2]), np.array(line[0][2:4])) for line in lines]


def _fboToImage(fbo):
    """opencv.py:101-108 -- the frame buffer's RGB float32 bytes, bottom-up, as a BGR top-down VIEW.  No pixel
    is touched here: ``doCanny`` uploads the raw bytes once and the kernels address them through the view's
    (negative) strides."""
    raw = np.frombuffer(fbo.read(components=3, dtype="f4"), dtype="f4")
    return raw.reshape((fbo.height, fbo.width, 3))[::-1, :, ::-1]


def _bgr_dev(image):
    cv._check_u8(image, 3, "detector input")
    eng = default_engine()
    return eng, torch.from_numpy(np.ascontiguousarray(image)).to(eng.device)[None]


def doCanny(image, thresh_soft=None, thresh_hard=None):
    """gray -> GaussianBlur(5,5) -> normalize(0,255,MINMAX) -> Canny(soft, hard); u8 edge map."""
    lo = CANNY_THRESH_SOFT if thresh_soft is None else thresh_soft
    hi = CANNY_THRESH_HARD if thresh_hard is None else thresh_hard
    if cv._is_f32(image, 3):
        # the simulator path (postProcessFbo, opencv.py:110-113): float32 frame-buffer view, consumed in place
        eng = default_engine()
        return eng.f32_front_canny(eng.upload_f32_view(image), lo, hi)[0].cpu().numpy()
    eng, t = _bgr_dev(image)
    return eng.front_canny(t, lo, hi, blur=True, normalize=True)[0].cpu().numpy()


def lsd(image, detector=0, scale=0.8, sigma_scale=0.6, quant=2.0, ang_th=22.5, log_eps=0.0, density_th=0.7,
        n_bins=1024):
    """BGR -> gray -> LSD; list of (xy float32, xy float32) pairs."""
    eng, t = _bgr_dev(image)
    gray = eng.bgr2gray(t)
    cap = 4096
    while True:
        segs, _, _, counts = eng.lsd(gray, refine=detector, scale=scale, sigma_scale=sigma_scale, quant=quant,
                                     ang_th=ang_th, log_eps=log_eps, density_th=density_th, n_bins=n_bins,
                                     max_segs=cap)
        n = int(counts[0].item())
        if n <= cap:
            break
        cap = n
    lines = segs[0, :n].cpu().numpy().reshape(-1, 1, 4) if n else None
    return lineMatrixToPairs(lines)


def prob(image, display=False):
    """BGR -> gray -> Canny(5,150) -> HoughLinesP(1, pi/180, 30, 50, 10); list of int32 pairs."""
    eng, t = _bgr_dev(image)
    edges = eng.front_canny(t, 5, 150, blur=False, normalize=False)
    if display:
        cv.imshow("Canny filter for probabilistic Hough", edges[0].cpu().numpy())
    cap = 4096
    while True:
        segs, counts = eng.hough_lines_p(edges, 1, np.pi / 180, 30, 50, 10, max_segs=cap)
        n = int(counts[0].item())
        if n <= cap:
            break
        cap = n
    lines = segs[0, :n].cpu().numpy().reshape(-1, 1, 4) if n else None
    return lineMatrixToPairs(lines)


def _getIntersections(lines):
    """All i<j intersections of (rho, theta) lines with the reference's skip rules; list of (x, y)."""
    arr = np.asarray(lines, dtype=np.float32).reshape(-1, 2)
    if len(arr) < 2:
        return []
    eng = default_engine()
    out, count = eng.pair_intersections(torch.from_numpy(arr))
    n = int(count.item())
    pts = out[:n].cpu().numpy()
    return [(float(x), float(y)) for x, y in pts]


def combineParallelLines(lines, max_distance=5, max_angle=3):
    """util.py:141-159 -- greedy merge of nearly parallel, nearby segments (first mergeable pair, restart).

    ``lines``: list of (p1, p2) point pairs (or (2, 2) arrays) as ``lsd`` / ``prob`` return them.  Like the
    reference, untouched segments come back as the caller's own objects and merged ones as new float64
    (2, 2) arrays; arithmetic is float64 throughout (the reference computes float32 inputs in float32)."""
    if len(lines) < 2:
        return list(lines)
    arr = np.array([[l[0][0], l[0][1], l[1][0], l[1][1]] for l in lines], dtype=np.float64)
    eng = default_engine()
    out, src, counts = eng.combine_parallel_lines(torch.from_numpy(arr)[None], None, max_distance, max_angle)
    m = int(counts[0].item())
    out, src = out[0, :m].cpu().numpy(), src[0, :m].cpu().numpy()
    res = []
    for k in range(m):
        if src[k] & 0x40000000:
            res.append(np.array([[out[k, 0], out[k, 1]], [out[k, 2], out[k, 3]]]))
        else:
            res.append(lines[int(src[k])])
    return res
