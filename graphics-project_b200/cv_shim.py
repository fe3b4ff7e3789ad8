"""cv2-signature front end of the engine: the drop-in seam for the reference.

The reference reaches its hot path through the module object bound as ``cv`` (``import cv2 as cv``:
opencv.py:1, opencv_fit_color.py:1, opencv_fit.py:1).  ``install()`` puts this module there
(``sys.modules['cv2']``) so reference code runs unchanged; every attribute that is NOT on the hot
path (imread, imshow, line, solveP3P, Rodrigues, constants, ...) is forwarded to the real cv2.
The hot-path functions below never call cv2: they upload the numpy image, run the CUDA engine
through the C ABI and download cv2-shaped results (dtype, shape and ``None``-when-empty as cv2).
Parameter combinations the reference does not use raise ``cv2.error`` -- there is no CPU fallback.
"""
import math
import sys

import numpy as np
import torch

from .engine import default_engine

try:  # the real cv2 is only used for non-hot-path attributes and for its exception type
    import cv2 as _real_cv2
    if getattr(_real_cv2, "__le_shim__", False):  # already replaced by install()
        _real_cv2 = _real_cv2._real_cv2
except Exception:  # pragma: no cover
    _real_cv2 = None

__le_shim__ = True

if _real_cv2 is not None:
    error = _real_cv2.error
else:  # pragma: no cover
    class error(Exception):
        pass

COLOR_BGR2GRAY = 6
NORM_MINMAX = 32
LSD_REFINE_NONE, LSD_REFINE_STD, LSD_REFINE_ADV = 0, 1, 2


def __getattr__(name):  # PEP 562: everything else comes from the real cv2
    if _real_cv2 is None:
        raise AttributeError(name)
    return getattr(_real_cv2, name)


def _err(msg):
    return error(f"graphics_project_b200.cv_shim: {msg}")


def _dev(a):
    eng = default_engine()
    return eng, torch.from_numpy(np.ascontiguousarray(a)).to(eng.device, non_blocking=False)


def _is_f32(img, channels):
    return (isinstance(img, np.ndarray) and img.dtype == np.float32 and
            ((channels == 1 and img.ndim == 2) or (channels == 3 and img.ndim == 3 and img.shape[2] == 3)))


def _check_u8(img, channels, what):
    if not isinstance(img, np.ndarray):
        raise _err(f"{what}: expected a numpy array")
    if img.dtype != np.uint8:
        raise _err(f"{what}: (-215:Assertion failed) depth == CV_8U; got {img.dtype}")
    if channels == 1 and img.ndim != 2:
        raise _err(f"{what}: (-215:Assertion failed) type == CV_8UC1; got shape {img.shape}")
    if channels == 3 and (img.ndim != 3 or img.shape[2] != 3):
        raise _err(f"{what}: expected an HxWx3 image; got shape {img.shape}")


# ----------------------------------------------------------------------------- a1-a4
def cvtColor(src, code, dst=None, dstCn=0):
    """cv.cvtColor(image, cv.COLOR_BGR2GRAY) -- opencv.py:22,205,213."""
    if code != COLOR_BGR2GRAY:
        raise _err(f"cvtColor: only COLOR_BGR2GRAY is on the hot path (code {code})")
    if _is_f32(src, 3):  # the simulator's frame-buffer image (opencv.py:101-113), any strides
        eng = default_engine()
        return eng.f32_bgr2gray(eng.upload_f32_view(src))[0].cpu().numpy()
    _check_u8(src, 3, "cvtColor")
    eng, t = _dev(src)
    return eng.bgr2gray(t[None])[0].cpu().numpy()


def GaussianBlur(src, ksize, sigmaX, dst=None, sigmaY=0, borderType=4):
    """cv.GaussianBlur(gray, (5,5), 0) -- opencv.py:25; opencv_fit_color.py:77."""
    if tuple(ksize) != (5, 5) or sigmaX != 0 or sigmaY != 0 or borderType != 4:
        raise _err("GaussianBlur: only ksize (5,5), sigma 0, BORDER_REFLECT_101 is on the hot path")
    if _is_f32(src, 1):
        eng, t = _dev(src)
        return eng.f32_gauss5(t[None])[0].cpu().numpy()
    _check_u8(src, 1, "GaussianBlur")
    eng, t = _dev(src)
    return eng.gauss5(t[None])[0].cpu().numpy()


def normalize(src, dst, alpha=1.0, beta=0.0, norm_type=4, dtype=-1, mask=None):
    """cv.normalize(blurred, None, 0, 255, cv.NORM_MINMAX) -- opencv.py:26."""
    if (alpha, beta, norm_type) != (0, 255, NORM_MINMAX) or mask is not None or dtype not in (-1, 0):
        raise _err("normalize: only (alpha=0, beta=255, NORM_MINMAX) on uint8 / float32 is on the hot path")
    if _is_f32(src, 1) and dtype == -1:  # float32 in, float32 out; doCanny then truncates with astype(np.uint8)
        eng, t = _dev(src)
        return eng.f32_normalize_minmax(t[None])[0].cpu().numpy()
    _check_u8(src, 1, "normalize")
    eng, t = _dev(src)
    return eng.normalize_minmax(t[None])[0].cpu().numpy()


def Canny(image, threshold1, threshold2, edges=None, apertureSize=3, L2gradient=False):
    """cv.Canny(img, lo, hi[, apertureSize=3]) -- opencv.py:26,214; opencv_fit_color.py:78."""
    if apertureSize != 3 or L2gradient:
        raise _err("Canny: only apertureSize=3, L2gradient=False is on the hot path")
    if isinstance(image, np.ndarray) and image.dtype == np.uint8 and image.ndim == 3 and image.shape[2] == 3:
        eng, t = _dev(image)  # cv2 accepts a 3-channel image: per pixel the channel with the largest gradient
        return eng.canny_c3(t[None], threshold1, threshold2)[0].cpu().numpy()
    _check_u8(image, 1, "Canny")
    eng, t = _dev(image)
    return eng.canny(t[None], threshold1, threshold2)[0].cpu().numpy()


# ----------------------------------------------------------------------------- a5
def _hough_std(image, rho, theta, threshold, srn, stn, min_theta, max_theta):
    if srn != 0 or stn != 0 or min_theta != 0 or abs(max_theta - math.pi) > 1e-12:
        raise _err("HoughLines: multi-scale (srn/stn) and theta ranges are not on the hot path")
    _check_u8(image, 1, "HoughLines")
    eng, t = _dev(image)
    cap = 4096
    while True:
        lines, votes, counts, _ = eng.hough_lines(t[None], rho, theta, threshold, max_lines=cap)
        n = int(counts[0].item())
        if n <= cap:
            break
        cap = n
    return lines[0, :n].cpu().numpy(), votes[0, :n].cpu().numpy()


def HoughLines(image, rho, theta, threshold, lines=None, srn=0, stn=0, min_theta=0, max_theta=math.pi):
    """cv.HoughLines(canny, 1, np.pi/180, thr, None, 0, 0) -> float32 (N,1,2) or None."""
    ln, _ = _hough_std(image, rho, theta, threshold, srn, stn, min_theta, max_theta)
    return ln.reshape(-1, 1, 2) if len(ln) else None


def HoughLinesWithAccumulator(image, rho, theta, threshold, lines=None, srn=0, stn=0, min_theta=0,
                              max_theta=math.pi):
    """-> float32 (N,1,3): rho, theta, votes."""
    ln, votes = _hough_std(image, rho, theta, threshold, srn, stn, min_theta, max_theta)
    if not len(ln):
        return None
    return np.concatenate([ln, votes.astype(np.float32)[:, None]], 1).reshape(-1, 1, 3)


# ----------------------------------------------------------------------------- a6
def _cv_round(v):
    return int(np.rint(v))


def HoughLinesP(image, rho, theta, threshold, lines=None, minLineLength=0, maxLineGap=0):
    """cv.HoughLinesP(edges, 1, np.pi/180, threshold, minLineLength, maxLineGap) -> int32 (N,1,4) or None."""
    _check_u8(image, 1, "HoughLinesP")
    eng, t = _dev(image)
    cap = 4096
    while True:
        segs, counts = eng.hough_lines_p(t[None], rho, theta, threshold, _cv_round(minLineLength),
                                         _cv_round(maxLineGap), max_segs=cap)
        n = int(counts[0].item())
        if n <= cap:
            break
        cap = n
    return segs[0, :n].cpu().numpy().reshape(-1, 1, 4) if n else None


# ----------------------------------------------------------------------------- a7
class _LineSegmentDetector:
    def __init__(self, refine, scale, sigma_scale, quant, ang_th, log_eps, density_th, n_bins):
        self.params = dict(refine=int(refine), scale=float(scale), sigma_scale=float(sigma_scale),
                           quant=float(quant), ang_th=float(ang_th), log_eps=float(log_eps),
                           density_th=float(density_th), n_bins=int(n_bins))
        if self.params["refine"] not in (LSD_REFINE_NONE, LSD_REFINE_STD, LSD_REFINE_ADV):
            raise _err("createLineSegmentDetector: refine must be LSD_REFINE_NONE, _STD or _ADV")

    def detect(self, image):
        """-> (lines float32 (N,1,4), width f64 (N,1), prec f64 (N,1), nfa f64 (N,1) for LSD_REFINE_ADV else None);
        all None if N == 0."""
        _check_u8(image, 1, "LineSegmentDetector.detect")
        eng, t = _dev(image)
        adv = self.params["refine"] == LSD_REFINE_ADV
        cap = 4096
        while True:
            segs, width, prec, counts, nfa = eng.lsd(t[None], max_segs=cap, want_nfa=True, **self.params)
            n = int(counts[0].item())
            if n <= cap:
                break
            cap = n
        if n == 0:
            return None, None, None, None
        return (segs[0, :n].cpu().numpy().reshape(-1, 1, 4), width[0, :n].cpu().numpy().reshape(-1, 1),
                prec[0, :n].cpu().numpy().reshape(-1, 1), nfa[0, :n].cpu().numpy().reshape(-1, 1) if adv else None)


def createLineSegmentDetector(refine=LSD_REFINE_STD, scale=0.8, sigma_scale=0.6, quant=2.0, ang_th=22.5,
                              log_eps=0, density_th=0.7, n_bins=1024):
    """cv.createLineSegmentDetector(...) -- opencv.py:206."""
    return _LineSegmentDetector(refine, scale, sigma_scale, quant, ang_th, log_eps, density_th, n_bins)


# ----------------------------------------------------------------------------- installation
def install():
    """Make ``import cv2`` resolve to this shim (call BEFORE importing the reference modules)."""
    mod = sys.modules[__name__]
    sys.modules["cv2"] = mod
    return mod


def uninstall():
    if _real_cv2 is not None:
        sys.modules["cv2"] = _real_cv2
