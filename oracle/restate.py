"""ctypes binding of oracle/cv_restate.c (TEST INFRASTRUCTURE ONLY)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")


def build(force=False):
    src = os.path.join(_HERE, "cv_restate.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "liboracle.so"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.or_fast_atan2.restype = C.c_float
        _lib.or_fast_atan2.argtypes = [C.c_float, C.c_float]
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _u8(a):
    a = np.ascontiguousarray(a)
    assert a.dtype == np.uint8
    return a


def bgr2gray(bgr):
    bgr = _u8(bgr)
    h, w, _ = bgr.shape
    out = np.empty((h, w), np.uint8)
    lib().or_bgr2gray(_p(bgr), _p(out), C.c_int(h), C.c_int(w))
    return out


def gauss5(gray):
    gray = _u8(gray)
    h, w = gray.shape
    out = np.empty_like(gray)
    lib().or_gauss5(_p(gray), _p(out), C.c_int(h), C.c_int(w))
    return out


def normalize_minmax(src):
    src = _u8(src)
    out = np.empty_like(src)
    lib().or_normalize_minmax(_p(src), _p(out), C.c_long(src.size))
    return out


def _f32(a):
    a = np.ascontiguousarray(a)
    assert a.dtype == np.float32
    return a


def f32_bgr2gray(view):
    """float32 BGR (H,W,3), any strides (the reference's _fboToImage view has negative ones) -> f32 gray."""
    assert view.dtype == np.float32 and view.ndim == 3 and view.shape[2] == 3
    h, w, _ = view.shape
    out = np.empty((h, w), np.float32)
    sy, sx, sc = (s // 4 for s in view.strides)
    lib().or_f32_bgr2gray(C.c_void_p(view.__array_interface__["data"][0]), C.c_long(sy), C.c_long(sx),
                          C.c_long(sc), _p(out), C.c_int(h), C.c_int(w))
    return out


def f32_gauss5(gray):
    gray = _f32(gray)
    h, w = gray.shape
    out = np.empty_like(gray)
    lib().or_f32_gauss5(_p(gray), _p(out), C.c_int(h), C.c_int(w))
    return out


def f32_normalize_minmax(src, as_u8=False):
    src = _f32(src)
    out = np.empty(src.shape, np.uint8 if as_u8 else np.float32)
    lib().or_f32_normalize_minmax(_p(src), None if as_u8 else _p(out), _p(out) if as_u8 else None,
                                  C.c_long(src.size))
    return out


def f32_do_canny(view, lo=30, hi=50):
    """doCanny (opencv.py:21-28) on a float32 BGR frame."""
    return canny(f32_normalize_minmax(f32_gauss5(f32_bgr2gray(view)), as_u8=True), lo, hi)


def canny(src, lo, hi, return_nms=False):
    """Gray [H,W], or a 3-channel image [H,W,3] (cv2 keeps, per pixel, the channel with the largest |dx|+|dy|)."""
    src = _u8(src)
    h, w = src.shape[:2]
    out = np.empty((h, w), np.uint8)
    nms = np.empty((h, w), np.uint8) if return_nms else None
    fn = lib().or_canny_c3 if src.ndim == 3 else lib().or_canny
    assert src.ndim == 2 or src.shape[2] == 3
    fn(_p(src), _p(out), C.c_int(h), C.c_int(w), C.c_double(lo), C.c_double(hi), _p(nms) if return_nms else None)
    return (out, nms) if return_nms else out


def hough_dims(h, w, rho=1.0, theta=np.pi / 180):
    return lib().or_hough_numangle(C.c_double(theta)), lib().or_hough_numrho(C.c_int(h), C.c_int(w), C.c_double(rho))


def hough_trig_std(rho, theta, numangle):
    ts = np.empty(numangle, np.float32)
    tc = np.empty(numangle, np.float32)
    lib().or_hough_trig_std(C.c_float(rho), C.c_float(theta), C.c_int(numangle), _p(ts), _p(tc))
    return ts, tc


def hough_lines(edges, rho, theta, threshold, max_lines=1 << 20, return_accum=False):
    """-> (lines (N,1,2) f32 or None, votes (N,) i32[, accum (numangle+2, numrho+2) i32])"""
    edges = _u8(edges)
    h, w = edges.shape
    na, nr = hough_dims(h, w, rho, theta)
    lines = np.empty((max_lines, 2), np.float32)
    votes = np.empty(max_lines, np.int32)
    accum = np.zeros((na + 2, nr + 2), np.int32)
    n = lib().or_hough_lines(_p(edges), C.c_int(h), C.c_int(w), C.c_float(rho), C.c_float(theta),
                             C.c_int(int(threshold)), _p(lines), _p(votes), C.c_int(max_lines), _p(accum))
    assert n <= max_lines
    ln = lines[:n].reshape(n, 1, 2).copy() if n else None
    res = (ln, votes[:n].copy())
    return res + (accum,) if return_accum else res


def hough_lines_p(edges, rho, theta, threshold, min_len, max_gap, max_segs=1 << 18, seed=None):
    """``seed`` replaces cv2's fixed cv::RNG(-1) state for ONE call (order-sensitivity measurements only)."""
    edges = _u8(edges)
    if seed is not None:
        lib().or_ppht_seed(C.c_uint64(seed))
        try:
            return hough_lines_p(edges, rho, theta, threshold, min_len, max_gap, max_segs)
        finally:
            lib().or_ppht_seed(C.c_uint64((1 << 64) - 1))
    h, w = edges.shape
    segs = np.empty((max_segs, 4), np.int32)
    n = lib().or_hough_lines_p(_p(edges), C.c_int(h), C.c_int(w), C.c_float(rho), C.c_float(theta),
                               C.c_int(int(threshold)), C.c_int(int(min_len)), C.c_int(int(max_gap)),
                               _p(segs), C.c_int(max_segs))
    assert n <= max_segs
    return segs[:n].reshape(n, 1, 4).copy() if n else None


def fast_atan2(y, x):
    return lib().or_fast_atan2(C.c_float(y), C.c_float(x))


def gauss_taps_q8(sigma, k):
    taps = np.empty(k, np.int32)
    lib().or_gauss_taps_q8(C.c_double(sigma), C.c_int(k), _p(taps))
    return taps


def lsd_scale_image(gray, scale=0.8, sigma_scale=0.6):
    gray = _u8(gray)
    h, w = gray.shape
    dh, dw = C.c_int(), C.c_int()
    lib().or_lsd_scaled_dims(C.c_int(h), C.c_int(w), C.c_double(scale), C.byref(dh), C.byref(dw))
    out = np.empty((dh.value, dw.value), np.uint8)
    lib().or_lsd_scale_image(_p(gray), C.c_int(h), C.c_int(w), C.c_double(scale), C.c_double(sigma_scale), _p(out))
    return out


def lsd(gray, refine=0, scale=0.8, sigma_scale=0.6, quant=2.0, ang_th=22.5, log_eps=0.0, density_th=0.7,
        n_bins=1024, max_segs=1 << 16, return_debug=False, return_nfa=False):
    """-> (lines (N,1,4) f32 or None, width (N,1) f64, prec (N,1) f64[, nfa (N,1) f64 with return_nfa]).
    All three refine modes are pinned bit for bit against cv2 4.13 (tests/test_oracle.py, tests/test_lsd2_oracle.py)."""
    assert refine in (0, 1, 2)
    gray = _u8(gray)
    h, w = gray.shape
    lines = np.empty((max_segs, 4), np.float32)
    wd = np.empty(max_segs, np.float64)
    pr = np.empty(max_segs, np.float64)
    nf = np.full(max_segs, -1.0, np.float64)
    dbg_s = dbg_a = None
    if return_debug:
        dh, dw = C.c_int(h), C.c_int(w)
        if scale != 1:
            lib().or_lsd_scaled_dims(C.c_int(h), C.c_int(w), C.c_double(scale), C.byref(dh), C.byref(dw))
        dbg_s = np.empty((dh.value, dw.value), np.uint8)
        dbg_a = np.empty((dh.value, dw.value), np.float64)
    lib().or_lsd_set_nfa_out(_p(nf))
    try:
        n = lib().or_lsd(_p(gray), C.c_int(h), C.c_int(w), C.c_int(refine), C.c_double(scale), C.c_double(sigma_scale),
                         C.c_double(quant), C.c_double(ang_th), C.c_double(log_eps), C.c_double(density_th),
                         C.c_int(n_bins), _p(lines), _p(wd), _p(pr), C.c_int(max_segs),
                         _p(dbg_s) if return_debug else None, _p(dbg_a) if return_debug else None)
    finally:
        lib().or_lsd_set_nfa_out(None)
    assert n <= max_segs
    if n == 0:
        res = (None, None, None) + ((None,) if return_nfa else ())
    else:
        res = (lines[:n].reshape(n, 1, 4).copy(), wd[:n].reshape(n, 1).copy(), pr[:n].reshape(n, 1).copy())
        if return_nfa:
            res = res + (nf[:n].reshape(n, 1).copy(),)
    return res + (dbg_s, dbg_a) if return_debug else res
