"""CPU: the C-ABI library loads and exports every symbol include/line_engine.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "line_engine.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(le_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_expected_entry_points():
    syms = header_symbols()
    for must in ("le_create", "le_destroy", "le_last_error", "le_front_canny_u8", "le_canny_u8", "le_hough_lines",
                 "le_hough_lines_p", "le_lsd", "le_vp_loss", "le_vp_sphere_hist", "le_pair_intersections"):
        assert must in syms


def test_library_exports_every_header_symbol():
    lib = ctypes.CDLL(os.path.join(ROOT, "graphics-project_b200", "liblineengine.so"))
    for s in header_symbols():
        assert hasattr(lib, s), f"{s} declared in include/line_engine.h but not exported"


def test_python_binding_matches_header():
    from graphics_project_b200 import _native
    assert sorted(_native.PROTOTYPES) == header_symbols()
    assert _native.lib.le_version().decode().endswith("sm_100a")


def test_pure_host_entry_points():
    """le_hough_dims / le_lsd_scaled_dims do no device work: callable without a GPU."""
    from graphics_project_b200 import _native
    na, nr = ctypes.c_int(), ctypes.c_int()
    import math
    assert _native.lib.le_hough_dims(1080, 1920, 1.0, math.pi / 180, ctypes.byref(na), ctypes.byref(nr)) == 0
    assert (na.value, nr.value) == (180, 6001)
    assert _native.lib.le_hough_dims(400, 600, 1.0, math.pi / 180, ctypes.byref(na), ctypes.byref(nr)) == 0
    assert (na.value, nr.value) == (180, 2001)
    sh, sw = ctypes.c_int(), ctypes.c_int()
    assert _native.lib.le_lsd_scaled_dims(1080, 1920, 0.8, ctypes.byref(sh), ctypes.byref(sw)) == 0
    assert (sh.value, sw.value) == (864, 1536)
    assert _native.lib.le_lsd_scaled_dims(397, 595, 0.8, ctypes.byref(sh), ctypes.byref(sw)) == 0
    assert (sh.value, sw.value) == (318, 476)
    assert _native.lib.le_hough_dims(0, 10, 1.0, 0.1, ctypes.byref(na), ctypes.byref(nr)) != 0


def test_no_gpu_fails_loudly():
    """On a box without a GPU the engine refuses to construct (no CPU fallback)."""
    import torch
    import pytest
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import graphics_project_b200 as gp
    with pytest.raises(RuntimeError):
        gp.Engine()
    ctx = ctypes.c_void_p()
    assert gp._native.lib.le_create(ctypes.byref(ctx), 0, 0, 0, 0) != 0
    assert b"CUDA" in gp._native.lib.le_last_error(None)


def test_product_never_imports_oracle():
    """The product package must not import, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "graphics-project_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", "Makefile")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("no CPU oracle", ""), f"{f} mentions oracle"
