"""ctypes binding of liblineengine.so (the C ABI declared in include/line_engine.h).

There is NO fallback: if the shared library is missing or does not export a symbol, importing
this module raises.  Build it with ``python -c "import __graft_entry__ as g; g.build()"`` or
``make -C graphics-project_b200/csrc``.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liblineengine.so")

LE_OK = 0
ERRORS = {-1: "LE_ERR_INVALID", -2: "LE_ERR_CUDA", -3: "LE_ERR_NOMEM", -4: "LE_ERR_OVERFLOW", -5: "LE_ERR_UNSUPPORTED"}

_vp, _i, _d, _sz, _ll = C.c_void_p, C.c_int, C.c_double, C.c_size_t, C.c_longlong

# name -> (restype, argtypes); mirrors include/line_engine.h one to one
PROTOTYPES = {
    "le_version": (C.c_char_p, []),
    "le_create": (_i, [C.POINTER(_vp), _i, _i, _i, _i]),
    "le_destroy": (_i, [_vp]),
    "le_last_error": (C.c_char_p, [_vp]),
    "le_workspace_bytes": (_sz, [_vp]),
    "le_check": (_i, [_vp, _vp]),
    "le_launch_count": (C.c_longlong, [_vp]),
    "le_timing_enable": (_i, [_vp, _i]),
    "le_timing_read": (_i, [_vp, _i, C.POINTER(_d), C.POINTER(_i)]),
    "le_kernel_name": (C.c_char_p, [_i]),
    "le_bgr2gray_u8": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp]),
    "le_gauss5_u8": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp]),
    "le_minmax_normalize_u8": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp]),
    "le_canny_u8": (_i, [_vp, _vp, _vp, _i, _i, _i, _d, _d, _vp]),
    "le_canny_u8c3": (_i, [_vp, _vp, _vp, _i, _i, _i, _d, _d, _vp]),
    "le_front_canny_u8": (_i, [_vp, _vp, _i, _i, _i, _vp, _i, _i, _i, _d, _d, _vp]),
    "le_hough_dims": (_i, [_i, _i, _d, _d, C.POINTER(_i), C.POINTER(_i)]),
    "le_hough_lines": (_i, [_vp, _vp, _i, _i, _i, _d, _d, _i, _vp, _vp, _vp, _vp, _i, _vp]),
    "le_hough_lines_p": (_i, [_vp, _vp, _i, _i, _i, _d, _d, _i, _i, _i, _vp, _vp, _i, _i, _vp]),
    "le_lsd_scaled_dims": (_i, [_i, _i, _d, C.POINTER(_i), C.POINTER(_i)]),
    "le_lsd": (_i, [_vp, _vp, _i, _i, _i, _i, _d, _d, _d, _d, _d, _d, _i, _vp, _vp, _vp, _vp, _vp, _i, _vp]),
    "le_segments_to_polar": (_i, [_vp, _vp, _i, _vp, _vp]),
    "le_vp_loss": (_i, [_vp, _vp, _i, _vp, _i, _d, _d, _vp, _vp]),
    "le_vp_sphere_hist": (_i, [_vp, _vp, _i, _d, _d, _vp, _i, _i, _vp]),
    "le_pair_intersections": (_i, [_vp, _vp, _i, _vp, _vp, _i, _vp]),
    "le_polar_batch": (_i, [_vp, _vp, _vp, _i, _i, _d, _vp, _vp, _vp]),
    "le_vp_fit_batch": (_i, [_vp, _vp, _vp, _i, _i, _d, _d, _i, _i, _i, _i, _d, _vp, _i, _i, _i, _vp, _vp]),
    "le_combine_parallel_lines": (_i, [_vp, _vp, _vp, _i, _i, _d, _d, _d, _d, _vp, _vp, _vp, _vp]),
    "le_face_quads": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _i, _i, _d, _vp, _vp, _vp, _vp, _i, _vp]),
    "le_f32_bgr2gray": (_i, [_vp, _vp, _ll, _ll, _ll, _ll, _vp, _i, _i, _i, _vp]),
    "le_f32_gauss5": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp]),
    "le_f32_minmax_normalize": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "le_f32_front_canny": (_i, [_vp, _vp, _ll, _ll, _ll, _ll, _vp, _i, _i, _i, _d, _d, _vp]),
}


def load(path=LIB_PATH):
    if not os.path.exists(path):
        raise ImportError(
            f"{path} not found: the CUDA extension is not built (run __graft_entry__.build()). "
            "This engine has no CPU fallback.")
    lib = C.CDLL(path)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    return lib


lib = load()
