"""CPU: the float32 frame-buffer path of the oracle against the golden vectors made by the
reference's own _fboToImage + doCanny (tests/golden/make_golden_f32.py) and against the live cv2 wheel."""
import hashlib
import os

import numpy as np
import pytest

from oracle import restate as R

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
IMAGES = ["demo_scarce", "sc_white_2", "sc_minecraft_cave", "old_sc_inf", "sc_cube", "sc_scarce_3_flat"]


def sha(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()[:10]


def fbo_view(bgr_u8):
    """The view _fboToImage (opencv.py:101-108) makes of a frame buffer showing this image."""
    raw = np.ascontiguousarray(bgr_u8[::-1, :, ::-1].astype(np.float32) / np.float32(255))
    return raw.reshape(bgr_u8.shape)[::-1, :, ::-1]


@pytest.fixture(scope="module")
def golden_f32():
    return dict(np.load(os.path.join(GOLDEN, "golden_f32.npz")))


@pytest.mark.parametrize("name", IMAGES)
def test_f32_chain_golden(name, golden_inputs, golden_f32):
    view = fbo_view(golden_inputs[name])
    h, w = view.shape[:2]
    g = R.f32_bgr2gray(view)
    b = R.f32_gauss5(g)
    nm = R.f32_normalize_minmax(b)
    want = golden_f32[name + "/sha"]
    assert sha(g) == want[0]
    if w % 16 == 0:  # cv2's scalar tail (last w % 16 columns of GaussianBlur) sums in another order
        assert sha(b) == want[1]
        assert sha(nm) == want[2]
        assert sha(R.f32_normalize_minmax(b, as_u8=True)) == want[3]
    edges = np.unpackbits(golden_f32[name + "/edges"])[:h * w].reshape(h, w).astype(np.uint8) * 255
    assert np.array_equal(R.f32_do_canny(view), edges)


@pytest.mark.parametrize("seed,shape", [(0, (400, 600)), (1, (270, 480)), (2, (64, 48)), (3, (37, 32)), (4, (5, 16)),
                                        (5, (8, 32)), (6, (1080, 1920))])
def test_f32_chain_vs_cv2(seed, shape):
    cv = pytest.importorskip("cv2")
    rng = np.random.default_rng(seed)
    raw = (rng.random(shape + (3,), dtype=np.float32) if seed % 2 == 0 else
           rng.integers(0, 256, shape + (3,)).astype(np.float32) / np.float32(255))
    view = raw[::-1, :, ::-1]
    g = cv.cvtColor(view, cv.COLOR_BGR2GRAY)
    assert np.array_equal(R.f32_bgr2gray(view), g)
    if shape[1] % 16 == 0:
        b = cv.GaussianBlur(g, (5, 5), 0)
        assert np.array_equal(R.f32_gauss5(g), b)
        nm = cv.normalize(b, None, 0, 255, cv.NORM_MINMAX)
        assert np.array_equal(R.f32_normalize_minmax(b), nm)
        assert np.array_equal(R.f32_normalize_minmax(b, as_u8=True), nm.astype(np.uint8))
        assert np.array_equal(R.f32_do_canny(view), cv.Canny(nm.astype(np.uint8), 30, 50))


def test_f32_constant_frame():
    cv = pytest.importorskip("cv2")
    z = np.full((32, 32, 3), 0.5, np.float32)
    nm = cv.normalize(cv.GaussianBlur(cv.cvtColor(z, cv.COLOR_BGR2GRAY), (5, 5), 0), None, 0, 255, cv.NORM_MINMAX)
    assert np.array_equal(R.f32_normalize_minmax(R.f32_gauss5(R.f32_bgr2gray(z))), nm)  # max == min: scale 0
    assert not R.f32_do_canny(z).any()
