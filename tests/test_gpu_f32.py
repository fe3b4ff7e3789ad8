"""GPU parity of the float32 frame-buffer ingest through the C ABI: gray / blur / normalize
bit-exact against the oracle, doCanny on a frame-buffer view against the reference's own golden edge maps."""
import os

import numpy as np
import pytest
import torch

from oracle import restate as R
from test_f32_oracle import IMAGES, fbo_view

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def golden_f32():
    return dict(np.load(os.path.join(GOLDEN, "golden_f32.npz")))


def golden_edges(golden_f32, name, h, w):
    return np.unpackbits(golden_f32[name + "/edges"])[:h * w].reshape(h, w).astype(np.uint8) * 255


@pytest.mark.parametrize("name", IMAGES)
def test_f32_stages_and_fused_golden(engine, name, golden_inputs, golden_f32):
    view = fbo_view(golden_inputs[name])
    h, w = view.shape[:2]
    up = engine.upload_f32_view(view)  # raw frame-buffer bytes + the view's negative strides
    g = engine.f32_bgr2gray(up)
    og = R.f32_bgr2gray(view)
    assert np.array_equal(g[0].cpu().numpy(), og)
    b = engine.f32_gauss5(g)
    ob = R.f32_gauss5(og)
    assert np.array_equal(b[0].cpu().numpy(), ob)
    assert np.array_equal(engine.f32_normalize_minmax(b)[0].cpu().numpy(), R.f32_normalize_minmax(ob))
    nm8 = engine.f32_normalize_minmax(b, as_u8=True)
    assert np.array_equal(nm8[0].cpu().numpy(), R.f32_normalize_minmax(ob, as_u8=True))
    want = golden_edges(golden_f32, name, h, w)
    assert np.array_equal(engine.canny(nm8, 30, 50)[0].cpu().numpy(), want)
    assert np.array_equal(engine.f32_front_canny(up, 30, 50)[0].cpu().numpy(), want)
    # the fused call leaves the edge lists: HoughLines straight from them (postProcessFbo, opencv.py:118)
    lines, votes, counts, _ = engine.hough_lines(None, 1, np.pi / 180, 70, max_lines=16384)
    ref, rv = R.hough_lines(want, 1, np.pi / 180, 70)
    n = int(counts[0])
    assert n == (0 if ref is None else len(ref))
    if n:
        assert np.array_equal(lines[0, :n].cpu().numpy(), ref[:, 0])
    engine.check()


@pytest.mark.parametrize("seed,n,h,w", [(0, 1, 270, 480), (1, 3, 64, 48), (2, 2, 37, 35), (3, 1, 5, 16), (4, 1, 8, 3),
                                        (5, 2, 129, 200)])
def test_f32_random_batches(engine, seed, n, h, w):
    rng = np.random.default_rng(seed)
    raw = rng.random((n, h, w, 3), dtype=np.float32)
    t = torch.from_numpy(raw).cuda()
    g = engine.f32_bgr2gray(t)
    b = engine.f32_gauss5(g)
    nm = engine.f32_normalize_minmax(b)
    e = engine.f32_front_canny(t, 30, 50)
    e2 = engine.f32_front_canny(t.flip(1), 50, 30)  # a strided torch view is copied by .flip; thresholds swap
    for i in range(n):
        og = R.f32_bgr2gray(raw[i])
        ob = R.f32_gauss5(og)
        assert np.array_equal(g[i].cpu().numpy(), og)
        assert np.array_equal(b[i].cpu().numpy(), ob)
        assert np.array_equal(nm[i].cpu().numpy(), R.f32_normalize_minmax(ob))
        assert np.array_equal(e[i].cpu().numpy(), R.f32_do_canny(raw[i]))
        assert np.array_equal(e2[i].cpu().numpy(), R.f32_do_canny(raw[i, ::-1]))
    # channel-reversed, row-reversed numpy view of the whole batch, consumed in place
    view = raw[:, ::-1, :, ::-1]
    e3 = engine.f32_front_canny(engine.upload_f32_view(view), 30, 50)
    for i in range(n):
        assert np.array_equal(e3[i].cpu().numpy(), R.f32_do_canny(view[i]))
    engine.check()


def test_f32_constant_frame(engine):
    z = torch.full((1, 32, 48, 3), 0.5, dtype=torch.float32, device="cuda")
    nm = engine.f32_normalize_minmax(engine.f32_gauss5(engine.f32_bgr2gray(z)))
    zc = np.full((32, 48, 3), 0.5, np.float32)
    assert np.array_equal(nm[0].cpu().numpy(), R.f32_normalize_minmax(R.f32_gauss5(R.f32_bgr2gray(zc))))
    assert not engine.f32_front_canny(z, 30, 50).any()


def test_reference_level_docanny_on_fbo(engine, golden_inputs, golden_f32):
    """detectors._fboToImage + detectors.doCanny == the reference's postProcessFbo front (opencv.py:101-113),
    and the cv2-signature shim reproduces doCanny's four float32 calls one by one."""
    import graphics_project_b200 as gp
    from graphics_project_b200 import cv_shim as cv

    class Fbo:
        def __init__(self, img):
            self.height, self.width = img.shape[:2]
            self._b = np.ascontiguousarray(img[::-1, :, ::-1].astype(np.float32) / np.float32(255)).tobytes()

        def read(self, components=3, dtype="f4"):
            return self._b

    name = "demo_scarce"
    img = golden_inputs[name]
    h, w = img.shape[:2]
    image = gp.detectors._fboToImage(Fbo(img))
    assert image.dtype == np.float32 and image.strides[0] < 0 and image.strides[2] < 0
    want = golden_edges(golden_f32, name, h, w)
    assert np.array_equal(gp.detectors.doCanny(image), want)
    gray = cv.cvtColor(image, cv.COLOR_BGR2GRAY)
    assert gray.dtype == np.float32
    blurred = cv.GaussianBlur(gray, (5, 5), 0)
    got = cv.Canny(cv.normalize(blurred, None, 0, 255, cv.NORM_MINMAX).astype(np.uint8), 30, 50)
    assert np.array_equal(got, want)
