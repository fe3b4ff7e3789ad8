"""CPU differential fuzz: the C oracle against the live cv2 wheel on seeded random inputs."""
import numpy as np
import pytest

from oracle import restate as R

cv = pytest.importorskip("cv2")


@pytest.mark.parametrize("seed", range(20))
def test_oracle_vs_cv2_random(seed):
    rng = np.random.default_rng(seed)
    h, w = int(rng.integers(3, 150)), int(rng.integers(3, 220))
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    if seed % 2:
        img = np.repeat(np.repeat(img[::7, ::9], 7, axis=0)[:h], 9, axis=1)[:, :w].copy()  # blocky: real edges
        h, w = img.shape[:2]
    g = cv.cvtColor(img, cv.COLOR_BGR2GRAY)
    assert np.array_equal(g, R.bgr2gray(img))
    b = cv.GaussianBlur(g, (5, 5), 0)
    assert np.array_equal(b, R.gauss5(g))
    assert np.array_equal(cv.normalize(b, None, 0, 255, cv.NORM_MINMAX), R.normalize_minmax(b))
    lo, hi = sorted(rng.integers(0, 400, 2).tolist())
    e = cv.Canny(b, lo, hi)
    assert np.array_equal(e, R.canny(b, lo, hi))
    thr = int(rng.integers(1, 30))
    ref = cv.HoughLinesWithAccumulator(e, 1, np.pi / 180, thr)
    lines, votes = R.hough_lines(e, 1, np.pi / 180, thr)
    if ref is None:
        assert lines is None
    else:
        assert np.array_equal(ref[:, 0, :2], lines[:, 0]) and np.array_equal(ref[:, 0, 2].astype(np.int32), votes)
    mlen, gap = int(rng.integers(0, 30)), int(rng.integers(0, 10))
    ref = cv.HoughLinesP(e, 1, np.pi / 180, thr, minLineLength=mlen, maxLineGap=gap)
    got = R.hough_lines_p(e, 1, np.pi / 180, thr, mlen, gap)
    assert (ref is None and got is None) or np.array_equal(ref, got)
    if h >= 8 and w >= 8:
        refine = seed % 2
        d = cv.createLineSegmentDetector(refine, scale=.8, sigma_scale=.6, quant=2.0, ang_th=22.5, log_eps=0,
                                         density_th=.7, n_bins=1024)
        ref = d.detect(g)[0]
        got = R.lsd(g, refine)[0]
        assert (ref is None and got is None) or np.array_equal(ref, got)
