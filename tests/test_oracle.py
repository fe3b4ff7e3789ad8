"""CPU: the oracle (oracle/cv_restate.c) against the golden vectors and against the live cv2 wheel."""
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import restate as R

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
IMAGES = ["demo_scarce", "sc_white_2", "sc_minecraft_cave", "old_sc_inf", "sc_cube", "sc_scarce_3_flat"]


def sha(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()[:10]


@pytest.mark.parametrize("name", IMAGES)
def test_front_chain_golden(name, golden_inputs, golden_outputs):
    img, o = golden_inputs[name], lambda k: golden_outputs[f"{name}/{k}"]
    g = R.bgr2gray(img)
    assert np.array_equal(g, o("gray"))
    b = R.gauss5(g)
    assert np.array_equal(b, o("blur"))
    nm = R.normalize_minmax(b)
    assert np.array_equal(nm, o("norm"))
    assert np.array_equal(R.canny(g, 5, 150), o("canny_prob"))
    assert np.array_equal(R.canny(b, 5, 10), o("canny_blur"))
    assert np.array_equal(R.canny(nm, 30, 50), o("docanny"))
    assert np.array_equal(R.canny(nm, 50, 30), o("docanny"))  # lo > hi are swapped


@pytest.mark.parametrize("name", IMAGES)
def test_hough_golden(name, golden_outputs):
    o = lambda k: golden_outputs[f"{name}/{k}"]  # noqa: E731
    for edges, thr, key in ((o("canny_blur"), 50, "hough50"), (o("docanny"), 70, "hough70")):
        lines, votes, accum = R.hough_lines(edges, 1, np.pi / 180, thr, return_accum=True)
        ref = o(key)
        assert np.array_equal(lines[:, 0], ref[:, 0, :2])
        assert np.array_equal(votes, ref[:, 0, 2].astype(np.int32))
        assert accum.sum() == 180 * int((edges > 0).sum())
    assert np.array_equal(R.hough_lines_p(o("canny_prob"), 1, np.pi / 180, 30, 50, 10), o("houghp"))
    assert np.array_equal(R.hough_lines_p(o("docanny"), 1, np.pi / 180, 30, 20, 3), o("houghp2"))


@pytest.mark.parametrize("name", IMAGES)
def test_lsd_golden(name, golden_outputs):
    o = lambda k: golden_outputs[f"{name}/{k}"]  # noqa: E731
    for refine, scale, key in ((0, .8, "lsd0"), (1, .8, "lsd1"), (0, .5, "lsd0h")):
        lines, width, prec = R.lsd(o("gray"), refine, scale)
        assert np.array_equal(lines, o(key))
        if key != "lsd0h":
            assert np.array_equal(width, o(key + "_w"))
            assert np.array_equal(prec, o(key + "_p"))


def test_golden_table_hashes(golden_inputs):
    """Appendix-B style table: the oracle reproduces the SHA-1 of every chain for the stored inputs."""
    table = json.load(open(os.path.join(GOLDEN, "golden_table.json")))["images"]
    for name, img in golden_inputs.items():
        row = table[name]
        g = R.bgr2gray(img)
        assert sha(g) == row["gray"]
        e = R.canny(g, 5, 150)
        assert [int((e > 0).sum()), sha(e)] == row["canny_5_150"]
        hp = R.hough_lines_p(e, 1, np.pi / 180, 30, 50, 10)
        assert [len(hp), sha(hp)] == row["houghp_30_50_10"]
        eb = R.canny(R.gauss5(g), 5, 10)
        assert [int((eb > 0).sum()), sha(eb)] == row["blur_canny_5_10"]
        hl, _ = R.hough_lines(eb, 1, np.pi / 180, 50)
        assert [len(hl), sha(hl)] == row["hough_50"]
        ls = R.lsd(g, 0, .8)[0]
        assert [len(ls), sha(ls)] == row["lsd_0_08"]


def test_edge_cases():
    blank = np.zeros((40, 60), np.uint8)
    assert R.hough_lines(blank, 1, np.pi / 180, 10)[0] is None
    assert R.hough_lines_p(blank, 1, np.pi / 180, 10, 5, 1) is None
    assert R.lsd(blank)[0] is None
    assert not R.canny(blank, 5, 10).any()
    const = np.full((16, 16), 77, np.uint8)
    assert np.array_equal(R.normalize_minmax(const), np.zeros_like(const))  # max == min -> scale 0


def test_oracle_vs_live_cv2():
    """Differential test on synthetic frames (sizes not multiple of 4/8, random noise) against cv2."""
    cv = pytest.importorskip("cv2")
    import graphics_project_b200.synth as synth
    rng = np.random.default_rng(0)
    cases = [synth.make_frame(1, 270, 480), synth.make_frame(2, 203, 317, "cave"),
             rng.integers(0, 256, (64, 97, 3), dtype=np.uint8)]
    for img in cases:
        g = cv.cvtColor(img, cv.COLOR_BGR2GRAY)
        assert np.array_equal(g, R.bgr2gray(img))
        b = cv.GaussianBlur(g, (5, 5), 0)
        assert np.array_equal(b, R.gauss5(g))
        assert np.array_equal(cv.normalize(b, None, 0, 255, cv.NORM_MINMAX), R.normalize_minmax(b))
        for src, lo, hi in ((g, 5, 150), (b, 5, 10), (b, 30, 50)):
            e = cv.Canny(src, lo, hi)
            assert np.array_equal(e, R.canny(src, lo, hi))
        e = cv.Canny(b, 5, 10)
        ref = cv.HoughLinesWithAccumulator(e, 1, np.pi / 180, 40)
        lines, votes = R.hough_lines(e, 1, np.pi / 180, 40)
        if ref is None:
            assert lines is None
        else:
            assert np.array_equal(ref[:, 0, :2], lines[:, 0]) and np.array_equal(ref[:, 0, 2].astype(np.int32), votes)
        e2 = cv.Canny(g, 5, 150)
        ref = cv.HoughLinesP(e2, 1, np.pi / 180, 30, minLineLength=20, maxLineGap=3)
        got = R.hough_lines_p(e2, 1, np.pi / 180, 30, 20, 3)
        assert (ref is None and got is None) or np.array_equal(ref, got)
        for refine in (0, 1):
            d = cv.createLineSegmentDetector(refine, scale=.8, sigma_scale=.6, quant=2.0, ang_th=22.5, log_eps=0,
                                             density_th=.7, n_bins=1024)
            ref = d.detect(g)[0]
            got = R.lsd(g, refine)[0]
            assert (ref is None and got is None) or np.array_equal(ref, got)


@pytest.mark.parametrize("name", ["demo_scarce", "sc_white_2", "old_sc_inf", "sc_cube", "sc_scarce_3_flat"])
def test_geometry_oracle_vs_reference_merge(name, golden_outputs):
    """oracle/geometry.py against util.combineParallelLines run by tests/golden/make_golden_merge.py."""
    import os
    from oracle import geometry as G
    gm = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_merge.npz"))
    for det, tol in (("prob", 1e-9), ("lsd", 1e-3)):
        out, src, merged = G.combine_parallel_lines(golden_outputs[f"{name}/ref_{det}"].reshape(-1, 4))
        ref = gm[f"{name}/merge_{det}"]
        assert out.shape == ref.shape and np.allclose(out, ref, rtol=0, atol=tol)
        assert len(src) == len(out) and merged.any()
