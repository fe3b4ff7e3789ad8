"""CPU: LSD_REFINE_ADV of the oracle against cv2 4.13 -- the golden vectors made with the
reference's own parameters (tests/golden/make_golden_lsd2.py) and the live wheel.

PARTIALLY PINNED: nfa() is exact; the rectangle's pixel enumeration reproduces cv2's counts
for 98.4 % of the rectangles, so a few per cent of the segments differ.  The bar written here: >= 99 % of cv2's
segments appear bit for bit in the oracle's output overall, >= 97 % of the oracle's appear in cv2's, never below
90 % on one image; for the common segments width / prec / NFA are identical for >= 96 % (measured on the six golden images, both
scales: 1364 of cv2's 1371 segments found, 1391 produced, 1322 with identical attributes)."""
import os

import numpy as np
import pytest

from oracle import restate as R

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
IMAGES = ["demo_scarce", "sc_white_2", "sc_minecraft_cave", "old_sc_inf", "sc_cube", "sc_scarce_3_flat"]


@pytest.fixture(scope="module")
def golden_lsd2():
    return dict(np.load(os.path.join(GOLDEN, "golden_lsd2.npz")))


def agreement(lines, width, prec, nfa, ref_lines, ref_w, ref_p, ref_nfa):
    mine = {tuple(l): (w, p, n) for l, w, p, n in zip(lines.tolist(), width, prec, nfa)}
    ref = {tuple(l): (w, p, n) for l, w, p, n in zip(ref_lines.tolist(), ref_w, ref_p, ref_nfa)}
    common = [k for k in ref if k in mine]
    same_attr = sum(mine[k][0] == ref[k][0] and mine[k][1] == ref[k][1] and abs(mine[k][2] - ref[k][2]) <= 1e-9
                    for k in common)
    return len(ref), len(mine), len(common), same_attr


def test_lsd_refine2_golden(golden_inputs, golden_lsd2):
    tot = np.zeros(4, np.int64)
    for name in IMAGES:
        g = R.bgr2gray(golden_inputs[name])
        for scale, key in ((.5, "h"), (.8, "f")):
            lines, width, prec, nfa = R.lsd(g, 2, scale, return_nfa=True)
            o = lambda s: golden_lsd2[f"{name}/lsd2{key}{s}"]  # noqa: E731
            a = agreement(lines[:, 0], width[:, 0], prec[:, 0], nfa[:, 0], o(""), o("_w"), o("_p"), o("_nfa"))
            assert a[2] >= 0.90 * a[0] and a[2] >= 0.90 * a[1], (name, scale, a)
            tot += a
    n_ref, n_mine, common, same = tot
    assert common >= 0.99 * n_ref, tot
    assert common >= 0.97 * n_mine, tot
    assert same >= 0.96 * common, tot  # (the NFA differs where the pixel count differs by a pixel or a row)


def test_lsd_refine2_nfa_formula_is_exact():
    """Every NFA cv2 returns for an unimproved rectangle (prec == 0.125) is nfa(n, k) of an integer pair."""
    cv = pytest.importorskip("cv2")
    import ctypes as C
    import math
    lib = R.lib()
    lib.or_lsd_nfa.restype = C.c_double
    rng = np.random.default_rng(3)
    img = np.full((160, 240), 40, np.uint8)
    for _ in range(12):
        p0, p1 = rng.integers(10, 230, 2), rng.integers(10, 150, 2)
        cv.line(img, (int(p0[0]), int(p1[0])), (int(p0[1]), int(p1[1])), int(rng.integers(120, 255)), int(rng.integers(1, 4)))
    d = cv.createLineSegmentDetector(2, scale=.8, sigma_scale=.6, quant=2.0, ang_th=22.5, log_eps=0, density_th=.7,
                                     n_bins=1024)
    lines, width, prec, nfa = d.detect(img)
    log_nt = 5 * (math.log10(192.0) + math.log10(128.0)) / 2 + math.log10(11.0)
    table = {round(lib.or_lsd_nfa(C.c_int(n), C.c_int(k), C.c_double(.125), C.c_double(log_nt)), 8)
             for n in range(1, 700) for k in range(0, n + 1)}
    vals = [v for v, p in zip(nfa[:, 0], prec[:, 0]) if p == 0.125]
    assert len(vals) >= 5 and all(round(v, 8) in table for v in vals)
