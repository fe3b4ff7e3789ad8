"""CPU: LSD_REFINE_ADV of the oracle against cv2 4.13 -- the golden vectors made with the
reference's own parameters (tests/golden/make_golden_lsd2.py) and the live wheel.  PINNED: lines, widths,
precisions and NFA values are bit-identical, including order, on the reference's fixture images at both scales
(the simulator runs ``lsd(image, 2, scale=0.5)``, opencv.py:142-145) and on seeded synthetic frames."""
import importlib.util
import os

import numpy as np
import pytest

from oracle import restate as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
IMAGES = ["demo_scarce", "sc_white_2", "sc_minecraft_cave", "old_sc_inf", "sc_cube", "sc_scarce_3_flat"]


@pytest.fixture(scope="module")
def golden_lsd2():
    return dict(np.load(os.path.join(GOLDEN, "golden_lsd2.npz")))


@pytest.mark.parametrize("name", IMAGES)
def test_lsd_refine2_golden(name, golden_inputs, golden_lsd2):
    g = R.bgr2gray(golden_inputs[name])
    for scale, key in ((.5, "h"), (.8, "f")):
        lines, width, prec, nfa = R.lsd(g, 2, scale, return_nfa=True)
        o = lambda s: golden_lsd2[f"{name}/lsd2{key}{s}"]  # noqa: E731
        assert np.array_equal(lines[:, 0], o("")), (name, scale)
        assert np.array_equal(width[:, 0], o("_w"))
        assert np.array_equal(prec[:, 0], o("_p"))
        assert np.array_equal(nfa[:, 0], o("_nfa"))
    assert np.array_equal(golden_lsd2[f"{name}/ref_lsd2"], golden_lsd2[f"{name}/lsd2h"])  # the reference's own wrapper


@pytest.mark.parametrize("seed", range(8))
def test_lsd_refine2_vs_cv2_synthetic(seed):
    cv = pytest.importorskip("cv2")
    spec = importlib.util.spec_from_file_location("le_synth", os.path.join(ROOT, "graphics-project_b200", "synth.py"))
    synth = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(synth)
    kind = ("cubes", "voxel", "cave")[seed % 3]
    h, w = [(270, 480), (400, 600), (333, 517), (135, 241)][seed % 4]
    g = cv.cvtColor(synth.make_frame(seed, h, w, kind), cv.COLOR_BGR2GRAY)
    for scale in (.5, .8):
        d = cv.createLineSegmentDetector(2, scale=scale, sigma_scale=.6, quant=2.0, ang_th=22.5, log_eps=0,
                                         density_th=.7, n_bins=1024)
        want = d.detect(g)
        got = R.lsd(g, 2, scale, return_nfa=True)
        if want[0] is None:
            assert got[0] is None
            continue
        for a, b in zip(got, want):
            assert np.array_equal(a, b), (seed, scale)


def test_lsd_refine2_nfa_formula_is_exact():
    """Every NFA cv2 returns for an unimproved rectangle (prec == 0.125) is nfa(n, k) of an integer pair: this is
    how the rectangle's pixel enumeration was identified."""
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
    got = R.lsd(img, 2, .8, return_nfa=True)
    assert all(np.array_equal(a, b) for a, b in zip(got, (lines, width, prec, nfa)))
