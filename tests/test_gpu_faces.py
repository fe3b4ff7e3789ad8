"""GPU parity: get_faces_from_pairs through the C ABI vs the reference's own outputs
(tests/golden/golden_faces.npz) and vs the float64 oracle on seeded random segment sets (bit-exact quads / keep
flags; corners to 1e-9)."""
import os

import numpy as np
import pytest
import torch

from oracle import geometry as G
from test_faces_oracle import COMBOS, IMAGES

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def golden_faces():
    return dict(np.load(os.path.join(GOLDEN, "golden_faces.npz")))


def _pairs(a):
    return [(np.array(r[0:2]), np.array(r[2:4])) for r in a.reshape(-1, 4)]


@pytest.mark.parametrize("name", IMAGES)
def test_get_faces_from_pairs_golden(name, golden_faces):
    import graphics_project_b200.vp as V
    probs = [(_pairs(golden_faces[f"{name}/{a}"]), _pairs(golden_faces[f"{name}/{b}"])) for _, (a, b) in COMBOS]
    batch = V.get_faces_from_pairs_batch(probs)  # the three calls of one frame in one launch
    for (key, _), (e1, e2), got_b in zip(COMBOS, probs, batch):
        ref = golden_faces[f"{name}/{key}"]
        got = V.get_faces_from_pairs(e1, e2)
        assert got == got_b
        assert isinstance(got, list) and all(len(f) == 4 and isinstance(f[0], tuple) for f in got)
        arr = np.array(got, dtype=np.float64).reshape(-1, 4, 2)
        assert arr.shape == ref.shape, (name, key)
        assert np.allclose(arr, ref, rtol=0, atol=5e-3)


def _random_lists(rng, n1, n2):
    """Grid-like clutter: many segments end near each other, so quads, shared pairs and overlaps all occur."""
    def segs(n, horizontal):
        out = np.zeros((n, 4))
        for k in range(n):
            c = rng.integers(0, 7, 2) * 60.0 + rng.normal(0, 3, 2)
            ln = rng.integers(1, 4) * 60.0
            d = np.array([ln, rng.normal(0, 4)]) if horizontal else np.array([rng.normal(0, 4), ln])
            out[k] = (c[0], c[1], c[0] + d[0], c[1] + d[1])
        return out
    return segs(n1, True), segs(n2, False)


@pytest.mark.parametrize("seed", range(6))
def test_face_quads_random_vs_oracle(engine, seed):
    rng = np.random.default_rng(seed)
    probs = 4
    max1, max2 = 40, 37
    c1 = rng.integers(0, max1 + 1, probs)
    c2 = rng.integers(0, max2 + 1, probs)
    c1[0], c2[0] = max1, max2
    c1[1] = 0
    s1 = np.zeros((probs, max1, 4))
    s2 = np.zeros((probs, max2, 4))
    for p in range(probs):
        a, b = _random_lists(rng, c1[p], c2[p])
        s1[p, :c1[p]], s2[p, :c2[p]] = a, b
    thr = (15, 8, 25)[seed % 3]
    quads, faces, keep, counts = engine.face_quads(torch.from_numpy(s1), torch.from_numpy(s2),
                                                   torch.from_numpy(c1.astype(np.int32)),
                                                   torch.from_numpy(c2.astype(np.int32)), thr, max_out=1 << 15)
    quads, faces, keep, counts = quads.cpu().numpy(), faces.cpu().numpy(), keep.cpu().numpy(), counts.cpu().numpy()
    total = 0
    for p in range(probs):
        rf, rq, rk = G.get_faces_from_pairs(s1[p, :c1[p]], s2[p, :c2[p]], thr)
        n = int(counts[p])
        assert n == len(rq), (seed, p)
        assert np.array_equal(quads[p, :n], rq)
        assert np.array_equal(keep[p, :n].astype(bool), rk)
        got = faces[p, :n][rk]
        assert got.shape == rf.shape
        assert np.allclose(got, rf, rtol=0, atol=1e-9, equal_nan=True)
        total += n
    assert total > 0
    engine.check()


def test_face_quads_truncation_reports_true_count(engine):
    rng = np.random.default_rng(9)
    a, b = _random_lists(rng, 40, 40)
    q, f, k, c = engine.face_quads(torch.from_numpy(a[None]), torch.from_numpy(b[None]), None, None, 25, max_out=8)
    rf, rq, rk = G.get_faces_from_pairs(a, b, 25)
    assert int(c[0]) == len(rq) > 8
    assert np.array_equal(q[0].cpu().numpy(), rq[:8])
