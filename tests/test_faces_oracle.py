"""CPU: the quad-search oracle (oracle/geometry.py get_faces_from_pairs) against the reference's own
outputs (tests/golden/golden_faces.npz, made by tests/golden/make_golden_faces.py)."""
import os

import numpy as np
import pytest

from oracle import geometry as G

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
IMAGES = ["demo_scarce", "sc_cube", "sc_white_2", "sc_scarce_3_flat", "old_sc_inf"]
COMBOS = (("zfaces", ("x", "y")), ("yfaces", ("z", "x")), ("xfaces", ("y", "z")))


@pytest.fixture(scope="module")
def golden_faces():
    return dict(np.load(os.path.join(GOLDEN, "golden_faces.npz")))


@pytest.mark.parametrize("name", IMAGES)
def test_faces_golden(name, golden_faces):
    """Same faces in the same order and orientation; the reference computes float32 segments in float32, the
    oracle in float64: corners within 5e-3 px."""
    for key, (a, b) in COMBOS:
        faces, quads, keep = G.get_faces_from_pairs(golden_faces[f"{name}/{a}"], golden_faces[f"{name}/{b}"])
        ref = golden_faces[f"{name}/{key}"]
        assert faces.shape == ref.shape, (name, key)
        assert np.allclose(faces, ref, rtol=0, atol=5e-3)
        assert len(quads) == len(keep) and keep.sum() == len(ref)


def test_faces_duplicate_rule_and_orientation():
    # two unit-ish squares sharing their two horizontal edges (same (i, k)), the second one narrower: the wider is
    # dropped (larger circumference, overlapping), whatever the order of the vertical edges
    h = np.array([[0, 0, 100, 0], [0, 50, 100, 50]], float)
    v = np.array([[0, 0, 0, 50], [60, 0, 60, 50], [100, 0, 100, 50]], float)
    faces, quads, keep = G.get_faces_from_pairs(h, v, threshold=1)
    assert [tuple(q) for q in quads] == [(0, 0, 1, 1), (0, 0, 1, 2), (0, 1, 1, 2)]
    assert keep.tolist() == [True, False, True]
    assert len(faces) == 2
