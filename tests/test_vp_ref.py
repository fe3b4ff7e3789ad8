"""CPU: oracle/vp_ref.py (restatement of the reference's pure-Python VP functions) against the reference's own
outputs in tests/golden/golden_vp.npz (made by tests/golden/make_golden.py importing /root/reference)."""
import os

import numpy as np
import pytest

from oracle import vp_ref

GOLD = os.path.join(os.path.dirname(__file__), "golden", "golden_vp.npz")


@pytest.fixture(scope="module")
def g():
    return np.load(GOLD)


def test_polar_lines_equal_reference(g):
    for pairs, polar in ((g["pairs_lsd"], g["polar_lsd"]), (g["pairs_prob"], g["polar_prob"])):
        edges = [(p[:2], p[2:]) for p in pairs]
        assert np.array_equal(vp_ref.edges_to_polar_lines(edges), polar)


def test_focal_points_and_losses_equal_reference(g):
    for (phi, theta), fp, ll, lp in zip(g["cands"], g["focal_points"], g["sum_loss_lsd"], g["sum_loss_prob"]):
        assert np.array_equal(np.array(vp_ref.get_focal_points(phi, theta)), fp)
        assert vp_ref.sum_loss(phi, theta, g["polar_lsd"]) == ll
        assert vp_ref.sum_loss(phi, theta, g["polar_prob"]) == lp
        assert vp_ref.sum_loss_vec(phi, theta, g["polar_lsd"]) == pytest.approx(ll, rel=1e-12)


def test_seeded_search_reproduces_reference_runs(g):
    """np.random.seed(s) + the reference's draw order: the restated search lands on the reference's own result."""
    runs = g["regress_runs_lsd"]
    for seed in (0, 3):
        (phi, theta), loss = vp_ref.regress_lines(g["polar_lsd"], 1000, 500, np.deg2rad(15), seed=seed,
                                                  loss=vp_ref.sum_loss_vec)
        assert (phi, theta) == pytest.approx(tuple(runs[seed, :2]), abs=1e-9)
        assert loss == pytest.approx(runs[seed, 2], rel=1e-9)
    (phi, theta), loss = vp_ref.regress_lines(g["polar_lsd"], 60, 20, np.deg2rad(15), seed=1)
    (phi2, theta2), loss2 = vp_ref.regress_lines(g["polar_lsd"], 60, 20, np.deg2rad(15), seed=1, loss=vp_ref.sum_loss_vec)
    assert (phi, theta) == (phi2, theta2) and loss == pytest.approx(loss2, rel=1e-12)
