"""oracle/vp_ref.py -- TEST INFRASTRUCTURE ONLY (never imported by the product path).

Restatement of the reference's pure-Python vanishing-point functions, used by the tests, by
``__graft_entry__.smoke()`` and by ``bench.py``'s CPU legs (the ``lines`` workload):

  edges_to_polar_lines   opencv_renewed.py:16-17 (and the inline copy with the 10 px filter,
                         opencv_fit_color.py:88)
  get_focal_points       opencv_fit_color.py:95-101
  loss_function          opencv_fit_color.py:14-15
  min_loss / sum_loss    opencv_fit_color.py:17-26
  regress_lines          opencv_fit_color.py:28-47   (random search; here with an explicit seed)

``regress_lines`` keeps the reference's SCALAR Python loops (two ``sum_loss`` calls per improving draw, list
comprehensions over the lines): this is what the reference executes, so it is what the CPU baseline times.
``sum_loss_vec`` is the same arithmetic vectorised over the lines, for scoring engine results quickly; the two
agree to summation order (checked in tests/test_vp_ref.py).

The reference hard-codes WIDTH = 600, HEIGHT = 400 (opencv_fit_color.py:7-8), its fixture size; ``width`` /
``height`` are parameters here so that both bench arms can fit 1080p frames with the frame's own size.

Pinned: ``tests/golden/golden_vp.npz`` holds the reference's own outputs (polar lines, focal points, losses and
seeded ``regress_lines`` runs, made by tests/golden/make_golden.py importing /root/reference); tests/test_vp_ref.py
checks this file against them.
"""
import numpy as np

WIDTH = 600
HEIGHT = 400


def edges_to_polar_lines(edges, min_len=None):
    """[(a, b), ...] endpoint pairs -> (N, 2) array of (rho, phi); ``min_len`` adds get_camera_angles' filter."""
    out = [(np.sign(np.arctan2(b[0] - a[0], b[1] - a[1])) * (a[1] * b[0] - b[1] * a[0]) / np.linalg.norm(b - a),
            np.fmod(-np.arctan2(b[0] - a[0], b[1] - a[1]) + np.pi, np.pi))
           for a, b in edges if min_len is None or np.linalg.norm(b - a) > min_len]
    return np.array(out)


def get_focal_points(phi, theta, width=WIDTH, height=HEIGHT):
    sc_points = [(1 / (np.tan(theta) * np.sin(phi)), 1 / np.tan(phi)),
                 (0, -np.tan(phi)),
                 (-np.tan(theta) / np.sin(phi), 1 / np.tan(phi))]
    return [np.array([height / 2 * p[0] + width / 2, height / 2 * p[1] + height / 2]) for p in sc_points]


def loss_function(point, rho, phi):
    return pow(point[1] * np.sin(phi) + point[0] * np.cos(phi) - rho, 2)


def min_loss(points, lines):
    return sum([min([loss_function(point, rho, phi) for point in points]) for rho, phi in lines]) / len(lines)


def sum_loss(phi, theta, lines, width=WIDTH, height=HEIGHT):
    return min_loss(get_focal_points(phi, theta, width, height), lines)


def sum_loss_vec(phi, theta, lines, width=WIDTH, height=HEIGHT):
    """sum_loss with the loop over the lines vectorised (same operations per line)."""
    lines = np.asarray(lines, np.float64).reshape(-1, 2)
    rho, ph = lines[:, 0], lines[:, 1]
    with np.errstate(all="ignore"):
        pts = get_focal_points(phi, theta, width, height)
        per = np.stack([(p[1] * np.sin(ph) + p[0] * np.cos(ph) - rho) ** 2 for p in pts])
        return float(per.min(0).sum() / len(lines))


def regress_lines(lines, iterations=500, refinement_iterations=100, refinement_area=0.3, seed=None, width=WIDTH,
                  height=HEIGHT, loss=sum_loss):
    """The reference's two-phase random search, draw for draw (np.random.rand order: phi then theta).  ``seed``
    seeds numpy's global generator like the golden runs did (np.random.seed); ``loss`` may be ``sum_loss_vec``."""
    if seed is not None:
        np.random.seed(seed)
    best_loss = 100000
    best_phi_theta = (0, 0)
    for _ in range(0, iterations):
        phi = np.random.rand() * np.pi
        theta = np.random.rand() * np.pi / 2
        if loss(phi, theta, lines, width, height) < best_loss:
            best_loss = loss(phi, theta, lines, width, height)
            best_phi_theta = (phi, theta)
    current_phi, current_theta = best_phi_theta
    for _ in range(0, refinement_iterations):
        phi = np.random.rand() * refinement_area + current_phi - refinement_area / 2
        theta = np.random.rand() * refinement_area + current_theta - refinement_area / 2
        if loss(phi, theta, lines, width, height) < best_loss:
            best_loss = loss(phi, theta, lines, width, height)
            best_phi_theta = (phi, theta)
    return best_phi_theta, best_loss
