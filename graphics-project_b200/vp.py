"""Vanishing-point fit with the reference's signatures (opencv_fit_color.py, opencv_renewed.py:16-32).

The reference's ``regress_lines`` is an unseeded random search: 1500 Python evaluations of
``sum_loss`` (~1 s for 75 lines).  Here the same loss is evaluated for whole candidate sets on the
GPU (``le_vp_loss``): a dense deterministic grid over the reference's search domain
(phi in [0,pi), theta in [0,pi/2)), seeds from the Gaussian-sphere histogram of pairwise
interpretation-plane intersections (``le_vp_sphere_hist``), then shrinking refinement grids around
the best candidate.  Returned values keep the reference's types: ``((phi, theta), loss)``.
"""
import numpy as np
import torch

from .engine import default_engine
from . import detectors

HEIGHT = 400
WIDTH = 600


def edges_to_polar_lines(edges):
    """opencv_renewed.py:16-17: list of (a, b) endpoint pairs -> (N, 2) array of (rho, phi)."""
    if len(edges) == 0:
        return np.zeros((0,), np.float64)  # np.array([]) in the reference
    segs = np.array([[a[0], a[1], b[0], b[1]] for a, b in edges], dtype=np.float64)
    eng = default_engine()
    return eng.segments_to_polar(torch.from_numpy(segs)).cpu().numpy()


def get_focal_points(phi, theta, width=WIDTH, height=HEIGHT):
    """opencv_fit_color.py:95-101 (scalar; three vanishing points in pixels)."""
    sc_points = [(1 / (np.tan(theta) * np.sin(phi)), 1 / np.tan(phi)),
                 (0, -np.tan(phi)),
                 (-np.tan(theta) / np.sin(phi), 1 / np.tan(phi))]
    return [np.array([height / 2 * p[0] + width / 2, height / 2 * p[1] + height / 2]) for p in sc_points]


def loss_function(point, rho, phi):
    """opencv_fit_color.py:14-15."""
    return pow(point[1] * np.sin(phi) + point[0] * np.cos(phi) - rho, 2)


def losses(lines, candidates, width=WIDTH, height=HEIGHT):
    """sum_loss for every (phi, theta) row of ``candidates`` -> float64 array (GPU)."""
    eng = default_engine()
    lines_t = torch.as_tensor(np.asarray(lines, dtype=np.float64))
    cand_t = torch.as_tensor(np.asarray(candidates, dtype=np.float64))
    return eng.vp_loss(lines_t, cand_t, width, height).cpu().numpy()


def sum_loss(phi, theta, lines, width=WIDTH, height=HEIGHT):
    """opencv_fit_color.py:25-26."""
    return float(losses(lines, [[phi, theta]], width, height)[0])


def sphere_seeds(lines, width=WIDTH, height=HEIGHT, n_az=360, n_el=90, top=8):
    """(phi, theta) hypotheses from the Gaussian-sphere histogram peaks.

    A polar line x cos(phi) + y sin(phi) = rho is turned back into two pixel points; the histogram
    votes every pairwise intersection direction v (folded to v_z >= 0).  A peak direction is read as
    the 'x' or the 'z' vanishing point of get_focal_points: d_x = (1/(tan t sin p), 1/tan p, 1),
    d_z = (-tan t / sin p, 1/tan p, 1)."""
    lines = np.asarray(lines, dtype=np.float64).reshape(-1, 2)
    if len(lines) < 2:
        return np.zeros((0, 2))
    rho, phi = lines[:, 0], lines[:, 1]
    p0 = np.stack([rho * np.cos(phi), rho * np.sin(phi)], 1)
    d = np.stack([-np.sin(phi), np.cos(phi)], 1) * 100.0
    segs = np.concatenate([p0 - d, p0 + d], 1)
    eng = default_engine()
    hist = eng.vp_sphere_hist(torch.from_numpy(segs), width, height, n_az, n_el)
    flat = hist.flatten()
    k = min(top, flat.numel())
    idx = torch.sort(flat, descending=True, stable=True).indices[:k].cpu().numpy()  # ties: lower bin first
    seeds = []
    for i in idx:
        be, ba = divmod(int(i), n_az)
        az = (ba + 0.5) / n_az * 2 * np.pi - np.pi
        el = (be + 0.5) / n_el * (np.pi / 2)
        vx, vy, vz = np.cos(el) * np.cos(az), np.cos(el) * np.sin(az), np.sin(el)
        if vz < 1e-6:
            continue
        x, y = vx / vz, vy / vz
        ph = np.arctan2(1.0, y) % np.pi  # 1/tan(phi) = y
        if abs(x) < 1e-9 or abs(np.sin(ph)) < 1e-9:
            continue
        th_x = np.arctan(1.0 / (x * np.sin(ph)))   # x = 1/(tan t sin p)
        th_z = np.arctan(-x * np.sin(ph))          # x = -tan t / sin p
        for th in (th_x, th_z):
            if 0 <= th < np.pi / 2:
                seeds.append((ph, th))
    return np.array(seeds).reshape(-1, 2)


def _grid(phi_lo, phi_hi, th_lo, th_hi, n_phi, n_th):
    ph = phi_lo + (np.arange(n_phi) + 0.5) / n_phi * (phi_hi - phi_lo)
    th = th_lo + (np.arange(n_th) + 0.5) / n_th * (th_hi - th_lo)
    return np.stack(np.meshgrid(ph, th, indexing="ij"), -1).reshape(-1, 2)


_ROUNDS = 4


def _search_grid(iterations, refinement_iterations):
    """(n_phi, n_th, n_ref): grid sizes of the two search phases; the reference's draw counts are lower bounds."""
    n_phi = max(192, int(np.ceil(np.sqrt(2 * max(iterations, 1)))))
    return n_phi, max(96, n_phi // 2), max(32, int(np.ceil(np.sqrt(max(refinement_iterations, 1)))))


def regress_lines_batch(polar, pcounts, iterations=500, refinement_iterations=100, refinement_area=0.3, width=WIDTH,
                        height=HEIGHT, want_hist=True, engine=None):
    """``regress_lines`` for every frame of a batch, one CTA per frame and no host round trip (``le_vp_fit_batch``):
    polar [N, max, 2] float64 CUDA tensor of (rho, phi) lines, pcounts [N] -> (fit [N, 3] = (phi, theta, loss),
    hist [N, 90, 360] Gaussian-sphere histograms or None).  Same candidates, in the same order, as
    ``regress_lines`` evaluates step by step."""
    eng = engine or default_engine()
    n_phi, n_th, n_ref = _search_grid(iterations, refinement_iterations)
    return eng.vp_fit_batch(polar, pcounts, width, height, n_phi, n_th, n_ref, _ROUNDS, float(refinement_area),
                            want_hist=want_hist)


def get_camera_angles_batch(frames, iterations=1000, refinement_iterations=500, width=None, height=None, max_segs=2048,
                            engine=None):
    """``get_camera_angles(image, method='lsd')`` (opencv_fit_color.py:71-93) for a batch of BGR frames
    ([N, H, W, 3] uint8 CUDA tensor): gray -> LSD -> segments longer than 10 px as polar lines -> fit, all on the
    device.  ``width`` / ``height`` default to the frame size (the reference's WIDTH / HEIGHT constants are its
    600 x 400 image size).  Returns (fit [N, 3] = (phi, theta, loss), hist [N, 90, 360], polar, pcounts, segs,
    counts); counts holds the true LSD segment count (> max_segs means truncated)."""
    eng = engine or default_engine()
    h, w = frames.shape[1], frames.shape[2]
    gray = eng.bgr2gray(frames)
    segs, _, _, counts = eng.lsd(gray, max_segs=max_segs)
    polar, pcounts = eng.polar_batch(segs, counts, 10.0)
    fit, hist = regress_lines_batch(polar, pcounts, iterations, refinement_iterations, 0.3,
                                    float(w if width is None else width), float(h if height is None else height),
                                    engine=eng)
    return fit, hist, polar, pcounts, segs, counts


def regress_lines(lines, iterations=500, refinement_iterations=100, refinement_area=0.3, width=WIDTH,
                  height=HEIGHT, verbose=True, use_seeds=True):
    """opencv_fit_color.py:28-47 signature and return type; deterministic GPU search.

    One launch chain and one 24-byte read-back: the frame goes through the batched kernels (``le_vp_fit_batch``
    with n = 1 -- sphere histogram, seeds, grid and refinement boxes in one CTA), the same candidates in the same
    order as ``regress_lines_stepwise``.  ``iterations`` / ``refinement_iterations`` are lower bounds on the
    number of candidates per phase."""
    lines = np.asarray(lines, dtype=np.float64).reshape(-1, 2)
    if len(lines) == 0:
        raise ZeroDivisionError("division by zero")  # min_loss divides by len(lines) in the reference
    eng = default_engine()
    polar = torch.from_numpy(np.ascontiguousarray(lines))[None].to(eng.device)
    pcounts = torch.tensor([len(lines)], dtype=torch.int32, device=eng.device)
    fit, _ = regress_lines_batch(polar, pcounts, iterations, refinement_iterations, refinement_area, width, height,
                                 want_hist=bool(use_seeds), engine=eng)
    phi, theta, best_loss = (float(v) for v in fit[0].cpu().numpy())
    best = (phi, theta) if best_loss < 100000 else (0, 0)
    if best_loss >= 100000:
        best_loss = 100000
    if verbose:
        print(" The loss is ", best_loss)
    return best, best_loss


def regress_lines_stepwise(lines, iterations=500, refinement_iterations=100, refinement_area=0.3, width=WIDTH,
                  height=HEIGHT, verbose=True, use_seeds=True):
    """The search of ``regress_lines`` phase by phase from the host (``le_vp_loss`` per candidate set, five round
    trips): kept as the cross-check of the fused kernel (tests/test_gpu_lines_batch.py), not used by the wrappers."""
    lines = np.asarray(lines, dtype=np.float64).reshape(-1, 2)
    best_loss = 100000
    best = (0, 0)
    if len(lines) == 0:
        raise ZeroDivisionError("division by zero")  # min_loss divides by len(lines) in the reference
    n_phi, n_th, n_ref = _search_grid(iterations, refinement_iterations)
    cand = _grid(0, np.pi, 0, np.pi / 2, n_phi, n_th)
    seeds = sphere_seeds(lines, width, height) if use_seeds else []
    if len(seeds):
        cand = np.concatenate([cand, seeds])
    ls = losses(lines, cand, width, height)
    ls = np.where(np.isnan(ls), np.inf, ls)
    i = int(np.argmin(ls))
    if ls[i] < best_loss:
        best_loss, best = float(ls[i]), (float(cand[i, 0]), float(cand[i, 1]))
    side = float(refinement_area)
    for _ in range(_ROUNDS):  # reference does one refinement pass; extra passes only shrink the box
        cphi, cth = best
        cand = _grid(cphi - side / 2, cphi + side / 2, cth - side / 2, cth + side / 2, n_ref, n_ref)
        ls = losses(lines, cand, width, height)
        ls = np.where(np.isnan(ls), np.inf, ls)
        i = int(np.argmin(ls))
        if ls[i] < best_loss:
            best_loss, best = float(ls[i]), (float(cand[i, 0]), float(cand[i, 1]))
        side *= 4.0 / n_ref
    if verbose:
        print(" The loss is ", best_loss)
    return best, best_loss


def which_line(focal_points, line, threshold=700):
    """opencv_fit_color.py:49-59."""
    x_loss, y_loss, z_loss = (loss_function(focal_points[0], *line), loss_function(focal_points[1], *line),
                              loss_function(focal_points[2], *line))
    if min([x_loss, y_loss, z_loss]) >= threshold:
        return None
    if x_loss <= y_loss and x_loss <= z_loss:
        return "x"
    elif y_loss <= x_loss and y_loss <= z_loss:
        return "y"
    else:
        return "z"


def classifyEdges(edges, threshold_multiplier=1.2):
    """opencv_renewed.py:19-32."""
    lines = edges_to_polar_lines(edges)
    phi_theta, loss = regress_lines(lines, iterations=1000, refinement_iterations=500,
                                    refinement_area=np.deg2rad(15))
    phi, theta = phi_theta
    focal_points = get_focal_points(phi, theta)
    pairs = [(edge, which_line(focal_points, line, threshold=loss * threshold_multiplier))
             for edge, line in zip(edges, lines)]
    x_edges = [line for line, which in pairs if which == "x"]
    y_edges = [line for line, which in pairs if which == "y"]
    z_edges = [line for line, which in pairs if which == "z"]
    return x_edges, y_edges, z_edges, phi, theta


def smoothEdges(x_edges, y_edges, z_edges):
    """opencv_renewed.py:56-60: combineParallelLines on each of the three classified edge lists -- one batched
    launch (one CTA per list) instead of three recursive Python scans."""
    lists = [list(x_edges), list(y_edges), list(z_edges)]
    max_k = max(len(l) for l in lists)
    if max_k < 2:
        return tuple(lists)
    arr = np.zeros((3, max_k, 4), np.float64)
    for i, l in enumerate(lists):
        for k, e in enumerate(l):
            arr[i, k] = (e[0][0], e[0][1], e[1][0], e[1][1])
    eng = default_engine()
    counts = torch.tensor([len(l) for l in lists], dtype=torch.int32)
    out, src, oc = eng.combine_parallel_lines(torch.from_numpy(arr), counts)
    out, src, oc = out.cpu().numpy(), src.cpu().numpy(), oc.cpu().numpy()
    res = []
    for i, l in enumerate(lists):
        r = []
        for k in range(int(oc[i])):
            if src[i, k] & 0x40000000:
                r.append(np.array([[out[i, k, 0], out[i, k, 1]], [out[i, k, 2], out[i, k, 3]]]))
            else:
                r.append(l[int(src[i, k])])
        res.append(r)
    return tuple(res)


def _edge_array(edges, max_k):
    arr = np.zeros((max_k, 4), np.float64)
    for k, e in enumerate(edges):
        arr[k] = (e[0][0], e[0][1], e[1][0], e[1][1])
    return arr


def get_faces_from_pairs_batch(problems, threshold=15):
    """[(edges1, edges2), ...] -> [faces, ...]: every pair list of one frame (or of many frames) in ONE launch,
    one CTA per problem (the reference's pipelines call get_faces_from_pairs three times per frame,
    opencv_renewed.py:176-178)."""
    problems = [(list(a), list(b)) for a, b in problems]
    if not problems:
        return []
    max1 = max(1, max(len(a) for a, _ in problems))
    max2 = max(1, max(len(b) for _, b in problems))
    s1 = np.stack([_edge_array(a, max1) for a, _ in problems])
    s2 = np.stack([_edge_array(b, max2) for _, b in problems])
    c1 = torch.tensor([len(a) for a, _ in problems], dtype=torch.int32)
    c2 = torch.tensor([len(b) for _, b in problems], dtype=torch.int32)
    eng = default_engine()
    cap = 4096
    while True:
        quads, faces, keep, counts = eng.face_quads(torch.from_numpy(s1), torch.from_numpy(s2), c1, c2, threshold, cap)
        counts = counts.cpu().numpy()
        if counts.max() <= cap:
            break
        cap = int(counts.max())
    faces, keep = faces.cpu().numpy(), keep.cpu().numpy()
    out = []
    for p in range(len(problems)):
        fl = []
        for a in range(int(counts[p])):
            if not keep[p, a]:
                continue
            if np.isnan(faces[p, a]).any():  # lineIntersection returned None: the reference fails here too
                raise TypeError("'NoneType' object is not subscriptable (parallel neighbouring segments in a face)")
            fl.append([(float(x), float(y)) for x, y in faces[p, a]])
        out.append(fl)
    return out


def get_faces_from_pairs(edges1, edges2, threshold=15):
    """opencv_renewed.py:62-112: quads (e1, e2, e3, e4) with e1, e3 from edges1 and e2, e4 from edges2 whose
    neighbouring segments lie within ``threshold`` px of each other; corners by lineIntersection; overlapping
    duplicates removed; list of faces, each a list of four (x, y) tuples, oriented like the reference's."""
    return get_faces_from_pairs_batch([(edges1, edges2)], threshold)[0]


def get_camera_angles(image, iterations=1000, method="hough", refinement_iterations=500):
    """opencv_fit_color.py:71-93."""
    if method == "hough":
        eng, t = detectors._bgr_dev(image)
        eng.front_canny(t, 5, 10, blur=True, normalize=False)
        cap = 4096
        while True:
            ln, _, counts, _ = eng.hough_lines(None, 1, np.pi / 180, 50, max_lines=cap)
            n = int(counts[0].item())
            if n <= cap:
                break
            cap = n
        if n == 0:
            raise TypeError("'NoneType' object is not subscriptable")  # lines[:,0,:] on None in the reference
        lines = ln[0, :n].cpu().numpy()
    elif method == "lsd":
        pairs = detectors.lsd(image)
        pairs = [(a, b) for a, b in pairs if np.linalg.norm(b - a) > 10]
        lines = edges_to_polar_lines(pairs)
    else:
        return None
    return regress_lines(lines, iterations=iterations, refinement_iterations=refinement_iterations)[0]
