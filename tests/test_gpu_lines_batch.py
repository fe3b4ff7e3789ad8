"""GPU parity of the batched `lines` pipeline (BASELINE config 5; SURVEY 8a a14 'lsd' branch): le_polar_batch
and le_vp_fit_batch against the single-frame path (vp.edges_to_polar_lines, vp.regress_lines -- themselves checked
against the reference's outputs in test_gpu_lsd_vp.py) and against the reference's own loss and seeded runs."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def ref_sum_loss(phi, theta, lines, width=600, height=400):
    """numpy restatement of sum_loss / get_focal_points (opencv_fit_color.py:14-26, 95-101)."""
    from graphics_project_b200 import vp
    fp = vp.get_focal_points(phi, theta, width, height)
    return np.mean([min((p[1] * np.sin(ph) + p[0] * np.cos(ph) - rho) ** 2 for p in fp) for rho, ph in lines])


def pad_lines(sets):
    mx = max(1, max(len(s) for s in sets))
    polar = np.zeros((len(sets), mx, 2))
    for i, s in enumerate(sets):
        polar[i, :len(s)] = s
    return torch.from_numpy(polar).cuda(), torch.tensor([len(s) for s in sets], dtype=torch.int32).cuda()


@pytest.fixture(scope="module")
def lsd_batch(engine):
    from graphics_project_b200 import synth
    frames = [synth.make_frame(s, 270, 480, kind=k) for s, k in ((0, "cubes"), (1, "voxel"), (2, "cave"), (3, "cubes"))]
    frames.insert(2, np.full((270, 480, 3), 128, np.uint8))  # a frame without a single segment
    t = torch.from_numpy(np.stack(frames)).cuda()
    segs, _, _, counts = engine.lsd(engine.bgr2gray(t), max_segs=1024)
    return t, segs, counts


def test_polar_batch_matches_single_frame_path(engine, lsd_batch):
    from graphics_project_b200 import vp
    _, segs, counts = lsd_batch
    polar, pcounts = engine.polar_batch(segs, counts, 10.0)
    segs_h, counts_h, polar_h, pcounts_h = segs.cpu().numpy(), counts.cpu().numpy(), polar.cpu().numpy(), pcounts.cpu().numpy()
    assert counts_h[2] == 0 and pcounts_h[2] == 0
    for f in range(len(counts_h)):
        pairs = [(p[:2], p[2:]) for p in segs_h[f, :counts_h[f]]]
        pairs = [(a, b) for a, b in pairs if np.linalg.norm(b - a) > 10]  # opencv_fit_color.py:88
        assert pcounts_h[f] == len(pairs)
        if pairs:
            want = vp.edges_to_polar_lines(pairs)
            assert np.allclose(polar_h[f, :len(pairs)], want, rtol=1e-12, atol=1e-12)
    assert (pcounts_h[[0, 1, 3, 4]] > 0).all() and (pcounts_h <= counts_h).all() and (pcounts_h < counts_h).any()


def test_vp_fit_batch_equals_stepwise_search(engine, golden_vp, lsd_batch):
    """Without seeds the fused search visits exactly the candidates regress_lines visits step by step."""
    from graphics_project_b200 import vp
    _, segs, counts = lsd_batch
    polar, pcounts = engine.polar_batch(segs, counts, 10.0)
    sets = [golden_vp["polar_lsd"], golden_vp["polar_lsd"][:7], np.zeros((0, 2))]
    sets += [polar[f, :int(pcounts[f])].cpu().numpy() for f in (0, 1, 3)]
    pl, pc = pad_lines(sets)
    fit, hist = vp.regress_lines_batch(pl, pc, 1000, 500, np.deg2rad(15), want_hist=False, engine=engine)
    assert hist is None
    fit = fit.cpu().numpy()
    assert tuple(fit[2]) == (0.0, 0.0, 100000.0)
    for i, lines in enumerate(sets):
        if len(lines) == 0:
            continue
        (phi, theta), loss = vp.regress_lines_stepwise(lines, 1000, 500, np.deg2rad(15), verbose=False, use_seeds=False)
        assert fit[i, 2] == pytest.approx(loss, rel=1e-9), i
        assert fit[i, 0] == pytest.approx(phi, abs=1e-9) and fit[i, 1] == pytest.approx(theta, abs=1e-9), i
        assert ref_sum_loss(fit[i, 0], fit[i, 1], lines) == pytest.approx(fit[i, 2], rel=1e-9)


def test_vp_fit_batch_with_sphere_seeds(engine, golden_vp):
    """a12 criterion for the fused search: scored by the reference loss it must be <= the median best
    loss of the seeded reference runs; the per-frame histogram equals le_vp_sphere_hist up to boundary pairs."""
    from graphics_project_b200 import vp
    lines = golden_vp["polar_lsd"]
    pl, pc = pad_lines([lines, lines[::2]])
    fit, hist = vp.regress_lines_batch(pl, pc, 1000, 500, np.deg2rad(15), engine=engine)
    fit, hist = fit.cpu().numpy(), hist.cpu().numpy()
    assert hist.shape == (2, 90, 360)
    runs = golden_vp["regress_runs_lsd"]
    assert ref_sum_loss(fit[0, 0], fit[0, 1], lines) == pytest.approx(fit[0, 2], rel=1e-9)
    assert fit[0, 2] <= np.median(runs[:, 2])
    _, loss = vp.regress_lines_stepwise(lines, 1000, 500, np.deg2rad(15), verbose=False)
    assert fit[0, 2] <= loss * 1.02
    # the single-frame wrapper IS the batch kernel with n = 1: identical result
    (phi1, th1), loss1 = vp.regress_lines(lines, 1000, 500, np.deg2rad(15), verbose=False)
    assert (phi1, th1, loss1) == tuple(fit[0])
    for i, ls in enumerate((lines, lines[::2])):
        m = len(ls)
        rho, phi = ls[:, 0], ls[:, 1]
        p0 = np.stack([rho * np.cos(phi), rho * np.sin(phi)], 1)
        d = np.stack([-np.sin(phi), np.cos(phi)], 1) * 100.0
        single = engine.vp_sphere_hist(torch.from_numpy(np.concatenate([p0 - d, p0 + d], 1)), 600, 400, 360, 90).cpu().numpy()
        assert abs(int(hist[i].sum()) - int(single.sum())) <= 2 and hist[i].sum() > 0.9 * m * (m - 1) / 2
        assert np.abs(hist[i] - single).sum() <= max(4, 0.002 * single.sum())


def test_get_camera_angles_batch(engine, lsd_batch):
    """Frames -> (phi, theta) on the device; the same lines through the single-frame search give the same loss."""
    from graphics_project_b200 import vp
    t, segs, counts = lsd_batch
    fit, hist, polar, pcounts, segs2, counts2 = vp.get_camera_angles_batch(t, 1000, 500, width=600, height=400,
                                                                           max_segs=1024, engine=engine)
    assert torch.equal(counts2, counts) and torch.equal(segs2[0, :int(counts[0])], segs[0, :int(counts[0])])
    fit = fit.cpu().numpy()
    assert tuple(fit[2]) == (0.0, 0.0, 100000.0) and int(hist[2].sum()) == 0
    for f in (0, 1, 3, 4):
        lines = polar[f, :int(pcounts[f])].cpu().numpy()
        assert 0 <= fit[f, 0] < np.pi + 0.3 and -0.3 <= fit[f, 1] < np.pi / 2 + 0.3
        assert ref_sum_loss(fit[f, 0], fit[f, 1], lines) == pytest.approx(fit[f, 2], rel=1e-9)
        _, loss = vp.regress_lines_stepwise(lines, 1000, 500, 0.3, verbose=False)
        assert fit[f, 2] <= loss * 1.02
