"""GPU parity: HoughLines (a5, bit-exact incl. order, votes and full accumulator) and HoughLinesP
(a6, bit-exact incl. emission order under cv::RNG(-1))."""
import numpy as np
import pytest
import torch

from oracle import restate as R

pytestmark = pytest.mark.gpu
IMAGES = ["demo_scarce", "sc_white_2", "sc_minecraft_cave", "old_sc_inf", "sc_cube", "sc_scarce_3_flat"]


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def run_hough(engine, edges_np, thr, rho=1.0, theta=np.pi / 180, cap=1 << 16):
    lines, votes, counts, accum = engine.hough_lines(dev(edges_np)[None], rho, theta, thr, max_lines=cap,
                                                     want_accum=True)
    n = int(counts[0])
    return lines[0, :n].cpu().numpy(), votes[0, :n].cpu().numpy(), accum[0].cpu().numpy()


@pytest.mark.parametrize("name", IMAGES)
def test_hough_lines_golden(engine, name, golden_outputs):
    o = lambda k: golden_outputs[f"{name}/{k}"]  # noqa: E731
    for ekey, thr, key in (("canny_blur", 50, "hough50"), ("docanny", 70, "hough70")):
        lines, votes, accum = run_hough(engine, o(ekey), thr)
        ref = o(key)
        assert lines.shape[0] == ref.shape[0]
        assert np.array_equal(lines, ref[:, 0, :2])
        assert np.array_equal(votes, ref[:, 0, 2].astype(np.int32))
        assert np.array_equal(accum, R.hough_lines(o(ekey), 1, np.pi / 180, thr, return_accum=True)[2])


@pytest.mark.parametrize("shape,rho,theta", [((1, 1), 1, np.pi / 180), ((9, 200), 1, np.pi / 180),
                                             ((203, 317), 1, np.pi / 180), ((120, 90), 2, np.pi / 90),
                                             ((64, 64), 0.5, np.pi / 360), ((713, 1277), 1, np.pi / 180)])
def test_hough_lines_vs_oracle_ragged(engine, shape, rho, theta):
    rng = np.random.default_rng(shape[0] * 7 + shape[1])
    h, w = shape
    edges = (rng.random((h, w)) < 0.03).astype(np.uint8) * 255
    if h > 20 and w > 20:
        edges[h // 2, :] = 255
        edges[:, w // 3] = 255
        np.fill_diagonal(edges, 255)
    for thr in (5, 30):
        lines, votes, accum = run_hough(engine, edges, thr, rho, theta)
        rl, rv, ra = R.hough_lines(edges, rho, theta, thr, return_accum=True)
        assert np.array_equal(accum, ra)
        assert accum.sum() == (accum.shape[0] - 2) * int((edges > 0).sum())
        if rl is None:
            assert len(lines) == 0
        else:
            assert np.array_equal(lines, rl[:, 0]) and np.array_equal(votes, rv)


def test_hough_blank_and_truncation(engine):
    z = torch.zeros((2, 50, 70), dtype=torch.uint8, device="cuda")
    lines, votes, counts, accum = engine.hough_lines(z, 1, np.pi / 180, 10, want_accum=True)
    assert counts.tolist() == [0, 0] and not accum.any()
    # truncation: counts holds the true number, the first max_lines entries are the top ones
    rng = np.random.default_rng(3)
    e = (rng.random((200, 300)) < 0.05).astype(np.uint8)
    full = run_hough(engine, e, 8)[0]
    lines, _, counts, _ = engine.hough_lines(dev(e)[None], 1, np.pi / 180, 8, max_lines=16)
    assert int(counts[0]) == len(full) > 16
    assert np.array_equal(lines[0].cpu().numpy(), full[:16])


def test_fused_canny_hough_reuses_edge_lists(engine, golden_inputs, golden_outputs):
    """edges=None votes from the queue the Canny pass left behind (no re-scan of the map)."""
    names = ["demo_scarce", "sc_white_2", "sc_cube"]
    batch = np.stack([golden_inputs[n] for n in names])
    engine.front_canny(dev(batch), 5, 10, blur=True)
    lines, votes, counts, accum = engine.hough_lines(None, 1, np.pi / 180, 50, want_accum=True)
    for i, n in enumerate(names):
        ref = golden_outputs[f"{n}/hough50"]
        k = int(counts[i])
        assert k == len(ref)
        assert np.array_equal(lines[i, :k].cpu().numpy(), ref[:, 0, :2])
        assert np.array_equal(votes[i, :k].cpu().numpy(), ref[:, 0, 2].astype(np.int32))
    engine.check()


def test_hough_full_size_properties(engine):
    """1080p / 4K: accumulator checksum == 180 * nnz(edges), and equality with the oracle."""
    import graphics_project_b200.synth as synth
    for h, w in ((1080, 1920), (2160, 3840)):
        img = synth.make_frame(4, h, w)
        edges = engine.front_canny(dev(img)[None], 5, 10, blur=True)
        lines, votes, counts, accum = engine.hough_lines(None, 1, np.pi / 180, 50, want_accum=True, max_lines=8192)
        nnz = int((edges > 0).sum())
        assert int(accum.sum()) == 180 * nnz
        rl, rv, ra = R.hough_lines(edges[0].cpu().numpy(), 1, np.pi / 180, 50, return_accum=True)
        assert np.array_equal(accum[0].cpu().numpy(), ra)
        k = int(counts[0])
        assert k == len(rl) and np.array_equal(lines[0, :k].cpu().numpy(), rl[:, 0])
        assert np.array_equal(votes[0, :k].cpu().numpy(), rv)


@pytest.mark.parametrize("name", IMAGES)
def test_hough_lines_p_golden(engine, name, golden_outputs):
    o = lambda k: golden_outputs[f"{name}/{k}"]  # noqa: E731
    for ekey, params, key in (("canny_prob", (30, 50, 10), "houghp"), ("docanny", (30, 20, 3), "houghp2")):
        segs, counts = engine.hough_lines_p(dev(o(ekey))[None], 1, np.pi / 180, *params, max_segs=2048)
        n = int(counts[0])
        ref = o(key)
        assert n == len(ref)
        assert np.array_equal(segs[0, :n].cpu().numpy(), ref[:, 0])


def test_hough_lines_p_vs_oracle_batch_and_ragged(engine):
    import graphics_project_b200.synth as synth
    frames = [synth.make_frame(s, 405, 720) for s in range(6)]
    edges = np.stack([R.canny(R.bgr2gray(f), 5, 150) for f in frames])
    edges[5] = 0  # a blank frame inside the batch
    segs, counts = engine.hough_lines_p(dev(edges), 1, np.pi / 180, 30, 50, 10)
    for i in range(6):
        ref = R.hough_lines_p(edges[i], 1, np.pi / 180, 30, 50, 10)
        n = int(counts[i])
        assert n == (0 if ref is None else len(ref))
        if n:
            assert np.array_equal(segs[i, :n].cpu().numpy(), ref[:, 0])
    rng = np.random.default_rng(1)
    for h, w in ((1, 1), (7, 300), (211, 97)):
        e = (rng.random((h, w)) < 0.2).astype(np.uint8) * 255
        segs, counts = engine.hough_lines_p(dev(e)[None], 1, np.pi / 180, 10, 5, 2)
        ref = R.hough_lines_p(e, 1, np.pi / 180, 10, 5, 2)
        n = int(counts[0])
        assert n == (0 if ref is None else len(ref))
        if n:
            assert np.array_equal(segs[0, :n].cpu().numpy(), ref[:, 0])


def test_hough_lines_p_full_size(engine):
    import graphics_project_b200.synth as synth
    for h, w in ((1080, 1920), (2160, 3840)):
        img = synth.make_frame(9, h, w)
        edges = engine.front_canny(dev(img)[None], 5, 150)
        segs, counts = engine.hough_lines_p(edges, 1, np.pi / 180, 30, 50, 10)
        ref = R.hough_lines_p(edges[0].cpu().numpy(), 1, np.pi / 180, 30, 50, 10)
        n = int(counts[0])
        assert n == len(ref) and np.array_equal(segs[0, :n].cpu().numpy(), ref[:, 0])
