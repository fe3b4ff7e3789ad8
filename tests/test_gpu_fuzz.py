"""GPU differential fuzz: seeded random sizes / contents / parameters, engine vs oracle (bit-exact rows)."""
import numpy as np
import pytest
import torch

from oracle import restate as R

pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def random_image(rng, h, w):
    kind = rng.integers(0, 4)
    if kind == 0:
        return rng.integers(0, 256, (h, w), dtype=np.uint8)
    if kind == 1:  # piecewise-constant blocks (flat areas + sharp edges)
        img = np.full((h, w), int(rng.integers(0, 256)), np.uint8)
        for _ in range(rng.integers(1, 8)):
            y0, x0 = int(rng.integers(0, h)), int(rng.integers(0, w))
            img[y0:y0 + int(rng.integers(1, h + 1)), x0:x0 + int(rng.integers(1, w + 1))] = int(rng.integers(0, 256))
        return img
    if kind == 2:  # smooth ramps (long weak chains)
        yy, xx = np.mgrid[0:h, 0:w]
        return np.clip(128 + 60 * np.sin(xx / rng.uniform(3, 40)) + 60 * np.cos(yy / rng.uniform(3, 40)), 0, 255).astype(np.uint8)
    img = np.zeros((h, w), np.uint8)  # sparse lines
    for _ in range(rng.integers(1, 6)):
        img[int(rng.integers(0, h)), :] = 255
        img[:, int(rng.integers(0, w))] = int(rng.integers(1, 256))
    return img


@pytest.mark.parametrize("seed", range(24))
def test_fuzz_canny_and_hough(engine, seed):
    rng = np.random.default_rng(1000 + seed)
    h = int(rng.integers(1, 200)) if seed % 3 else int(rng.integers(1, 50)) * 4
    w = int(rng.integers(1, 300)) if seed % 2 else int(rng.integers(1, 75)) * 4
    g = random_image(rng, h, w)
    lo, hi = sorted(rng.integers(0, 300, 2).tolist())
    blur = bool(rng.integers(0, 2))
    e = engine.front_canny(dev(g)[None], lo, hi, blur=blur)[0].cpu().numpy()
    ref = R.canny(R.gauss5(g) if blur else g, lo, hi)
    assert np.array_equal(e, ref), (seed, h, w, lo, hi, blur)
    thr = int(rng.integers(1, 40))
    lines, votes, counts, accum = engine.hough_lines(None, 1, np.pi / 180, thr, max_lines=1 << 16, want_accum=True)
    rl, rv, ra = R.hough_lines(ref, 1, np.pi / 180, thr, return_accum=True)
    assert np.array_equal(accum[0].cpu().numpy(), ra)
    n = int(counts[0])
    assert n == (0 if rl is None else len(rl))
    if n:
        assert np.array_equal(lines[0, :n].cpu().numpy(), rl[:, 0]) and np.array_equal(votes[0, :n].cpu().numpy(), rv)
    engine.check()


@pytest.mark.parametrize("seed", range(16))
def test_fuzz_hough_lines_p(engine, seed):
    rng = np.random.default_rng(2000 + seed)
    h, w = int(rng.integers(8, 160)), int(rng.integers(8, 240))
    g = random_image(rng, h, w)
    e = R.canny(g, 20, 60)
    if seed % 4 == 0:
        e = (rng.random((h, w)) < 0.15).astype(np.uint8) * 255
    thr, mlen, gap = int(rng.integers(2, 40)), int(rng.integers(0, 40)), int(rng.integers(0, 12))
    segs, counts = engine.hough_lines_p(dev(e)[None], 1, np.pi / 180, thr, mlen, gap, max_segs=8192)
    ref = R.hough_lines_p(e, 1, np.pi / 180, thr, mlen, gap)
    n = int(counts[0])
    assert n == (0 if ref is None else len(ref)), (seed, h, w, thr, mlen, gap)
    if n:
        assert np.array_equal(segs[0, :n].cpu().numpy(), ref[:, 0])


@pytest.mark.parametrize("seed", range(10))
def test_fuzz_lsd(engine, seed):
    rng = np.random.default_rng(3000 + seed)
    h, w = int(rng.integers(20, 200)), int(rng.integers(20, 300))
    g = random_image(rng, h, w)
    refine = int(rng.integers(0, 2))
    scale = [0.8, 0.5, 1.0][seed % 3]
    segs, width, prec, counts = engine.lsd(dev(g)[None], refine=refine, scale=scale)
    ref, rw, rp = R.lsd(g, refine, scale)
    n = int(counts[0])
    assert n == (0 if ref is None else len(ref)), (seed, h, w, refine, scale)
    if n:
        assert np.abs(segs[0, :n].cpu().numpy() - ref[:, 0]).max() <= 0.5
        assert np.allclose(width[0, :n].cpu().numpy(), rw[:, 0], rtol=1e-6, atol=1e-6)
