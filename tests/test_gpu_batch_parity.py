"""GPU parity at the BASELINE batch shapes: whole batches through the batched API, every
frame (or a strided sample of the big ones) against the CPU oracle, plus the size-independent properties a batch
offers (repeated frames give repeated results, the accumulator sums to angles x edge pixels)."""
import numpy as np
import pytest
import torch

import graphics_project_b200 as gp
from oracle import restate as R

pytestmark = pytest.mark.gpu


def _batch(n, h, w, kind, unique):
    u = gp.synth.make_batch(unique, h, w, kind)
    return np.stack([u[i % unique] for i in range(n)]), u


def test_canny_hough_batch_256x1080p(engine):
    """BASELINE config 2 as bench.py runs it: 256 x 1080p, accumulator written; 8 distinct frames vs the oracle
    (edges, accumulator, lines, order), every other frame against its twin."""
    frames, u = _batch(256, 1080, 1920, "cubes", 8)
    t = torch.from_numpy(frames).cuda()
    edges = engine.front_canny(t, 5, 10, blur=True)
    lines, votes, counts, accum = engine.hough_lines(None, 1.0, np.pi / 180, 50, max_lines=1024, want_accum=True)
    engine.check()
    c = counts.cpu().numpy()
    for i in range(8):
        e = R.canny(R.gauss5(R.bgr2gray(u[i])), 5, 10)
        assert np.array_equal(edges[i].cpu().numpy(), e)
        ln, v, acc = R.hough_lines(e, 1, np.pi / 180, 50, return_accum=True)
        assert np.array_equal(accum[i].cpu().numpy(), acc)
        n = 0 if ln is None else len(ln)
        assert c[i] == n
        k = min(n, 1024)
        assert np.array_equal(lines[i, :k].cpu().numpy(), ln[:k, 0]) and np.array_equal(votes[i, :k].cpu().numpy(), v[:k])
    nedge = (edges.view(256, -1) != 0).sum(1)
    assert torch.equal(accum.view(256, -1).sum(1), nedge * 180)  # every edge pixel votes once per angle
    for i in range(8, 256):
        assert c[i] == c[i % 8]
    assert torch.equal(edges[8:16], edges[:8]) and torch.equal(edges[248:], edges[:8])
    assert torch.equal(lines[248:, :64], lines[:8, :64]) and torch.equal(accum[200:208], accum[:8])


def test_houghp_batch_4k(engine):
    """BASELINE config 3 shape (4K, prob() chain), 24 frames of 6 distinct ones: every distinct frame against the
    oracle in cv2's order, twins identical."""
    frames, u = _batch(24, 2160, 3840, "cubes", 6)
    t = torch.from_numpy(frames).cuda()
    edges = engine.front_canny(t, 5, 150)
    segs, counts = engine.hough_lines_p(edges, 1, np.pi / 180, 30, 50, 10, max_segs=1024)
    engine.check()
    c = counts.cpu().numpy()
    for i in range(6):
        e = R.canny(R.bgr2gray(u[i]), 5, 150)
        assert np.array_equal(edges[i].cpu().numpy(), e)
        sp = R.hough_lines_p(e, 1, np.pi / 180, 30, 50, 10)
        n = 0 if sp is None else len(sp)
        assert c[i] == n and n > 20
        assert np.array_equal(segs[i, :n].cpu().numpy(), sp[:, 0])
    for i in range(6, 24):
        assert c[i] == c[i % 6] and torch.equal(segs[i, :c[i]], segs[i % 6, :c[i]])


def test_houghp_batch_1080p_dense_and_sparse(engine):
    """More frames than SMs, mixed densities (cube, cave: ~10x the edge pixels, blank): exact PPHT per frame."""
    kinds = ["cubes", "cave", "voxel"]
    u = [gp.synth.make_frame(s, 540, 960, kinds[s % 3]) for s in range(9)]
    u.append(np.full((540, 960, 3), 77, np.uint8))
    frames = np.stack([u[i % 10] for i in range(160)])
    t = torch.from_numpy(frames).cuda()
    edges = engine.front_canny(t, 5, 150)
    segs, counts = engine.hough_lines_p(edges, 1, np.pi / 180, 30, 50, 10, max_segs=2048)
    engine.check()
    c = counts.cpu().numpy()
    for i in range(10):
        e = R.canny(R.bgr2gray(u[i]), 5, 150)
        sp = R.hough_lines_p(e, 1, np.pi / 180, 30, 50, 10)
        n = 0 if sp is None else len(sp)
        assert c[i] == n
        if n:
            assert np.array_equal(segs[i, :n].cpu().numpy(), sp[:, 0])
    assert c[9] == 0
    for i in range(10, 160):
        assert c[i] == c[i % 10] and torch.equal(segs[i, :c[i]], segs[i % 10, :c[i]])


@pytest.mark.parametrize("refine", [0, 1, 2])
def test_lsd_batch_of_200(engine, refine):
    """More frames than SMs (the 8-warp instantiation, several CTAs per SM), three scene kinds: 40 distinct frames
    against the oracle, the four copies of each identical (the speculative grower is deterministic)."""
    u = np.stack([gp.synth.make_frame(s, 270, 480, ("cubes", "voxel", "cave")[s % 3]) for s in range(40)])
    gray = engine.bgr2gray(torch.from_numpy(np.concatenate([u] * 5)).cuda())
    segs, width, prec, counts, nfa = engine.lsd(gray, refine=refine, want_nfa=True)
    engine.check()
    c = counts.cpu().numpy()
    for i in range(40):
        r = R.lsd(gray[i].cpu().numpy(), refine, return_nfa=True)
        nr = 0 if r[0] is None else len(r[0])
        assert c[i] == nr
        if nr:
            assert np.abs(segs[i, :nr].cpu().numpy() - r[0][:, 0]).max() <= 0.5
        for rep in range(1, 5):
            j = i + 40 * rep
            assert c[j] == nr and torch.equal(segs[j, :nr], segs[i, :nr])


def test_lsd_batch_1080p_voxel(engine):
    """BASELINE config 4 shape: 48 x 1080p voxel frames (6 distinct) within 0.5 px of the oracle, twins identical."""
    frames, u = _batch(48, 1080, 1920, "voxel", 6)
    gray = engine.bgr2gray(torch.from_numpy(frames).cuda())
    segs, _, _, counts = engine.lsd(gray, max_segs=4096)
    engine.check()
    c = counts.cpu().numpy()
    for i in range(6):
        r = R.lsd(R.bgr2gray(u[i]))[0]
        n = 0 if r is None else len(r)
        assert c[i] == n and n > 50
        assert np.abs(segs[i, :n].cpu().numpy() - r[:, 0]).max() <= 0.5
    for i in range(6, 48):
        assert c[i] == c[i % 6] and torch.equal(segs[i, :c[i]], segs[i % 6, :c[i]])


def test_two_devices_in_one_process():
    """cudaFuncSetAttribute is per device -- a second engine on another GPU of the same process must
    launch every opt-in-shared-memory kernel, and the calls must leave the caller's current device alone."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    frames = torch.from_numpy(np.stack([gp.synth.make_frame(s, 270, 480) for s in range(2)]))
    res = []
    torch.cuda.set_device(0)
    for d in (0, 1):
        eng = gp.Engine(d)
        t = frames.to(f"cuda:{d}")
        with torch.cuda.device(d):
            edges = eng.front_canny(t, 5, 10, blur=True)
            lines, _, counts, accum = eng.hough_lines(None, 1, np.pi / 180, 50, want_accum=True)
            segs, pc = eng.hough_lines_p(edges, 1, np.pi / 180, 30, 50, 10)
            ls, _, _, lc = eng.lsd(eng.bgr2gray(t))
            eng.check()
        res.append([x.cpu() for x in (edges, lines[:, :32], counts, segs[:, :16], pc, ls[:, :16], lc)])
    assert torch.cuda.current_device() == 0
    for a, b in zip(*res):
        assert torch.equal(a, b)


def test_device_is_restored_after_calls(engine):
    """The ABI runs on its context's device and restores the caller's current device."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    t = torch.from_numpy(gp.synth.make_frame(0, 120, 160))[None].cuda(0)
    torch.cuda.set_device(1)
    try:
        e = engine.front_canny(t, 5, 10)
        assert torch.cuda.current_device() == 1 and e.device.index == 0
    finally:
        torch.cuda.set_device(0)


def test_fresh_contexts_on_concurrent_streams(engine):
    """The first call of a freshly created context allocates and clears its scratch; three contexts doing that at
    once on their own (non-blocking) streams must not see a buffer being cleared under a running kernel
    (le_ensure used to clear on the legacy default stream without waiting: intermittent illegal address in the
    multi-stream host pipelines)."""
    u = np.stack([gp.synth.make_frame(s, 540, 960, "voxel") for s in range(6)])
    t = torch.from_numpy(np.concatenate([u] * 4)).cuda()  # 24 frames
    gray = engine.bgr2gray(t)
    want_s, _, _, want_c = engine.lsd(gray, max_segs=1024)
    edges = engine.front_canny(t, 5, 150)
    want_p, want_pc = engine.hough_lines_p(edges, 1, np.pi / 180, 30, 50, 10, max_segs=512)
    torch.cuda.synchronize()
    for rep in range(3):
        engs = [gp.Engine(0) for _ in range(3)]
        streams = [torch.cuda.Stream() for _ in range(3)]
        outs = []
        for e, s in zip(engs, streams):
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                g = e.bgr2gray(t)
                segs, _, _, c = e.lsd(g, max_segs=1024)
                ed = e.front_canny(t, 5, 150)
                p, pc = e.hough_lines_p(ed, 1, np.pi / 180, 30, 50, 10, max_segs=512)
                outs.append((segs, c, p, pc))
        torch.cuda.synchronize()
        for e in engs:
            e.check()
        for segs, c, p, pc in outs:
            assert torch.equal(c, want_c) and torch.equal(pc, want_pc)
            for i in range(24):
                assert torch.equal(segs[i, :int(c[i])], want_s[i, :int(c[i])])
                assert torch.equal(p[i, :int(pc[i])], want_p[i, :int(pc[i])])
        for e in engs:
            e.close()


def test_houghp_dense_frames_use_the_global_order_path(engine):
    """Noise frames: far more edge pixels than the shared-memory list of k_ppht_order holds (the shuffle then runs
    in place in global memory), thousands of triggers, almost no accepted segment -- still cv2's exact result."""
    rng = np.random.default_rng(11)
    frames = rng.integers(0, 256, (3, 270, 480), dtype=np.uint8)
    frames[2, 100:170] = 90  # a flat band: long runs without edge pixels
    t = torch.from_numpy(frames).cuda()
    edges = engine.canny(t, 40, 120)
    assert int((edges[0] != 0).sum()) > 20000
    segs, counts = engine.hough_lines_p(edges, 1, np.pi / 180, 30, 20, 3, max_segs=8192)
    engine.check()
    for i in range(3):
        e = edges[i].cpu().numpy()
        assert np.array_equal(e, R.canny(frames[i], 40, 120))
        sp = R.hough_lines_p(e, 1, np.pi / 180, 30, 20, 3)
        n = 0 if sp is None else len(sp)
        assert int(counts[i]) == n
        if n:
            assert np.array_equal(segs[i, :n].cpu().numpy(), sp[:, 0])


def test_refused_parameters_and_reported_overflow(engine):
    """What the engine cannot do exactly it refuses (never a silent wrong answer): rho beyond the 16-bit counters,
    the order-free HoughLinesP mode, ambiguous / malformed tensors; a peak list that overflows is reported by
    check()."""
    z = torch.zeros((1, 1080, 1920), dtype=torch.uint8, device="cuda")
    with pytest.raises(gp.EngineError) as ei:
        engine.hough_lines(z, 40.0, np.pi / 180, 50)
    assert ei.value.code == -5
    with pytest.raises(gp.EngineError) as ei:
        engine.hough_lines_p(z, 40.0, np.pi / 180, 30, 50, 10)
    assert ei.value.code == -5
    with pytest.raises(gp.EngineError) as ei:
        engine.hough_lines_p(z, 1, np.pi / 180, 30, 50, 10, exact_order=False)
    assert ei.value.code == -5 and "order" in str(ei.value)
    with pytest.raises(ValueError):
        engine.front_canny(torch.zeros((64, 64, 3), dtype=torch.uint8, device="cuda"), 5, 10)
    with pytest.raises(ValueError):
        engine.front_canny(z, 5, 10, out=torch.zeros((1, 1080, 1919), dtype=torch.uint8, device="cuda"))
    with pytest.raises(TypeError):
        engine.front_canny(z, 5, 10, out=torch.zeros((1, 1080, 1920), dtype=torch.int32, device="cuda"))
    with pytest.raises(ValueError):
        engine.hough_lines(z, 1, np.pi / 180, 50, out=(torch.zeros((1, 64, 2), device="cuda"), None,
                                                      torch.zeros((2,), dtype=torch.int32, device="cuda"), None))
    # more local maxima than the peak list holds: counts still says how many, check() says the lines are unreliable
    rng = np.random.default_rng(5)
    noise = torch.from_numpy((rng.random((1, 2160, 3840)) < 0.3).astype(np.uint8) * 255).cuda()
    lines, _, counts, _ = engine.hough_lines(noise, 1, np.pi / 180, 2, max_lines=256)
    assert int(counts[0]) > 65536
    with pytest.raises(gp.EngineError) as ei:
        engine.check()
    assert ei.value.code == -4
    # the flag is per call: the next call on the same context is clean again
    engine.hough_lines(z, 1, np.pi / 180, 50)
    engine.check()
