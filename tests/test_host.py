"""CPU: host-side logic -- sharding, result gather over gloo (world_size 2), synthetic data, VP scalars."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import graphics_project_b200 as gp
from graphics_project_b200 import dist as gdist, synth, vp, detectors


def test_shard_range_partitions():
    for total in (0, 1, 7, 256, 1024, 1025):
        for world in (1, 2, 3, 8):
            spans = [gdist.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gather_worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = gdist.shard_range(total, rank, world)
    counts = torch.arange(lo, hi, dtype=torch.int32) % 5
    payload = torch.zeros((hi - lo, 6, 4), dtype=torch.int32)
    for i in range(hi - lo):
        payload[i, :counts[i]] = lo + i
    c, p = gdist.gather_results(counts, payload, total)
    # pipelined form: two steps in flight with separate output slots, waited one step late
    h0 = gdist.gather_results(counts, payload, total, async_op=True, slot=0)
    h1 = gdist.gather_results(counts + 1, payload, total, async_op=True, slot=1)
    c0, p0 = h0.wait()
    c1, _ = h1.wait()
    assert torch.equal(c0, c) and torch.equal(p0, p) and torch.equal(c1, c + 1)
    q.put((rank, c.numpy().copy(), p.numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


def _gather_vp_worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = gdist.shard_range(total, rank, world)
    ids = torch.arange(lo, hi)
    pcounts = (ids % 4).to(torch.int32)
    fit = torch.stack([ids * 0.5, ids * 0.25, ids + 100.0], 1).to(torch.float64)
    hist = torch.zeros((hi - lo, 3, 5), dtype=torch.int32)
    hist[:, 1, 2] = ids.to(torch.int32) + 1
    c, f, h = gdist.gather_vp(pcounts, fit, hist, total)
    c2, f2, scene = gdist.gather_vp(pcounts, fit, hist, total, scene_hist=True)
    assert torch.equal(c, c2) and torch.equal(f, f2)
    q.put((rank, c.numpy().copy(), f.numpy().copy(), h.numpy().copy(), scene.numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [5, 6])
def test_gather_vp_gloo_world2(total):
    """(pcounts, fit, sphere histograms) of the lines pipeline: per-frame gather and the scene-level reduce."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gather_vp_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    ids = np.arange(total)
    for _, c, f, h, scene in got:
        assert np.array_equal(c, ids % 4)
        assert np.array_equal(f, np.stack([ids * 0.5, ids * 0.25, ids + 100.0], 1))
        assert h.shape == (total, 3, 5) and np.array_equal(h[:, 1, 2], ids + 1) and h.sum() == (ids + 1).sum()
        assert scene.shape == (3, 5) and scene[1, 2] == (ids + 1).sum() and scene.sum() == (ids + 1).sum()


@pytest.mark.parametrize("total", [7, 8])
def test_gather_results_gloo_world2(total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for _, c, p in got:
        assert c.shape == (total,) and p.shape == (total, 6, 4)
        assert np.array_equal(c, np.arange(total) % 5)
        for i in range(total):
            assert (p[i, :c[i]] == i).all() and (p[i, c[i]:] == 0).all()


def test_synth_deterministic_and_shaped():
    a = synth.make_frame(5, 135, 240)
    b = synth.make_frame(5, 135, 240)
    assert a.dtype == np.uint8 and a.shape == (135, 240, 3) and np.array_equal(a, b)
    assert not np.array_equal(a, synth.make_frame(6, 135, 240))
    batch = synth.make_batch(5, 54, 96, unique=2)
    assert batch.shape == (5, 54, 96, 3) and np.array_equal(batch[0], batch[2]) and np.array_equal(batch[1], batch[3])
    assert tuple(a[0, 0]) == synth.BACKGROUND_BGR


def test_vp_scalar_functions_match_reference_golden(golden_vp):
    for (p, t), fp in zip(golden_vp["cands"], golden_vp["focal_points"]):
        assert np.allclose(np.array(vp.get_focal_points(p, t)), fp, rtol=1e-12, atol=1e-9)
    # which_line ordering rules (x <= y <= z tie order, None at the threshold)
    fp = [np.array([0.0, 0.0]), np.array([0.0, 0.0]), np.array([5.0, 5.0])]
    assert vp.which_line(fp, (0.0, 0.3)) == "x"
    assert vp.which_line(fp, (1e3, 0.3), threshold=700) is None
    assert vp.loss_function((3.0, 4.0), 1.0, 0.0) == pytest.approx(4.0)


def test_line_matrix_to_pairs():
    m = np.arange(8, dtype=np.int32).reshape(2, 1, 4)
    pairs = detectors.lineMatrixToPairs(m)
    assert len(pairs) == 2 and np.array_equal(pairs[1][0], [4, 5]) and np.array_equal(pairs[1][1], [6, 7])
    with pytest.raises(TypeError):
        detectors.lineMatrixToPairs(None)  # reference behaviour: lines[0] on None


def test_shim_rejects_off_path_parameters_without_gpu():
    from graphics_project_b200 import cv_shim as cv
    img = np.zeros((8, 8), np.uint8)
    with pytest.raises(cv.error):
        cv.Canny(img.astype(np.float32), 5, 10)
    with pytest.raises(cv.error):
        cv.Canny(img, 5, 10, apertureSize=5)
    with pytest.raises(cv.error):
        cv.HoughLines(np.zeros((8, 8, 3), np.uint8), 1, np.pi / 180, 10)
    with pytest.raises(cv.error):
        cv.cvtColor(img, cv.COLOR_BGR2GRAY)
    with pytest.raises(cv.error):
        cv.GaussianBlur(img, (7, 7), 0)
    with pytest.raises(cv.error):
        cv.createLineSegmentDetector(3)
    assert cv.createLineSegmentDetector(2).params["refine"] == 2  # LSD_REFINE_ADV is on the path
    assert cv.COLOR_BGR2GRAY == 6 and cv.NORM_MINMAX == 32
    assert callable(cv.imread)  # non-hot-path attributes are forwarded to the real cv2


def test_shim_install_lets_reference_import_unchanged():
    """INTEGRATION.md route A: with the shim at sys.modules['cv2'] the reference's own modules import and
    bind the GPU entry points (skipped where the reference checkout is absent, e.g. on the GPU box)."""
    import subprocess
    import sys
    if not os.path.isdir("/root/reference"):
        pytest.skip("reference checkout not present")
    code = r'''
import sys, types, os
sys.dont_write_bytecode = True
sys.path.insert(0, %r)
import graphics_project_b200 as gp
shim = gp.cv_shim.install()
for name in ("moderngl", "matplotlib", "matplotlib.pyplot", "mpl_toolkits", "mpl_toolkits.mplot3d"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules["moderngl"].Framebuffer = object; sys.modules["moderngl"].Context = object
sys.modules["mpl_toolkits.mplot3d"].Axes3D = object
sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
os.chdir("/root/reference"); sys.path.insert(0, "/root/reference")
import opencv, opencv_fit_color, opencv_renewed
assert opencv.cv is shim and opencv_fit_color.cv is shim
assert opencv.cv.Canny is shim.Canny and opencv.cv.HoughLinesP is shim.HoughLinesP
assert opencv.cv.imread is not None and opencv.cv.COLOR_BGR2GRAY == 6
print("ok")
''' % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stderr[-2000:]
