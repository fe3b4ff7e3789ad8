"""GPU: one engine context driven through every chain back to back on small and odd-sized frames (scratch buffers
are shared between the chains: counters, queues, the LSD planes), each result checked against its oracle."""
import numpy as np
import pytest
import torch

from oracle import geometry as G
from oracle import restate as R

pytestmark = pytest.mark.gpu


def test_all_chains_share_one_context():
    import graphics_project_b200 as gp
    eng = gp.Engine(0)
    fr = np.stack([gp.synth.make_frame(s, 135, 240, k) for s, k in ((0, "cubes"), (1, "cave"), (2, "voxel"))])
    t = torch.from_numpy(fr).cuda()
    for rep in range(2):  # second pass: every scratch buffer is being reused
        e = eng.front_canny(t, 5, 10, blur=True)
        lines, votes, counts, acc = eng.hough_lines(None, 1, np.pi / 180, 30, want_accum=True)
        for i in range(3):
            ref = R.canny(R.gauss5(R.bgr2gray(fr[i])), 5, 10)
            assert np.array_equal(e[i].cpu().numpy(), ref)
            ln, v, a = R.hough_lines(ref, 1, np.pi / 180, 30, return_accum=True)
            assert int(counts[i]) == (0 if ln is None else len(ln)) and np.array_equal(acc[i].cpu().numpy(), a)
        e2 = eng.front_canny(t, 30, 50, blur=True, normalize=True)
        assert np.array_equal(e2[1].cpu().numpy(), R.canny(R.normalize_minmax(R.gauss5(R.bgr2gray(fr[1]))), 30, 50))
        view = np.ascontiguousarray(fr[0][::-1, :, ::-1].astype(np.float32) / np.float32(255))[::-1, :, ::-1]
        e32 = eng.f32_front_canny(eng.upload_f32_view(view), 30, 50)
        assert np.array_equal(e32[0].cpu().numpy(), R.f32_do_canny(view))
        segp, pc = eng.hough_lines_p(e, 1, np.pi / 180, 30, 20, 3)
        sp = R.hough_lines_p(e[2].cpu().numpy(), 1, np.pi / 180, 30, 20, 3)
        assert int(pc[2]) == (0 if sp is None else len(sp))
        gray = eng.bgr2gray(t)
        for refine, scale in ((0, .8), (1, .8), (2, .5), (0, 1.0)):
            out = eng.lsd(gray, refine=refine, scale=scale, want_nfa=True)
            for i in range(3):
                ref = R.lsd(gray[i].cpu().numpy(), refine, scale)
                n = int(out[3][i])
                assert n == (0 if ref[0] is None else len(ref[0])), (refine, scale, i)
                if n:
                    assert np.abs(out[0][i, :n].cpu().numpy() - ref[0][:, 0]).max() <= 0.5
        odd = eng.bgr2gray(torch.from_numpy(np.stack([gp.synth.make_frame(5, 97, 131, "cave")])).cuda())
        o = eng.lsd(odd, refine=1)  # byte-wise blur / one-pixel gradient kernels (width not a multiple of 4)
        ro = R.lsd(odd[0].cpu().numpy(), 1)
        assert int(o[3][0]) == (0 if ro[0] is None else len(ro[0]))
        sg = out[0][1, :int(out[3][1])].cpu().numpy().astype(np.float64)
        hz = sg[np.abs(sg[:, 2] - sg[:, 0]) >= np.abs(sg[:, 3] - sg[:, 1])]
        vt = sg[np.abs(sg[:, 2] - sg[:, 0]) < np.abs(sg[:, 3] - sg[:, 1])]
        if len(hz) and len(vt):
            q, fc, kp, qc = eng.face_quads(torch.from_numpy(hz)[None], torch.from_numpy(vt)[None], None, None, 15)
            rf, rq, rk = G.get_faces_from_pairs(hz, vt, 15)
            assert int(qc[0]) == len(rq) and np.array_equal(kp[0, :len(rq)].cpu().numpy().astype(bool), rk)
        if len(sg) >= 2:
            mo, ms, mc = eng.combine_parallel_lines(torch.from_numpy(sg)[None])
            assert int(mc[0]) == len(G.combine_parallel_lines(sg)[0])
        eng.check()
    eng.close()
