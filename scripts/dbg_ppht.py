import sys, ctypes as C; sys.path.insert(0, "/root/repo")
import numpy as np, torch
import graphics_project_b200 as gp
from graphics_project_b200._native import lib
from oracle import restate as R
eng = gp.Engine(0)
for (h, w, seed) in ((270, 480, 0), (270, 480, 1)):
    img = gp.synth.make_frame(seed, h, w)
    g = R.bgr2gray(img); e = R.canny(R.gauss5(g), 5, 10)
    cap = 5000
    tr_ref = np.zeros((cap, 10), np.int32)
    R.lib().or_ppht_trace(tr_ref.ctypes.data_as(C.c_void_p), cap)
    ref = R.hough_lines_p(e, 1, np.pi/180, 30, 50, 10)
    nref = R.lib().or_ppht_trace_count(); R.lib().or_ppht_trace(None, 0)
    tr = torch.zeros((cap, 10), dtype=torch.int32, device="cuda")
    lib.le_debug_ppht_trace.argtypes = [C.c_void_p, C.c_int]
    lib.le_debug_ppht_trace(C.c_void_p(tr.data_ptr()), cap)
    segs, counts = eng.hough_lines_p(torch.from_numpy(e).cuda()[None], 1, np.pi/180, 30, 50, 10)
    torch.cuda.synchronize(); lib.le_debug_ppht_trace(None, 0)
    tr = tr.cpu().numpy()
    print("ref trace", nref)
    for i in range(nref):
        if not np.array_equal(tr[i], tr_ref[i]):
            print("first trace diff at", i); print(" gpu", tr[max(0,i-1):i+3].tolist()); print(" ref", tr_ref[max(0,i-1):i+3].tolist()); break
    else:
        print("traces equal")
