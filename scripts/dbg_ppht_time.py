import sys, ctypes as C; sys.path.insert(0, "/root/repo")
import numpy as np, torch
import graphics_project_b200 as gp
from graphics_project_b200._native import lib
eng = gp.Engine(0)
frames = np.stack([gp.synth.make_frame(s, 1080, 1920) for s in range(4)])
t = torch.from_numpy(np.concatenate([frames]*64)).cuda()   # 256 frames, frame 0 traced
edges = eng.front_canny(t, 5, 150)
cap = 4096
tr = torch.zeros((cap, 10), dtype=torch.int32, device="cuda")
lib.le_debug_ppht_trace.argtypes = [C.c_void_p, C.c_int]
for n in (1, 256):
    tr.zero_()
    lib.le_debug_ppht_trace(C.c_void_p(tr.data_ptr()), cap)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    eng.hough_lines_p(edges[:n], 1, np.pi/180, 30, 50, 10); torch.cuda.synchronize()
    e0.record(); segs, counts = eng.hough_lines_p(edges[:n], 1, np.pi/180, 30, 50, 10); e1.record(); torch.cuda.synchronize()
    lib.le_debug_ppht_trace(None, 0)
    o = tr[cap-2:].cpu().numpy().view(np.int64).ravel()
    print("n", n, "ms", round(e0.elapsed_time(e1), 3), "cycles: snapshot", o[0], "vote batches", o[1], "walk1", o[2], "walk2+refresh", o[3], "batches", o[4], "votes", o[5], "segs", int(counts[0]))
