"""How much of cv2.HoughLinesP's output survives a different pixel visiting order?  (CPU only, oracle + cv2.)

north_star allows HoughLinesP to match cv2 "under the same seeded pixel visiting order, or within 1 px of endpoint
distance with >= 99 % segment recall".  This script runs cv2's own algorithm (the oracle restatement, checked equal
to cv2 with cv2's seed) with OTHER seeds of cv::RNG and measures the recall of cv2's segments: which pick first
reaches the threshold on a line decides the (rho, theta) cell the walk follows, so endpoints move by many pixels
and segments split differently.  Result (profiles/r2_ppht_order_sensitivity.txt): 22-33 % recall at 1 px on
cube-edge frames, ~10 % on cave frames -- cv2 does not meet the relaxed criterion against itself, so no order-free
variant can; the engine therefore implements the exact order only."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cv2
import numpy as np
from oracle import restate as R
import importlib.util
spec = importlib.util.spec_from_file_location("synth", os.path.join(os.path.dirname(__file__), "..", "graphics-project_b200", "synth.py"))
synth = importlib.util.module_from_spec(spec); spec.loader.exec_module(synth)


def recall(ref, got, tol):
    ok = 0
    for s in ref:
        a = np.hypot(got[:, 0] - s[0], got[:, 1] - s[1]); b = np.hypot(got[:, 2] - s[2], got[:, 3] - s[3])
        c = np.hypot(got[:, 0] - s[2], got[:, 1] - s[3]); d = np.hypot(got[:, 2] - s[0], got[:, 3] - s[1])
        ok += bool((np.maximum(a, b) <= tol).any() or (np.maximum(c, d) <= tol).any())
    return ok / max(1, len(ref))


if __name__ == "__main__":
    for kind, hw in (("cubes", (1080, 1920)), ("cubes", (2160, 3840)), ("voxel", (1080, 1920)), ("cave", (400, 600))):
        for frame in range(2):
            e = cv2.Canny(cv2.cvtColor(synth.make_frame(frame, *hw, kind), cv2.COLOR_BGR2GRAY), 5, 150)
            ref = cv2.HoughLinesP(e, 1, np.pi / 180, 30, minLineLength=50, maxLineGap=10)[:, 0]
            assert np.array_equal(R.hough_lines_p(e, 1, np.pi / 180, 30, 50, 10)[:, 0], ref)
            for seed in (12345, 987654321):
                got = R.hough_lines_p(e, 1, np.pi / 180, 30, 50, 10, seed=seed)[:, 0]
                print(f"{kind} {hw[1]}x{hw[0]} frame {frame} seed {seed}: cv2 {len(ref)} segments, other order {len(got)}; "
                      f"recall of cv2's segments at 1 px {recall(ref, got, 1):.3f}, at 3 px {recall(ref, got, 3):.3f}")
