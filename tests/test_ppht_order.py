"""CPU: cv2.HoughLinesP's result is a function of its pixel visiting order.  north_star's relaxed criterion
("within 1 px of endpoint distance with >= 99 % segment recall") is not met by cv2's own algorithm under another
RNG seed, so the engine offers the exact order only (le_hough_lines_p rejects exact_order = 0).  The oracle with
cv2's seed equals cv2; with another seed it recalls well under half of cv2's segments."""
import os
import sys

import numpy as np

from oracle import restate as R

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts"))


def test_other_visiting_order_does_not_reproduce_cv2():
    import cv2
    from ppht_order_sensitivity import recall, synth
    e = cv2.Canny(cv2.cvtColor(synth.make_frame(0, 540, 960, "cubes"), cv2.COLOR_BGR2GRAY), 5, 150)
    ref = cv2.HoughLinesP(e, 1, np.pi / 180, 30, minLineLength=50, maxLineGap=10)[:, 0]
    assert np.array_equal(R.hough_lines_p(e, 1, np.pi / 180, 30, 50, 10)[:, 0], ref)
    assert np.array_equal(R.hough_lines_p(e, 1, np.pi / 180, 30, 50, 10, seed=(1 << 64) - 1)[:, 0], ref)
    for seed in (12345, 987654321):
        got = R.hough_lines_p(e, 1, np.pi / 180, 30, 50, 10, seed=seed)[:, 0]
        assert recall(ref, got, 1.0) < 0.6, "north_star's relaxed HoughLinesP criterion became attainable?"
    # the seed override lasts one call
    assert np.array_equal(R.hough_lines_p(e, 1, np.pi / 180, 30, 50, 10)[:, 0], ref)
