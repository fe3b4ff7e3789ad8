"""Golden vectors for LSD_REFINE_ADV, from the cv2 wheel with the reference's own parameters.

Run ONCE in the build container (needs /root/reference and the cv2 wheel):

    python tests/golden/make_golden_lsd2.py

The reference's simulator calls ``lsd(image, 2, scale=0.5)`` (opencv.py:142-145), i.e.
``cv.createLineSegmentDetector(2, scale=.5, sigma_scale=.6, quant=2.0, ang_th=22.5, log_eps=0, density_th=.7,
n_bins=1024).detect(gray)``.  Stored per golden input image, for scale .5 and .8: lines (N,4) float32, width,
prec, nfa (N,) float64; `<name>/ref_lsd2` is what the reference's own wrapper returns for (image, 2, scale=0.5).
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import import_reference  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    import cv2 as cv
    inputs = np.load(os.path.join(OUT, "golden_inputs.npz"))
    opencv, _, _ = import_reference()
    out = {}
    for name in inputs.files:
        img = inputs[name]
        g = cv.cvtColor(img, cv.COLOR_BGR2GRAY)
        for scale, key in ((.5, "h"), (.8, "f")):
            d = cv.createLineSegmentDetector(2, scale=scale, sigma_scale=.6, quant=2.0, ang_th=22.5, log_eps=0.0,
                                             density_th=.7, n_bins=1024)
            lines, width, prec, nfa = d.detect(g)
            out[f"{name}/lsd2{key}"] = lines[:, 0]
            out[f"{name}/lsd2{key}_w"] = width[:, 0]
            out[f"{name}/lsd2{key}_p"] = prec[:, 0]
            out[f"{name}/lsd2{key}_nfa"] = nfa[:, 0]
            print(name, scale, len(lines))
        pairs = opencv.lsd(img, 2, scale=0.5)
        ref = np.array([[a[0], a[1], b[0], b[1]] for a, b in pairs], dtype=np.float32)
        assert np.array_equal(ref, out[f"{name}/lsd2h"])
        out[f"{name}/ref_lsd2"] = ref
    np.savez_compressed(os.path.join(OUT, "golden_lsd2.npz"), **out)


if __name__ == "__main__":
    main()
