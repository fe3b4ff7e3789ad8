"""Golden vectors for the quad search get_faces_from_pairs, from the reference itself.

Run ONCE in the build container (needs /root/reference and the cv2 wheel):

    python tests/golden/make_golden_faces.py

For each fixture image the reference's own chain runs exactly as its `lines`/`mixed` pipelines do
(opencv_renewed.py:160-178): lsd(image) -> classifyEdges (np.random seeded, regress_lines is a random search)
-> smoothEdges -> splitEdges -> get_faces_from_pairs(x, y), (z, x), (y, z).  Stored in golden_faces.npz:
the three classified edge lists that enter get_faces_from_pairs, as (n, 4) float64 [x1, y1, x2, y2] (exact
values; `<name>/<axis>_f32` marks the segments that were float32 in the reference, which then does its
arithmetic in float32) and the three returned face lists as (m, 4, 2) float64.
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import import_reference  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
IMAGES = ["demo_scarce", "sc_cube", "sc_white_2", "sc_scarce_3_flat", "old_sc_inf"]


def edge_array(edges):
    a = np.array([[e[0][0], e[0][1], e[1][0], e[1][1]] for e in edges], dtype=np.float64).reshape(-1, 4)
    f32 = np.array([np.asarray(e[0]).dtype == np.float32 and np.asarray(e[1]).dtype == np.float32 for e in edges],
                   dtype=bool)
    return a, f32


def main():
    inputs = np.load(os.path.join(OUT, "golden_inputs.npz"))
    opencv, fit, renewed = import_reference()
    out = {}
    for name in IMAGES:
        img = inputs[name]
        np.random.seed(1234)
        edges = opencv.lsd(img)
        x, y, z, phi, theta = renewed.classifyEdges(edges)
        x, y, z = renewed.smoothEdges(x, y, z)
        x, y, z = renewed.splitEdges(x, y, z, 0.1)
        for ax, e in (("x", x), ("y", y), ("z", z)):
            out[f"{name}/{ax}"], out[f"{name}/{ax}_f32"] = edge_array(e)
        for key, (e1, e2) in (("zfaces", (x, y)), ("yfaces", (z, x)), ("xfaces", (y, z))):
            t0 = time.time()
            faces = renewed.get_faces_from_pairs(e1, e2)
            out[f"{name}/{key}"] = np.array(faces, dtype=np.float64).reshape(-1, 4, 2)
            print(name, key, len(e1), len(e2), "->", len(faces), f"{time.time() - t0:.1f}s")
    np.savez_compressed(os.path.join(OUT, "golden_faces.npz"), **out)


if __name__ == "__main__":
    main()
