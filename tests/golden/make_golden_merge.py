"""Golden vectors for the segment-merging step, from the reference itself.

Run ONCE in the build container (needs /root/reference):

    python tests/golden/make_golden_merge.py

Feeds the reference's own prob() / lsd() outputs stored in golden_outputs.npz (ref_prob int32, ref_lsd
float32, as lists of point pairs exactly like opencv.py:209,218 return them) to the reference's
util.combineParallelLines (util.py:141-159) and stores the merged segment lists as (m, 4) float64.
"""
import os
import sys

import numpy as np

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    sys.dont_write_bytecode = True
    sys.path.insert(0, REF)
    sys.setrecursionlimit(20000)
    import util  # pure numpy module of the reference
    gold = np.load(os.path.join(OUT, "golden_outputs.npz"))
    out = {}
    for key in gold.files:
        if not (key.endswith("/ref_prob") or key.endswith("/ref_lsd")):
            continue
        a = gold[key]
        pairs = [(np.array(r[0:2]), np.array(r[2:4])) for r in a.reshape(-1, 4)]
        merged = util.combineParallelLines(pairs)
        m = np.array([[l[0][0], l[0][1], l[1][0], l[1][1]] for l in merged], dtype=np.float64).reshape(-1, 4)
        out[key.replace("/ref_", "/merge_")] = m
        print(key, a.dtype, len(pairs), "->", len(m))
    np.savez_compressed(os.path.join(OUT, "golden_merge.npz"), **out)


if __name__ == "__main__":
    main()
