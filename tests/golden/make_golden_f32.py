"""Golden fixtures of the float32 frame-buffer path, made by the reference itself.

Run ONCE in the build container (needs /root/reference and the cv2 wheel):

    python tests/golden/make_golden_f32.py

A fake frame buffer (the only thing ``_fboToImage`` needs: ``read``, ``height``, ``width``) holds each stored
golden input as the simulator would: RGB float32 in [0,1], bottom-up.  The reference's own
``opencv._fboToImage`` (opencv.py:101-108) and ``opencv.doCanny`` (opencv.py:21-28) then run on it, exactly as
``postProcessFbo`` (opencv.py:110-113) calls them.  Stored in golden_f32.npz per image: the edge map
(bit-packed) and SHA-1 prefixes of the float32 intermediates (gray, blur, normalize) taken with the same cv2
calls doCanny makes.  The float input is NOT stored: tests rebuild it from golden_inputs.npz with `fbo_bytes`.
"""
import hashlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import import_reference  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def fbo_bytes(bgr_u8):
    """What fbo.read(components=3, dtype='f4') returns for a frame showing this image: RGB, bottom-up, [0,1]."""
    return np.ascontiguousarray(bgr_u8[::-1, :, ::-1].astype(np.float32) / np.float32(255)).tobytes()


class FakeFbo:
    def __init__(self, bgr_u8):
        self.height, self.width = bgr_u8.shape[:2]
        self._bytes = fbo_bytes(bgr_u8)

    def read(self, components=3, dtype="f4"):
        assert components == 3 and dtype == "f4"
        return self._bytes


def sha(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()[:10]


def main():
    import cv2 as cv
    inputs = np.load(os.path.join(OUT, "golden_inputs.npz"))
    opencv, _, _ = import_reference()
    out = {}
    for name in inputs.files:
        img = inputs[name]
        view = opencv._fboToImage(FakeFbo(img))
        assert view.dtype == np.float32 and view.strides[0] < 0 and view.strides[2] < 0
        edges = opencv.doCanny(view)
        g = cv.cvtColor(view, cv.COLOR_BGR2GRAY)
        b = cv.GaussianBlur(g, (5, 5), 0)
        nm = cv.normalize(b, None, 0, 255, cv.NORM_MINMAX)
        assert np.array_equal(edges, cv.Canny(nm.astype(np.uint8), 30, 50))
        out[name + "/edges"] = np.packbits(edges > 0)
        out[name + "/sha"] = np.array([sha(g), sha(b), sha(nm), sha(nm.astype(np.uint8))])
        print(name, img.shape, int((edges > 0).sum()), out[name + "/sha"])
    np.savez_compressed(os.path.join(OUT, "golden_f32.npz"), **out)


if __name__ == "__main__":
    main()
