"""CPU oracle for the line-detection hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import this package.  The product package (``graphics-project_b200``) never does.

Two oracles live here:

* ``oracle.restate`` -- ctypes binding of ``cv_restate.c``, a plain-C restatement of the OpenCV
  arithmetic the reference calls (gray, blur, normalize, Canny, HoughLines, HoughLinesP, LSD).  It
  exposes the intermediates cv2 hides (NMS map, full accumulator).  Pinned bit-exact against the
  cv2 4.13.0 wheel of this image and the golden vectors in ``tests/golden``.
* ``oracle.cvref``   -- the reference's own call chains issued verbatim against the real ``cv2``
  module (the reference *is* "Python calling cv2"; reference pins opencv-python==4.10.0.84,
  this image ships 4.13.0), and numpy restatements of its pure-Python VP functions.
"""
