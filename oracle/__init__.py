"""CPU oracle for the line-detection hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import this package.  The product package (``graphics-project_b200``) never does.

Two oracles live here:

* ``oracle.restate`` -- ctypes binding of ``cv_restate.c``, a plain-C restatement of the OpenCV
  arithmetic the reference calls (gray, blur, normalize, Canny, HoughLines, HoughLinesP, LSD).  It
  exposes the intermediates cv2 hides (NMS map, full accumulator).  Pinned bit-exact against the
  cv2 4.13.0 wheel of this image and the golden vectors in ``tests/golden``.
* ``oracle.geometry`` -- float64 restatement of the reference's segment-merge and face-quad search (util.py,
  opencv_renewed.py), pinned on the reference's own outputs (``tests/golden/golden_merge.npz``, ``golden_faces.npz``).
* ``oracle.vp_ref``   -- restatement of the reference's pure-Python vanishing-point functions
  (opencv_fit_color.py:14-47, 95-101; opencv_renewed.py:16-17): the seeded random search of ``regress_lines``
  with the reference's own scalar loops (the CPU baseline of the ``lines`` workload) and a vectorised scorer.

The reference's cv2 call chains themselves (the reference *is* "Python calling cv2"; it pins
opencv-python==4.10.0.84, this image ships 4.13.0) are issued verbatim against the real ``cv2`` module by
``bench.py``'s CPU legs and by ``tests/test_oracle.py::test_oracle_vs_live_cv2``.
"""
