"""CPU oracle for the MiniCV match + two-view-estimation hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in ``minicv_b200`` (the product) may import
this package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs do, and only as the checker or as
the timed CPU baseline -- never as a fallback for the CUDA path.

The reference (``/root/reference``) forwards every function on this path to an
un-vendored, un-pinned OpenCV (``src/MiniCVNative/CMakeLists.txt:31``,
``buildnative.sh:15,48``), so the arithmetic restated here is OpenCV's, pinned
by this repo to ``cv2`` 4.13.0 (opencv-python-headless 4.13.0.92).  The
reference itself holds no golden vectors for this path (``src/Test/Program.fs``
has no assertions), so every restatement is pinned instead against outputs of
cv2 4.13.0 frozen under ``tests/golden/`` by ``tests/golden/make_golden.py``.
"""
