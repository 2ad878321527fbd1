"""CPU oracle for the pybot geometry-and-matching hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``pybot_b200/`` imports this package.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs may import or execute anything in here, and only
as the checker or the timed CPU baseline - never as the product path.

Layout
------
``ref_loader``  imports the real reference from ``/root/reference`` (authoring
                container only; the GPU box does not have it).
``port``        restatement of the reference's Python on the path (numpy +
                the OpenCV calls the reference itself makes), file:line cited.
``model``       pure-numpy models of the third-party arithmetic the reference
                delegates to OpenCV 4.13.0 (reprojectImageTo3D, projectPoints
                + Jacobian, Rodrigues, BFMatcher Hamming kNN), pinned against
                cv2 by ``tests/test_oracle_pinning.py``.
``c/``          C restatement of BFMatcher's Hamming kNN for sizes where the
                numpy model is too slow (built by ``oracle/build.py``).
``make_golden`` regenerates ``tests/golden/*.npz`` by running the real
                reference (needs ``/root/reference``).

Parity pinning status: the reference holds NO numerical
tests for this path.  a1-a8 (RigidTransform.__mul__, Camera.project,
StereoCamera.reconstruct*) are pinned by running the reference itself here and
committing the outputs as golden fixtures.  a9 (Hamming kNN) and a10 (BA
residual/Jacobian) have no implementation in the reference at all: PARITY
UNPINNED by the reference; they are pinned against the installed OpenCV
(cv2.BFMatcher, cv2.projectPoints) which is what a pybot user would call.
"""
