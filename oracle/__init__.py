"""ORACLE - CPU restatement of the reference's hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package; the product path
(``cse_b200``) never does and fails loudly when its CUDA library is missing.

* ``ops`` / ``models`` - Keras 2.2.4 / TF 1.15 layer semantics and the four member
  architectures in torch CPU (fp64 gold, fp32 = the reference's own precision).
  PARITY UNPINNED: the reference's arithmetic lives in keras==2.2.4 and
  tensorflow-gpu==1.15.0 (requirements.txt:3-4), neither installable here, and the
  reference ships no tests / vectors / weights.
* ``vote`` - numpy restatement of the soft vote + CSV text round trip.  PINNED against
  vectors produced by the reference's own functions (tools/make_golden_vote.py).
* ``resize`` - numpy restatement of select_frames + OpenCV's 8-bit INTER_LINEAR resize.  PINNED against cv2 itself
  and against the reference's own clip loaders (tools/make_golden_clips.py).
* ``bench.py`` additionally uses ``vote`` for the post-timing self-check of the predictions it timed.
"""
