"""CPU oracle for the EM-POSE LGD hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and there only as the checker.  The
product path (``em-pose_b200/``) never imports this package and fails loudly
when its CUDA library is missing.

Parity status
-------------
* Everything the reference ships in-tree for this path (IEF loop, MLP, LSTM
  wrapper, sensor frames, reconstruction loss) is PINNED: ``tests/golden`` holds
  outputs of the unmodified reference modules run in the build container
  (``tests/golden/make_golden.py``), and the restatement here reproduces them.
* The SMPL-H LBS arithmetic itself lives in a third-party dependency that is
  absent from ``/root/reference`` (human_body_prior fork @ 821a0e7e,
  reference ``requirements.txt:9``) and the reference has no tests or golden
  vectors at that boundary, so that part is "parity unpinned": it restates the
  published SMPL / smplx ``lbs`` formulation and is checked by known-answer and
  self-consistency tests only (see ``oracle/smplh_lbs.py``).
"""
