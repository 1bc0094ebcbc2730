"""ORACLE -- test infrastructure, NOT product code.

CPU restatement (numpy / scipy / plain C) of the algorithm the reference runs on the
`run_interaction` path and its siblings.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import this package; the product
(`cellregmap_b200`) never does.

Pinning status
--------------
* `oracle.math_port` (QSCov/PMat/ScoreStatistic, dense P/score forms, modified Liu, qmin):
  PINNED by the reference's own known answers in cellregmap/test/test_math.py:38-91
  (tests/test_oracle_goldens.py).
* `oracle.sugar_port`, `oracle.brent_port`, `oracle.lmm_port`, `oracle.chiscore_port`,
  `oracle/qfc_oracle.c`, and therefore every end-to-end output of `oracle.crm_port`:
  PARITY UNPINNED.  They restate third-party dependencies that are pinned only by lower
  bounds in the reference's setup.cfg:27-34 (glimix-core>=3.1.12, numpy-sugar>=1.5.1,
  chiscore>=0.2.3 -> chi2comb, brent-search, optimix), are not vendored under /root/reference,
  are not installed in this image and cannot be installed (no network).  The restatement
  follows the published algorithms (Brent 1973 localmin; Davies 1980 AS 155; Liu-Tang-Zhang
  2009 with the SKAT kurtosis modification; Lippert et al. 2011/2014 FaST-LMM likelihood) and
  the reference's own call sites, and is cross-checked against independent formulations
  (dense textbook likelihoods, Imhof quadrature, scipy distributions) in tests/.
"""
