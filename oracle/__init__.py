"""CPU oracle for the blueice likelihood hot path.  TEST INFRASTRUCTURE ONLY.

This package is a NumPy/SciPy restatement of the reference's algorithm for the hot path
named in BASELINE.json (anchor-grid morphing -> mixture density -> log-likelihood, unbinned
and binned, plus the histogram-template lookup that feeds it).  Every function cites the
reference file:line it follows (paths relative to the reference checkout of
JelleAalbers/blueice v1.2.1).

Rules (task statement, section 3):
  * Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
    legs may import anything from here -- as the checker or as the timed CPU baseline, never
    as part of the product.  `blueice_b200` never imports `oracle`.
  * Parity pinning: the restatement is checked against
      - SciPy's own RegularGridInterpolator / stats.poisson (the third-party numerics the
        reference calls), bit-for-bit, in tests/test_oracle_pins.py;
      - the reference's own known-answer tests (tests/test_BeestonBarlow.py:32,68-76,120-131,
        tests/test_likelihood.py:17-18, tests/test_binned_likelihood.py:21-22);
      - outputs of the unmodified reference run in the build container, committed as
        tests/golden/*.npz by tests/golden/make_golden.py.
    `multihist.Histdd.lookup` (piecewise pdf lookup) is restated from memory of multihist
    0.6.x: PARITY UNPINNED for that one function (no reference test reaches source.py:243).
"""
from . import morph, unbinned, hist, binned, pipeline  # noqa: F401
