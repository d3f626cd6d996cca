"""
oracle/ -- CPU restatement of the geosss geodesic slice-sampling hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import anything from here, and only as the
checker or as the timed CPU baseline -- never as a fallback for the CUDA path.
``geosss_b200`` never imports this package.

Every function cites the reference file:line (relative to the upstream geosss
v0.3.5 tree) it restates.  Arithmetic is float64 numpy/scipy exactly like the
reference, so tape-fed trajectories agree with it to rounding (<= 1e-15).

Parity status: PINNED.  The reference's own tests hold no numeric vectors for
this path, so the oracle is pinned against outputs of the
reference itself: ``oracle/make_golden.py`` imports the upstream package (in the
build container only) and writes ``tests/golden/*.npz``; ``tests/test_oracle_golden.py``
checks every oracle function against those vectors.
"""

from . import sphere, targets, samplers, tape, philox, ess, exact  # noqa: F401
