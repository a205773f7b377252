"""CPU oracle for the MSACL hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

This package is a NumPy restatement (vectorised over env instances / replay windows) of the
reference algorithms on the path named by BASELINE.json:north_star.  Each function cites the
reference file:line it follows.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import it; the product package
(`msacl_b200`) never does and fails loudly when its CUDA library is missing.

Parity pin: the reference ships no tests or golden vectors, so the oracle
is pinned against outputs of the reference itself, generated in the build container by
`tests/golden/make_golden.py` (imports `/root/reference` through a gymnasium stub) and
committed as `tests/golden/*.npz`; `tests/test_oracle_golden.py` replays them.

Numerics contract: the reference's dtype flow under NumPy 2.x (Python-float constants are
"weak", so scalar math on float32 state stays float32; float64 appears only where the
reference builds float64 arrays).  The oracle reproduces that flow operation by operation.
"""
