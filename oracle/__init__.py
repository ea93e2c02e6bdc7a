"""CPU oracle for the DBAT bundle-adjustment hot path.

TEST INFRASTRUCTURE ONLY.  This package is a NumPy/SciPy restatement of the
reference MATLAB algorithm (niclasborlin/dbat v0.9.2.0, `code/bundle/**`,
`code/misc/{buildserialindices,serialize,deserialize,buildweightmatrix}.m`).
Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` may import it; the product package
`dbat_b200` never does.

Parity status: PINNED.  The restatement is checked in `tests/test_oracle_golden.py`
against the reference's own golden result files for the camcal XML script
project (`data/script/camcaldemo/result/*`, copied as small fixtures into
`tests/golden/camcaldemo/` by `tests/golden/make_golden.py`): 18-digit IO and EO
estimates, EO posterior standard deviations, sigma0 and the top-50 image
residuals.  No MATLAB/Octave exists in the build container or on the GPU box,
so the reference itself cannot be executed.

Indices are 0-based throughout (the reference is 1-based); every function
cites the reference file:line it follows.
"""
