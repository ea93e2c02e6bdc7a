"""dbat_b200 — B200-native inner loop of the Damped Bundle Adjustment Toolbox.

Public surface mirrors the reference entry points for the hot path:
bundle, levenberg_marquardt, levenberg_marquardt_powell, gauss_newton_armijo, bundle_cov,
and the start-value steps that precede them in every demo, resect and forwintersect.
All numerical work runs in libdbatgpu.so (hand-written sm_100a CUDA, include/dbat_gpu.h).
"""
from .bundle import (Problem, bundle, bundle_cov, gauss_markov, gauss_newton_armijo, levenberg_marquardt,
                     levenberg_marquardt_powell, make_termfun)
from .photogrammetry import forwintersect, resect
from .dbatstruct import (buildserialindices, buildweightmatrix, deserialize, new_struct,
                         serialize, seteoest_depend)

__all__ = ['Problem', 'bundle', 'bundle_cov', 'gauss_markov', 'gauss_newton_armijo', 'levenberg_marquardt',
           'levenberg_marquardt_powell', 'make_termfun', 'forwintersect', 'resect', 'buildserialindices',
           'buildweightmatrix', 'deserialize', 'new_struct', 'serialize', 'seteoest_depend']
