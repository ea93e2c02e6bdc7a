"""dbat_b200 — B200-native inner loop of the Damped Bundle Adjustment Toolbox.

Public surface mirrors the reference entry points for the hot path:
bundle, levenberg_marquardt, levenberg_marquardt_powell, gauss_newton_armijo, bundle_cov,
the start-value steps that precede them in every demo, resect and forwintersect, and the
report that follows them, bundle_result_file (with its correlation / significance statistics), and
the ingest in front of them (loadpm, prob2dbatstruct, loadcpt/matchcpt/setcpt, setcamvals, setcamest,
seteoest, cleareo, clearop), so that a reference demo script ports call for call.
All numerical work runs in libdbatgpu.so (hand-written sm_100a CUDA, include/dbat_gpu.h).
"""
from .bundle import (Problem, bundle, bundle_cov, gauss_markov, gauss_newton_armijo, levenberg_marquardt,
                     levenberg_marquardt_powell, make_termfun)
from .photogrammetry import forwintersect, resect
from .report import (angles, bundle_residuals, bundle_result_file, camangles, corrmat, coverage, cumchi2,
                     high_eo_correlations, high_io_correlations, high_op_correlations,
                     test_distortion_params, writestats)
from .ingest import (cleareo, clearop, filterprob, legacyloadeotable, loadcpt, loadctrlpts, loadeotable, loadimagepts,
                     loadimagetable, loadpm, loadpm3dtbl, loadpmreport, loadpsz, ps2pmstruct, matchcpt, matcheo, prob2dbatstruct,
                     setcamest, setcamvals, setcpt, seteoest, setprioreo)
from .script import rundbatscript
from .dbatstruct import (buildserialindices, buildweightmatrix, deserialize, new_struct,
                         serialize, seteoest_depend)

__all__ = ['Problem', 'bundle', 'bundle_cov', 'gauss_markov', 'gauss_newton_armijo', 'levenberg_marquardt',
           'levenberg_marquardt_powell', 'make_termfun', 'forwintersect', 'resect', 'bundle_result_file', 'corrmat', 'cumchi2',
           'high_io_correlations', 'high_eo_correlations', 'high_op_correlations', 'test_distortion_params',
           'bundle_residuals', 'coverage', 'angles', 'camangles', 'writestats', 'loadpm', 'prob2dbatstruct', 'loadcpt', 'matchcpt', 'setcpt',
           'setcamvals', 'setcamest', 'seteoest', 'cleareo', 'clearop', 'legacyloadeotable', 'matcheo', 'setprioreo', 'loadimagepts', 'loadctrlpts',
           'loadimagetable', 'loadeotable', 'loadpm3dtbl', 'loadpmreport', 'filterprob', 'loadpsz', 'ps2pmstruct', 'rundbatscript', 'buildserialindices',
           'buildweightmatrix', 'deserialize', 'new_struct', 'serialize', 'seteoest_depend']
