"""`bundle_result_file` (dbat_b200/report.py) against the result files the reference itself wrote.

Every golden report under tests/golden/ is regenerated from its input project and diffed line by line;
only the lines that describe the run rather than the result (UUID, paths, date, versions, host, CPU
times) are exempt.  A report pins, to its printed precision, every IO/EO value and standard
deviation, the >95 % correlations, the chi-square significances, coverage, ray counts, residual and
precision extremes, ray angles and the control point tables.  On CPU the covariances come from the
oracle's bundle_cov; the `-m gpu` twin in test_gpu_parity.py feeds the same writer from the device.
"""
import copy
import os
import re

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, 'golden')

RUN_SPECIFIC = re.compile(
    r'^\s*(Computation UUID|Input file name|Ctrl pt file|EO file|Last Bundle Run:|DBAT version:|MATLAB version:|'
    r'Host system:|Host name:|Bundle:|Post-cov (prep|CIO|CEO|COP):)')


def report_diff(lines, golden_path):
    """Lines that differ between a generated report and a golden one, run-specific lines aside."""
    gold = [l.rstrip('\n') for l in open(golden_path)]
    bad = []
    if len(gold) != len(lines):
        bad.append(('length', len(gold), len(lines)))
    for n, (a, b) in enumerate(zip(gold, lines)):
        if a != b and not (RUN_SPECIFIC.match(a) and RUN_SPECIFIC.match(b)):
            bad.append((n + 1, a, b))
    return bad


def camcal_pm_run(pm, x0desc='Camera calibration from EXIF value'):
    from oracle.loaders import camcal_pm_struct
    from oracle.photogrammetry import resect, forwintersect
    from oracle.bundle import bundle as obundle
    G = os.path.join(GOLD, 'camcalpm')
    s = camcal_pm_struct(os.path.join(G, pm), os.path.join(G, 'camcal-fixed.txt'))
    cpId = np.asarray(s.OP.id)[s.prior.OP.isCtrl]
    s1, _, fail = resect(s, 'all', cpId, 1, 0, cpId)
    s2, _, _ = forwintersect(s1, 'all', True)
    s2.proj.x0desc = x0desc                                  # camcaldemo.m:109
    return obundle(copy.deepcopy(s2), 'gna')


@pytest.mark.parametrize('pm,report', [
    ('camcal-pmexport.txt', 'dbatexports/camcal-dbatreport.txt'),
    ('camcal-pmexport5.txt', 'camcalpm/camcal-dbatreport5.txt')])
def test_camcal_result_files_reproduce_the_reference_reports(pm, report, tmp_path):
    """camcaldemo.m / camcaldemo2.m: the 610- and 313-line result files, line for line."""
    from oracle.bundle import bundle_cov as ocov
    from dbat_b200.report import bundle_result_file
    s, ok, it, s0, E = camcal_pm_run(pm)
    f = tmp_path / 'report.txt'
    s, lines = bundle_result_file(s, E, str(f), cov=ocov)
    assert report_diff(lines, os.path.join(GOLD, report)) == []
    assert open(f).read().split('\n')[:-1] == lines          # what is returned is what is written
    # the struct comes back with posterior std / cov filled in (bundle_result_file.m:139-173)
    assert s.post.std.IO.shape == s.IO.val.shape and s.post.std.EO.shape == s.EO.val.shape
    assert s.post.std.OP.shape == s.OP.val.shape and s.post.cov.OP.shape == (3, 3, s.OP.val.shape[1])
    assert np.all(s.post.std.OP[:, s.prior.OP.isCtrl] == 0)
