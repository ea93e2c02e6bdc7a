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


_NUM = re.compile(r'-?\d+\.?\d*(?:e[-+]?\d+)?')


_TIE = re.compile(r'^(\s*(?:Max|X|Y|Z): -?[0-9.]+ ou) \(CP\d+, pt \d+\)\s*$')


def _same(a, b, rtol):
    """Equal up to rtol on every printed number (and one unit of its last printed digit)."""
    if a == b:
        return True
    if rtol > 0 and _TIE.match(a) and _TIE.match(b) and _TIE.match(a).group(1) == _TIE.match(b).group(1):
        # "X: 0.001 ou (CP2, pt 1002)": the arg-max label of a maximum over control points.  When two points tie to
        # the printed digits (prague2016 weighted: |dZ| of CP2 and CP4 differ by 2e-9 relative, 1e-12 of the
        # coordinate - below the convergence tolerance of the run) the label is decided by rounding noise.
        return True
    if rtol == 0 or _NUM.sub('#', a) != _NUM.sub('#', b):
        return False
    for x, y in zip(_NUM.findall(a), _NUM.findall(b)):
        mant = x.lower().split('e')
        ulp = 10.0 ** (-len(mant[0].split('.')[1]) if '.' in mant[0] else 0) * 10.0 ** (int(mant[1]) if len(mant) > 1 else 0)
        if abs(float(x) - float(y)) > rtol * abs(float(x)) + 1.01 * ulp:
            return False
    return True


def _drop_nullspace(lines):
    """The printed null-space basis of a rank-deficient run is one of many (bundle.m:396-407)."""
    out, skip = [], False
    for l in lines:
        if 'Null-space suggest' in l:
            skip = True
        elif 'Problems related to the processing' in l:
            skip = False
        if not skip:
            out.append(l)
    return out


def report_diff(lines, golden_path, rtol=0.0, first_error_rtol=None):
    """Lines that differ between a generated report and a golden one, run-specific lines aside.
    first_error_rtol: separate tolerance for the 'First error' line, which depends on the start values
    only (resection of an ill-conditioned quartic: solver and LAPACK-build dependent)."""
    gold = _drop_nullspace([l.rstrip('\n') for l in open(golden_path)])
    lines = _drop_nullspace(lines)
    bad = []
    if len(gold) != len(lines):
        bad.append(('length', len(gold), len(lines)))
    for n, (a, b) in enumerate(zip(gold, lines)):
        tol = first_error_rtol if first_error_rtol is not None and 'First error:' in a else rtol
        if not _same(a, b, tol) and not (RUN_SPECIFIC.match(a) and RUN_SPECIFIC.match(b)):
            bad.append((n + 1, a, b))
    return bad


def camcal_pm_run(pm, x0desc='Camera calibration from EXIF value', model=3):
    from oracle.loaders import camcal_pm_struct
    from oracle.photogrammetry import resect, forwintersect
    from oracle.bundle import bundle as obundle
    G = os.path.join(GOLD, 'camcalpm')
    s = camcal_pm_struct(os.path.join(G, pm), os.path.join(G, 'camcal-fixed.txt'), model=model)
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


@pytest.mark.parametrize('pm', ['missing-obs', '1ray'])
def test_failed_run_result_files_reproduce_the_reference_reports(pm):
    """camcaldemo_missing_obs.m / camcaldemo_1ray.m: code -4 at iteration 0; the result file lists the
    structurally suspect parameters, NaN deviations (no factorisation) and the start values.  Photo 21's
    start value comes from a resection whose quartic has a near-double root, where LAPACK builds differ
    in the 7th digit (see test_gpu_parity's resect tolerance): 1e-5 relative on the printed numbers."""
    from oracle.bundle import bundle_cov as ocov
    from dbat_b200.report import bundle_result_file
    s, ok, it, s0, E = camcal_pm_run('camcal-pmexport-%s.txt' % pm)
    assert E.code == -4
    s, lines = bundle_result_file(s, E, None, cov=ocov)
    assert report_diff(lines, os.path.join(GOLD, 'camcalpm', 'camcal-dbatreport-%s.txt' % pm), rtol=1e-5) == []


def test_no_datum_result_file_reproduces_the_reference_report():
    """camcaldemo_no_datum.m: code -2, numerical rank 428 of 435; everything but the (non-unique)
    null-space basis is reproduced exactly."""
    from oracle.loaders import camcal_pm_struct
    from oracle.bundle import bundle as obundle, bundle_cov as ocov
    from dbat_b200.report import bundle_result_file
    G = os.path.join(GOLD, 'camcalpm')
    s = camcal_pm_struct(os.path.join(G, 'camcal-pmexport.txt'), None, keep_loaded=True)
    s, ok, it, s0, E = obundle(copy.deepcopy(s), 'gna')
    s, lines = bundle_result_file(s, E, None, cov=ocov)
    assert '         Numerical rank: 428 (deficiency: 7)' in lines
    assert sum('Vector ' in l for l in lines) == 7
    assert report_diff(lines, os.path.join(G, 'camcal-dbatreport-no-datum.txt')) == []


@pytest.mark.parametrize('model', [-1, 1, 2, 3, 4, 5])
def test_all_distortion_models_result_files_reproduce_the_reference_reports(model):
    """camcaldemo_allmodels.m: the same project under every lens distortion model (the forward model -1,
    the legacy backward models 1-2 without affine terms, 3-5): six 608/610-line result files, exactly."""
    from oracle.bundle import bundle_cov as ocov
    from dbat_b200.report import bundle_result_file
    s, ok, it, s0, E = camcal_pm_run('camcal-pmexport.txt', model=model)
    assert ok and it == 9
    s, lines = bundle_result_file(s, E, None, cov=ocov)
    assert report_diff(lines, os.path.join(GOLD, 'dbatexports', 'camcal-dbatreport-model%d.txt' % model)) == []


def test_script_pipeline_result_file_reproduces_the_reference_report():
    """data/script/camcaldemo (camcaldemo.xml run by the reference's script runner): its result/report.txt."""
    from oracle.loaders import load_camcal_script
    from oracle.photogrammetry import resect, forwintersect
    from oracle.bundle import bundle as obundle, bundle_cov as ocov
    from dbat_b200.report import bundle_result_file
    root = os.path.join(GOLD, 'camcaldemo')
    s = load_camcal_script(root)
    cpId = np.asarray(s.OP.id)[s.prior.OP.isCtrl]
    s1, _, fail = resect(s, 'all', cpId, 1, 0, cpId)
    s2, _, _ = forwintersect(s1, 'all', True)
    s3, ok, it, s0, E = obundle(copy.deepcopy(s2), 'gna')
    s3, lines = bundle_result_file(s3, E, None, cov=ocov)
    assert report_diff(lines, os.path.join(root, 'result', 'report.txt')) == []


def prague_run(root, stub, cpfile=None, prior_eo=False, **kw):
    from oracle.loaders import prague_cam_struct, load_pm_export, set_prior_eo_positions
    from oracle.photogrammetry import resect, forwintersect
    from oracle.bundle import bundle as obundle
    s = prague_cam_struct(root, stub, cpfile, **kw)
    if prior_eo:
        prob = load_pm_export(os.path.join(root, 'pmexports', 'wsmart-with-orient-pmexport.txt'))
        set_prior_eo_positions(s, prob, os.path.join(root, 'ref', 'fake-camera-positions.txt'))
    s.EO.val[:] = np.nan                                     # cleareo / clearop (prague2016_pm.m:194-195)
    s.OP.val[:, ~s.prior.OP.isCtrl] = np.nan
    cpId = np.asarray(s.OP.id)[s.prior.OP.isCtrl]
    s1, _, fail = resect(s, 'all', cpId, 1, 0, cpId)
    assert not fail
    s2, _, _ = forwintersect(s1, 'all', True)
    return obundle(copy.deepcopy(s2), 'gna')


@pytest.mark.parametrize('orient', ['no', 'with'])
@pytest.mark.parametrize('project,stub,cps', [
    ('prague2016cam', 'fixed', None), ('prague2016cam', 'weighted', None),
    ('prague2016sxb', 'f-op0', 'fixed'), ('prague2016sxb', 'w-op0', 'weighted'),
    ('prague2016sxb', 'w-op1', 'weighted'), ('prague2016sxb', 'wsmart', 'weighted')])
def test_prague2016_result_files_reproduce_the_reference_reports(project, stub, cps, orient):
    """prague2016_pm('c1','c2','s1'..'s4'): fixed camera (legacy model 1), fixed or weighted control
    points (prior OP observations, so the Ctrl measurements tables carry real prior/posterior/diff
    numbers), a check point (s3) and 1100 smart points (s4), each from the export PhotoModeler wrote
    without and with its own orientation stage - twelve result files, exactly."""
    from oracle.bundle import bundle_cov as ocov
    from dbat_b200.report import bundle_result_file
    root = os.path.join(GOLD, project)
    s, ok, it, s0, E = prague_run(root, stub, 'ctrlpts-%s.txt' % cps if cps else None, orient=orient)
    assert ok
    s, lines = bundle_result_file(s, E, None, cov=ocov)
    assert report_diff(lines, os.path.join(root, 'dbatexports', '%s-%s-orient-dbatreport.txt' % (stub, orient))) == []


@pytest.mark.parametrize('use_prior', [True, False])
def test_sxb_prior_eo_result_files_reproduce_the_reference_reports(use_prior):
    """sxb_prior_eo.m: with / without prior observations of four camera positions.  Exact but for the
    'First error' line, which is a function of the resection start values only (ill-conditioned quartic,
    6e-6 and 1e-5 relative, cf. test_sxb_prior_eo_demo_matches_golden_reports): 2e-5 on that line."""
    from oracle.bundle import bundle_cov as ocov
    from dbat_b200.report import bundle_result_file
    root = os.path.join(GOLD, 'prague2016sxb')
    s, ok, it, s0, E = prague_run(root, 'wsmart', 'ctrlpts-weighted.txt', use_prior, shift_cp=False, orient='with')
    assert ok
    s, lines = bundle_result_file(s, E, None, cov=ocov)
    rep = os.path.join(root, 'dbatexports', 'sxb-%sprior-eo-dbatreport.txt' % ('' if use_prior else 'no-'))
    exact = report_diff(lines, rep)
    assert len(exact) == 1 and 'First error' in exact[0][1]
    assert report_diff(lines, rep, rtol=2e-5) == []


# ----------------------------------------------------------------------------- device twins
def _device_pipeline(s):
    import dbat_b200
    cpId = np.asarray(s.OP.id)[s.prior.OP.isCtrl]
    s1, _, fail = dbat_b200.resect(s, 'all', cpId, 1, 0, cpId)
    assert not fail
    s2, _, _ = dbat_b200.forwintersect(s1, 'all', True)
    return dbat_b200.bundle(s2, 'gna')


@pytest.mark.gpu
@pytest.mark.parametrize('pm,report', [
    ('camcal-pmexport.txt', 'dbatexports/camcal-dbatreport.txt'),
    ('camcal-pmexport5.txt', 'camcalpm/camcal-dbatreport5.txt')])
def test_camcal_result_files_from_the_device(pm, report):
    """The camcal demos with every number from the device: start values (resect, forwintersect), bundle,
    and CIOF / CEO / COP from the device factorisation through the default `bundle_cov`.  The result file
    is the reference's to the printed digits; 1e-5 relative covers the 'First error' line (device
    resection of image 21, see test_gpu_parity) and last-digit rounding."""
    from oracle.loaders import camcal_pm_struct
    from dbat_b200.report import bundle_result_file
    G = os.path.join(GOLD, 'camcalpm')
    s = camcal_pm_struct(os.path.join(G, pm), os.path.join(G, 'camcal-fixed.txt'))
    s.proj.x0desc = 'Camera calibration from EXIF value'
    s, ok, it, s0, E = _device_pipeline(s)
    assert ok
    s, lines = bundle_result_file(s, E)
    assert report_diff(lines, os.path.join(GOLD, report), rtol=1e-5, first_error_rtol=1e-4) == []


@pytest.mark.gpu
@pytest.mark.parametrize('stub', ['fixed', 'weighted'])
def test_prague2016_result_files_from_the_device(stub):
    """prague2016_pm('c1'|'c2') on the device: legacy model 1, fixed camera, fixed / weighted control
    points (prior OP observations enter the posterior covariances and the Ctrl measurement tables)."""
    from oracle.loaders import prague_cam_struct
    from dbat_b200.report import bundle_result_file
    root = os.path.join(GOLD, 'prague2016cam')
    s = prague_cam_struct(root, stub)
    s.EO.val[:] = np.nan
    s.OP.val[:, ~s.prior.OP.isCtrl] = np.nan
    s, ok, it, s0, E = _device_pipeline(s)
    assert ok
    s, lines = bundle_result_file(s, E)
    assert report_diff(lines, os.path.join(root, 'dbatexports', '%s-no-orient-dbatreport.txt' % stub), rtol=1e-5,
                       first_error_rtol=1e-4) == []


def test_result_file_through_the_device_covariance_interface():
    """The default covariance provider is `dbat_b200.bundle_cov`, which assembles sparse block-diagonal
    CEO / COP and CIOF from what the device returns (`Problem.cov`: (N,k,k) blocks, camera part of CXX).
    Here a stand-in Problem serves those arrays from the oracle, so the whole host path of the device
    twins above - sparse block extraction included - is exercised without a GPU."""
    from oracle.bundle import bundle_cov as ocov
    from dbat_b200.report import bundle_result_file, _diag_blocks

    s, ok, it, s0, E = camcal_pm_run('camcal-pmexport5.txt')

    class StandIn:
        n = E.numParams

        def cov(self, which, s0_):
            assert s0_ == E.s0
            if which == 'cxx_cam':
                nC = self.n - len(s.bundle.serial.OP.dest)
                return np.asarray(ocov(s, E, 'CXX'))[:nC, :nC]
            k = {'cio': s.IO.val.shape[0], 'ceo': 6, 'cop': 3}[which]
            return _diag_blocks(np.asarray(ocov(s, E, which.upper())), k)

    E.problem = StandIn()
    import dbat_b200
    for w in ('CIOF', 'CEOF', 'CIO', 'CEO', 'COP'):             # the assembled matrices equal the oracle's
        np.testing.assert_allclose(dbat_b200.bundle_cov(s, E, w).toarray(), np.asarray(ocov(s, E, w)), rtol=0, atol=1e-18)
    s, lines = bundle_result_file(s, E)
    assert report_diff(lines, os.path.join(GOLD, 'camcalpm', 'camcal-dbatreport5.txt')) == []


def test_sxb_script_result_file_reproduces_the_reference_report():
    """data/script/sxb (sxb.xml): aerial block with object coordinates of 1e6 m, calibrated camera, 14
    weighted control points, 2 check points (Check measurements tables), two image-point files with
    different sigmas.  Exact but for 'First error' (start values from a resection at 1e6 m: 6e-5)."""
    from oracle.loaders import load_sxb_script
    from oracle.photogrammetry import resect, forwintersect
    from oracle.bundle import bundle as obundle, bundle_cov as ocov
    from dbat_b200.report import bundle_result_file
    root = os.path.join(GOLD, 'sxb')
    s = load_sxb_script(root)
    cpId = np.asarray(s.OP.id)[s.prior.OP.isCtrl]
    s1, _, fail = resect(s, 'all', cpId, 1, 0, cpId)
    assert not fail
    s2, _, _ = forwintersect(s1, 'all', True)
    s3, ok, it, s0, E = obundle(copy.deepcopy(s2), 'gna')
    assert ok and it == 4 and (E.numParams, E.numObs) == (1173, 2434)
    s3, lines = bundle_result_file(s3, E, None, cov=ocov)
    rep = os.path.join(root, 'result', 'report.txt')
    exact = report_diff(lines, rep)
    assert len(exact) == 1 and 'First error' in exact[0][1]
    assert report_diff(lines, rep, rtol=1e-4) == []
    assert sum('Check point delta' in l for l in lines) == 1


def roma_struct():
    from oracle.loaders import load_roma_script
    return load_roma_script(os.path.join(GOLD, 'romabundledemo'))


ROMA_CAMERA = dict(cc=24.5425002994450807, px=18.0816295411089882, py=12.0164475994770967,
                   K1=0.000221523347553201198, K2=-1.86984852928280264e-07)      # result/EOS5DMarkII.xml


def test_roma_script_result_file_reproduces_the_reference_report():
    """data/script/romabundledemo: the reference's mid-size benchmark (7.28 s / 5 iterations in
    MATLAB) - 60 images, 26321 object points, 90561 image points, n = 79321, self-calibration of cc, pp, K1,
    K2, no control points, datum by `seteoest(s,'depend',1)`, i.e. the datum and problem shape of the
    north-star workload.  Forward intersection from the prior EO, GNA: 5 iterations, 8641.39 -> 185.94,
    sigma0 0.582769, and the whole 800-line result file - every camera station with deviations and
    correlations, coverage, 26321-point statistics - exactly; the calibrated camera agrees with the
    18-digit values of result/EOS5DMarkII.xml to 1e-13.  The posterior covariances of this size come from the
    oracle's block path (`oracle/bundle.py:_prepare_blocks`, checked against the dense factor below)."""
    from oracle.photogrammetry import forwintersect
    from oracle.dbatstruct import seteoest_depend
    from oracle.bundle import bundle as obundle, bundle_cov as ocov
    from dbat_b200.report import bundle_result_file
    s = roma_struct()
    s, ids, _ = forwintersect(s, 'all', True)
    assert len(ids) == 26321
    seteoest_depend(s, 0)
    s, ok, it, s0, E = obundle(s, 'gna')
    assert ok and it == 5 and (E.numParams, E.numObs, E.redundancy) == (79321, 181122, 101801)
    assert abs(s0 - 0.582769) < 6e-7 and abs(E.res[0] - 8641.39) < 6e-3 and abs(E.res[-1] - 185.94) < 6e-4
    io = s.IO.val[:, 0]
    got = dict(cc=io[0], px=io[1], py=-io[2], K1=-io[5], K2=-io[6])
    for k, v in ROMA_CAMERA.items():
        assert abs(got[k] - v) < 1e-13 * abs(v) + 1e-20, k
    s, lines = bundle_result_file(s, E, None, cov=ocov)
    assert report_diff(lines, os.path.join(GOLD, 'romabundledemo', 'result', 'report.txt')) == []


def test_oracle_block_covariance_path_equals_the_dense_factor():
    """oracle/bundle.py: `bundle_cov(..., blocks=True)` (what large problems use) against the literal
    dense restatement of bundle_cov.m, on a self-calibrating project, on one with single coordinates of some
    points held fixed, and on one with prior OP observations."""
    from oracle.bundle import bundle as obundle, bundle_cov as ocov

    def partly_fixed_points():
        from oracle.loaders import camcal_pm_struct
        from oracle.photogrammetry import resect, forwintersect
        G = os.path.join(GOLD, 'camcalpm')
        s = camcal_pm_struct(os.path.join(G, 'camcal-pmexport.txt'), os.path.join(G, 'camcal-fixed.txt'))
        cpId = np.asarray(s.OP.id)[s.prior.OP.isCtrl]
        s = forwintersect(resect(s, 'all', cpId, 1, 0, cpId)[0], 'all', True)[0]
        s.bundle.est.OP[2, 5:9] = False                        # z of four points fixed: 2 x 2 point blocks
        s.bundle.est.OP[0:2, 20] = False                       # x, y of one point fixed: a 1 x 1 block
        return obundle(s, 'gna')

    for run in (lambda: camcal_pm_run('camcal-pmexport.txt'), partly_fixed_points,
                lambda: prague_run(os.path.join(GOLD, 'prague2016sxb'), 'w-op1', 'ctrlpts-weighted.txt')):
        s, ok, it, s0, E = run()
        assert ok
        dense = {w: np.asarray(ocov(s, E, w)) for w in ('CIOF', 'CEOF', 'CIO', 'CEO', 'COP')}
        E.final.factorized = None
        for w, D in dense.items():
            B = ocov(s, E, w, blocks=True)
            B = B.toarray() if hasattr(B, 'toarray') else B
            assert np.abs(B - D).max() <= 1e-9 * max(np.abs(D).max(), 1e-300), w


@pytest.mark.gpu
def test_roma_result_file_from_the_device():
    """The roma script project on the device: forward intersection, dependent datum, GNA self-calibration
    (n = 79321, reduced camera system of order 358), covariances and result file.  5 iterations and the
    reference's report to the printed digits (1e-5 relative, one unit of the last digit); the calibrated
    camera to 1e-9 of the reference's 18-digit values (the north-star tolerance)."""
    import dbat_b200
    from dbat_b200.report import bundle_result_file
    s = roma_struct()
    s, ids, _ = dbat_b200.forwintersect(s, 'all', True)
    dbat_b200.seteoest_depend(s, 0)
    s, ok, it, s0, E = dbat_b200.bundle(s, 'gna')
    assert ok and it == 5 and (E.numParams, E.numObs) == (79321, 181122)
    io = s.IO.val[:, 0]
    got = dict(cc=io[0], px=io[1], py=-io[2], K1=-io[5], K2=-io[6])
    for k, v in ROMA_CAMERA.items():
        assert abs(got[k] - v) < 1e-9 * abs(v), k
    s, lines = bundle_result_file(s, E)
    assert report_diff(lines, os.path.join(GOLD, 'romabundledemo', 'result', 'report.txt'), rtol=1e-5) == []


@pytest.mark.gpu
def test_failed_run_result_file_from_the_device():
    """camcaldemo_missing_obs.m on the device: code -4, and the post-mortem on the device-exported Jacobian -
    "Structural rank: 417 (deficiency: 6)" with the six suspected parameters - then the result file of the
    failed run (NaN deviations, start values) against the reference's."""
    import dbat_b200
    from oracle.loaders import camcal_pm_struct
    G = os.path.join(GOLD, 'camcalpm')
    s = camcal_pm_struct(os.path.join(G, 'camcal-pmexport-missing-obs.txt'), os.path.join(G, 'camcal-fixed.txt'))
    s.proj.x0desc = 'Camera calibration from EXIF value'
    cpId = np.asarray(s.OP.id)[s.prior.OP.isCtrl]
    s1, _, fail = dbat_b200.resect(s, 'all', cpId, 1, 0, cpId)
    s2, _, _ = dbat_b200.forwintersect(s1, 'all', True)
    s3, ok, it, s0, E = dbat_b200.bundle(s2, 'gna')
    assert E.code == -4
    w = E.weakness.structural
    assert not hasattr(w, 'error'), w.error
    assert w.rank == 417 and w.deficiency == 6
    assert w.suspectedParams == ['OX-12/13', 'OY-12/13', 'OZ-12/13', 'OX-59/60', 'OY-59/60', 'OZ-59/60']
    s3, lines = dbat_b200.bundle_result_file(s3, E)
    assert report_diff(lines, os.path.join(G, 'camcal-dbatreport-missing-obs.txt'), rtol=1e-4) == []


@pytest.mark.gpu
@pytest.mark.parametrize('damping,iters', [('lm', 7), ('lmp', 5)])
def test_roma_other_dampings_on_the_device(damping, iters):
    """LM and LMP on the roma project: the oracle needs 7 and 5 iterations and lands on GNA's minimum
    (sigma0 0.582769, last error 185.9397; run here on CPU and recorded in DESIGN.md §2); the device must
    take the same number of iterations to the same point."""
    import dbat_b200
    s = roma_struct()
    s, _, _ = dbat_b200.forwintersect(s, 'all', True)
    dbat_b200.seteoest_depend(s, 0)
    s, ok, it, s0, E = dbat_b200.bundle(s, damping)
    assert ok and it == iters
    assert abs(s0 - 0.582769) < 6e-7 and abs(E.res[-1] - 185.9397) < 6e-4
    io = s.IO.val[:, 0]
    got = dict(cc=io[0], px=io[1], py=-io[2], K1=-io[5], K2=-io[6])
    for k, v in ROMA_CAMERA.items():
        assert abs(got[k] - v) < 1e-7 * abs(v), k
