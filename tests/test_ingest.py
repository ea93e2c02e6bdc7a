"""dbat_b200/ingest.py (loadpm, prob2dbatstruct, loadcpt / matchcpt / setcpt, setcamvals, setcamest,
seteoest, cleareo, clearop) against the oracle's project loaders and, end to end, against the reference's
result files: the demo scripts below are the reference's demos, call for call."""
import copy
import os

import numpy as np
import pytest

from dbat_b200 import ingest
from test_report_golden import GOLD, report_diff

CAMCAL = os.path.join(GOLD, 'camcalpm')


def camcaldemo(pm, model=3):
    """code/demo/camcaldemo.m:40-98 (camcaldemo_allmodels.m for the other models)."""
    prob = ingest.loadpm(os.path.join(CAMCAL, pm))
    s0 = ingest.prob2dbatstruct(prob)
    if not s0.prior.OP.isCtrl.any():
        s0.prior.OP.isCtrl = np.asarray(s0.OP.id) > 1000
    pts = ingest.loadcpt(os.path.join(CAMCAL, 'camcal-fixed.txt'))
    i, j = ingest.matchcpt(s0, pts)
    s0 = ingest.setcpt(s0, pts, i, j)
    s0.IO.model.distModel[:] = model
    s0 = ingest.setcamvals(s0, 'default', 7.3)
    s0 = ingest.setcamest(s0, 'all', 'not', 'sk')
    s0 = ingest.seteoest(s0, 'all')
    s0 = ingest.cleareo(s0)
    return ingest.clearop(s0)


def assert_same_struct(a, b):
    for path in ('IO.val', 'EO.val', 'OP.val', 'IP.val', 'IP.std', 'IP.img', 'IP.op', 'IP.sigmas', 'OP.id',
                 'IO.sensor.pxSize', 'IO.sensor.imSize', 'IO.model.distModel', 'bundle.est.IO', 'bundle.est.EO',
                 'bundle.est.OP', 'prior.OP.use', 'prior.OP.isCtrl', 'prior.OP.isCheck', 'prior.EO.use', 'prior.IO.use'):
        x, y = a, b
        for k in path.split('.'):
            x, y = getattr(x, k), getattr(y, k)
        if path == 'OP.val':        # the demo's mean control-point offset is summed in a different order
            np.testing.assert_allclose(np.asarray(x), np.asarray(y), rtol=1e-12, atol=1e-9, err_msg=path)
        else:
            np.testing.assert_array_equal(np.asarray(x), np.asarray(y), err_msg=path)
    ctrl = a.prior.OP.isCtrl
    np.testing.assert_allclose(a.prior.OP.val[:, ctrl], b.prior.OP.val[:, ctrl], rtol=1e-12, atol=1e-9)
    np.testing.assert_array_equal(a.prior.OP.std[:, ctrl], b.prior.OP.std[:, ctrl])
    assert list(a.OP.label) == list(b.OP.label) and list(a.EO.name) == list(b.EO.name)
    assert a.proj.title == b.proj.title and (a.IO.model.nK, a.IO.model.nP) == (b.IO.model.nK, b.IO.model.nP)


@pytest.mark.parametrize('pm', ['camcal-pmexport.txt', 'camcal-pmexport5.txt', 'camcal-pmexport-1ray.txt',
                                'camcal-pmexport-missing-obs.txt'])
@pytest.mark.parametrize('model', [3, 1, -1])
def test_camcaldemo_ingest_equals_the_oracle_loader(pm, model):
    from oracle.loaders import camcal_pm_struct
    s = camcaldemo(pm, model)
    o = camcal_pm_struct(os.path.join(CAMCAL, pm), os.path.join(CAMCAL, 'camcal-fixed.txt'), model=model)
    assert_same_struct(s, o)


def prague2016_pm(root, stub, cps, orient='no'):
    """code/demo/prague2016_pm.m:140-195: loaded (fixed) camera, control points from the reference file
    moved by the mean offset to PhotoModeler's frame, EO and OP cleared."""
    prob = ingest.loadpm(os.path.join(root, 'pmexports', '%s-%s-orient-pmexport.txt' % (stub, orient)))
    s0 = ingest.prob2dbatstruct(prob)
    s0 = ingest.setcamvals(s0, 'loaded')
    s0 = ingest.setcamest(s0, 'not', 'all')
    ctrlPts = ingest.loadcpt(os.path.join(root, 'ref', 'ctrlpts-%s.txt' % cps))
    assert np.all(np.isin(prob.ctrlPts[:, 0], ctrlPts.id))
    _, ia, ib = np.intersect1d(prob.ctrlPts[:, 0], ctrlPts.id, return_indices=True)
    ctrlPts.pos = ctrlPts.pos + np.mean(prob.ctrlPts[ia, 1:4].T - ctrlPts.pos[:, ib], axis=1, keepdims=True)
    i, j = ingest.matchcpt(s0, ctrlPts, 'id')
    s0 = ingest.setcpt(s0, ctrlPts, i, j)
    s0 = ingest.cleareo(s0)
    return ingest.clearop(s0)


PRAGUE_CASES = [('prague2016cam', 'fixed', 'fixed'), ('prague2016cam', 'weighted', 'weighted'),
                ('prague2016sxb', 'f-op0', 'fixed'), ('prague2016sxb', 'w-op0', 'weighted'),
                ('prague2016sxb', 'w-op1', 'weighted'), ('prague2016sxb', 'wsmart', 'weighted')]


@pytest.mark.parametrize('project,stub,cps', PRAGUE_CASES)
def test_prague2016_ingest_equals_the_oracle_loader(project, stub, cps):
    from oracle.loaders import prague_cam_struct
    root = os.path.join(GOLD, project)
    s = prague2016_pm(root, stub, cps)
    o = prague_cam_struct(root, stub, 'ctrlpts-%s.txt' % cps)
    o.EO.val[:] = np.nan
    o.OP.val[:, ~o.prior.OP.isCtrl] = np.nan
    assert_same_struct(s, o)


def _solve_and_report(s):
    """resect -> forwintersect -> bundle (oracle, CPU) -> result file, as every demo continues."""
    from oracle.photogrammetry import resect, forwintersect
    from oracle.bundle import bundle as obundle, bundle_cov as ocov
    from dbat_b200.report import bundle_result_file
    cpId = np.asarray(s.OP.id)[s.prior.OP.isCtrl]
    s1, _, fail = resect(s, 'all', cpId, 1, 0, cpId)
    assert not fail
    s2, _, _ = forwintersect(s1, 'all', True)
    s3, ok, it, s0, E = obundle(copy.deepcopy(s2), 'gna')
    assert ok
    return bundle_result_file(s3, E, None, cov=ocov)[1]


def test_camcaldemo_from_file_to_result_file():
    """camcaldemo.m with this package's ingest in front of the solver: the reference's 610-line report."""
    s = camcaldemo('camcal-pmexport.txt')
    s.proj.x0desc = 'Camera calibration from EXIF value'
    assert report_diff(_solve_and_report(s), os.path.join(GOLD, 'dbatexports', 'camcal-dbatreport.txt')) == []


@pytest.mark.parametrize('project,stub,cps', [PRAGUE_CASES[1], PRAGUE_CASES[4], PRAGUE_CASES[5]])
def test_prague2016_from_file_to_result_file(project, stub, cps):
    root = os.path.join(GOLD, project)
    lines = _solve_and_report(prague2016_pm(root, stub, cps))
    assert report_diff(lines, os.path.join(root, 'dbatexports', '%s-no-orient-dbatreport.txt' % stub)) == []


def test_setters_follow_the_reference_argument_rules():
    s = camcaldemo('camcal-pmexport5.txt')
    s = ingest.setcamest(s, 'not', 'all')
    assert not s.bundle.est.IO.any()
    s = ingest.setcamest(s, 'cc', 'pp', 'K2')                       # K2 brings K1 (setcamest.m:84-86)
    assert list(np.flatnonzero(s.bundle.est.IO[:, 0])) == [0, 1, 2, 5, 6]
    s = ingest.setcamest(s, 'P1')                                   # P1 brings P2 (setcamest.m:95-98)
    assert s.bundle.est.IO[8:10].all()
    s = ingest.setcamest(s, 'not', 'K2')                            # not K2 takes K2 and K3 (:87-88)
    assert list(np.flatnonzero(s.bundle.est.IO[:, 0])) == [0, 1, 2, 5, 8, 9]
    s.IO.model.distModel[:] = 2
    with pytest.raises(ValueError):
        ingest.setcamest(s, 'as')                                   # no affine terms below model 3
    s = ingest.setcamest(s, 'all')
    assert not s.bundle.est.IO[3:5].any() and s.bundle.est.IO[[0, 1, 2, 5, 6, 7, 8, 9]].all()
    s = ingest.setcamvals(s, 'default', 7.3, 'K1', 1e-3, 'pp', [3.5, -2.5])
    assert s.IO.val[0, 0] == 7.3 and s.IO.val[5, 2] == 1e-3 and list(s.IO.val[1:3, 4]) == [3.5, -2.5]
    s = ingest.seteoest(s, 'none')
    assert not s.bundle.est.EO.any()
    s = ingest.seteoest(s, [1, 2], 'pos')
    assert s.bundle.est.EO[0:3, 1:3].all() and s.bundle.est.EO.sum() == 6
    s.EO.val[:] = np.arange(30.0).reshape(6, 5) * [[1], [2], [-1], [1], [1], [1]]
    s = ingest.seteoest(s, 'depend', 1)
    assert not s.bundle.est.EO[:, 0].any() and s.bundle.est.EO.sum() == 30 - 7 and not s.bundle.est.EO[1, 4]
    s = ingest.seteoest(s, 'depend', 2, 'z')                        # base camera 2, longest baseline along z
    assert not s.bundle.est.EO[:, 1].any() and s.bundle.est.EO.sum() == 30 - 7 and not s.bundle.est.EO[2, 0]
    with pytest.raises(ValueError):
        ingest.seteoest(s, 'depend', 1, 'w')
    with pytest.raises(NotImplementedError):
        ingest.prob2dbatstruct(ingest.loadpm(os.path.join(CAMCAL, 'camcal-pmexport5.txt')), True)


@pytest.mark.parametrize('usePriorEO', [True, False])
def test_sxb_prior_eo_demo_from_file_to_result_file(usePriorEO):
    """code/demo/sxb_prior_eo.m:30-75 call for call (prior observations of four camera positions matched
    by image label), then resect / forwintersect / bundle / result file against the reference's report."""
    root = os.path.join(GOLD, 'prague2016sxb')
    prob = ingest.loadpm(os.path.join(root, 'pmexports', 'wsmart-with-orient-pmexport.txt'))
    s0 = ingest.prob2dbatstruct(prob)
    s0 = ingest.setcamvals(s0, 'loaded')
    s0 = ingest.setcamest(s0, 'not', 'all')
    ctrlPts = ingest.loadcpt(os.path.join(root, 'ref', 'ctrlpts-weighted.txt'))
    i, j = ingest.matchcpt(s0, ctrlPts, 'id')
    s0 = ingest.setcpt(s0, ctrlPts, i, j)
    if usePriorEO:
        EOtbl = ingest.legacyloadeotable(os.path.join(root, 'ref', 'fake-camera-positions.txt'), (False, True))
        i, j = ingest.matcheo(s0, EOtbl)
        assert len(i) == 4
        s0 = ingest.setprioreo(s0, EOtbl, i, j)
    s0 = ingest.cleareo(s0)
    s0 = ingest.clearop(s0)
    assert np.isnan(s0.EO.val).sum() == (30 - 12 if usePriorEO else 30)
    lines = _solve_and_report(s0)
    rep = os.path.join(root, 'dbatexports', 'sxb-%sprior-eo-dbatreport.txt' % ('' if usePriorEO else 'no-'))
    assert report_diff(lines, rep, first_error_rtol=2e-5) == []


def test_format_string_table_readers():
    """loadimagepts / loadctrlpts / loadimagetable / loadeotable on the script projects' tables."""
    sxb = os.path.join(GOLD, 'sxb')
    pts = ingest.loadimagepts(os.path.join(sxb, 'measurements', 'markpts.txt'), 'id,im,x,y')
    assert pts.pos.shape == (2, 47) and (pts.id[0], pts.im[0]) == (317, 1) and np.isnan(pts.std).all()
    assert pts.pos[0, 0] == 5007.6667 and pts.pos[1, 0] == 7275.6667
    cal = ingest.loadimagepts(os.path.join(GOLD, 'camcaldemo', 'measurements', 'markpts.txt'), 'im,id,x,y,sxy')
    assert np.all(cal.std == 0.1) and cal.pos.shape[1] == 2074
    cp = ingest.loadctrlpts(os.path.join(sxb, 'reference', 'sxb-control.txt'), 'id,label,x,y,z,sx,sy,sz')
    assert len(cp.id) == 16 and cp.name[0] == 'B2.16' and list(cp.std[:, 0]) == [0.02, 0.02, 0.04]
    fx = ingest.loadctrlpts(os.path.join(GOLD, 'camcaldemo', 'reference', 'camcal-fixed.txt'), 'id,label,x,y,z')
    assert np.all(fx.std == 0) and list(fx.pos[:, 1]) == [1.0, 1.0, 0.0]
    ims = ingest.loadimagetable(os.path.join(sxb, 'images', 'images.txt'), 'id,path')
    assert list(ims.id) == [1, 2, 3, 4, 5] and ims.path[0].endswith('8811.jpg') and np.all(ims.cam == 1)
    eo = ingest.loadeotable(os.path.join(GOLD, 'romabundledemo', 'prior', 'initial_eo.txt'), 'id,x,y,z,omega,phi,kappa')
    assert eo.pos.shape == (3, 60) and list(eo.ang[:, 0]) == [39.43, 7.46, 99.59] and np.isnan(eo.std).all()
    with pytest.raises(ValueError):
        ingest.loadimagepts(os.path.join(sxb, 'measurements', 'markpts.txt'), 'id,im,x')


@pytest.mark.parametrize('stub', ['fixed', 'weighted'])
def test_external_verification_against_photomodeler(stub):
    """prague2016_pm.m:255-300, the reference's cross-software check: the adjusted object points and their
    posterior standard deviations against PhotoModeler's own 3-D point table for the same project
    (`pmexports/<stub>-no-orient-3dpts.txt`, printed to 1e-6 m).  They agree to that printing resolution -
    positions (after undoing the demo's mean control-point offset) to 1.5e-6 m, standard deviations to 1e-6 m."""
    from oracle.photogrammetry import resect, forwintersect
    from oracle.bundle import bundle as obundle, bundle_cov as ocov
    from dbat_b200.report import bundle_result_file
    root = os.path.join(GOLD, 'prague2016cam')
    s0 = prague2016_pm(root, stub, stub)
    prob = ingest.loadpm(os.path.join(root, 'pmexports', '%s-no-orient-pmexport.txt' % stub))
    ctrlPts = ingest.loadcpt(os.path.join(root, 'ref', 'ctrlpts-%s.txt' % stub))
    _, ia, ib = np.intersect1d(prob.ctrlPts[:, 0], ctrlPts.id, return_indices=True)
    meanOffset = np.mean(prob.ctrlPts[ia, 1:4].T - ctrlPts.pos[:, ib], axis=1, keepdims=True)
    cpId = np.asarray(s0.OP.id)[s0.prior.OP.isCtrl]
    s1, _, fail = resect(s0, 'all', cpId, 1, 0, cpId)
    s2, _, _ = forwintersect(s1, 'all', True)
    s3, ok, it, sig0, E = obundle(copy.deepcopy(s2), 'gna')
    s3, _ = bundle_result_file(s3, E, None, cov=ocov)
    pts3d = ingest.loadpm3dtbl(os.path.join(root, 'pmexports', '%s-no-orient-3dpts.txt' % stub))
    _, i, j = np.intersect1d(pts3d.id, s3.OP.id, return_indices=True)
    assert len(i) == len(pts3d.id) == 100
    assert np.abs(s3.OP.val[:, j] - meanOffset - pts3d.pos[:, i]).max() < 1.5e-6
    assert np.abs(s3.post.std.OP[:, j] - pts3d.std[:, i]).max() < 1e-6
    # camera stations and total error against PhotoModeler's status report (values printed to 1e-6, deviations
    # to 1e-3): its own adjustment stopped a little short of the minimum, as the reference's paper reports
    pm = ingest.loadpmreport(os.path.join(root, 'pmexports', '%s-no-orient-pmreport.txt' % stub))
    assert pm.EO.shape == (6, 21)
    assert np.abs(s3.EO.val[0:3] - pm.EO[0:3]).max() < 5e-5 and np.abs(s3.EO.val[3:6] - pm.EO[3:6]).max() < np.deg2rad(1e-3)
    assert np.abs(s3.post.std.EO[0:3] - pm.EOstd[0:3]).max() < 6e-4
    assert np.abs(s3.post.std.EO[3:6] - pm.EOstd[3:6]).max() < np.deg2rad(6e-4)
    assert abs(sig0 / pm.lastError - 1) < 2e-3
    assert [len(v) for v in pts3d.vis][:2] == [21, 21] and max(len(v) for v in pts3d.vis) == 21


def ps_postproc(psFile):
    """code/demo/ps_postproc.m:30-105 (no point filtering): PhotoScan project -> PhotoModeler-shaped prob ->
    struct, forward lens model, camera fixed unless PhotoScan adjusted it, EO / OP start values as
    PhotoScan estimated them."""
    psz = ingest.loadpsz(psFile)
    prob = ingest.ps2pmstruct(psz)
    s0 = ingest.prob2dbatstruct(prob)
    s0.IO.model.distModel[:] = -1
    assert not psz.camera.isAdjusted
    return s0, psz


def test_photoscan_project_from_archive_to_result_file():
    """ps_postproc('') on data/prague2016/sxb/psprojects/sxb.psz (zip of doc.xml + binary PLY tables): 5
    cameras placed by the chunk's similarity transform, 16 markers as weighted control points, 1166 tie
    points with 4018 projections at 1 px and 48 marker measurements at 0.1 px.  Starting from PhotoScan's
    own estimates the bundle needs 3 iterations (93.8342 -> 48.1954, sigma0 0.710294) and the result file
    is the reference's `sxb-dbatreport.txt`, iteration count and first error included - which pins the
    transform chain of loadpsz (camera-to-chunk, axis flip, chunk-to-world) as well."""
    from oracle.bundle import bundle as obundle, bundle_cov as ocov
    from dbat_b200.report import bundle_result_file
    root = os.path.join(GOLD, 'prague2016sxb')
    s0, psz = ps_postproc(os.path.join(root, 'psprojects', 'sxb.psz'))
    assert s0.IP.val.shape[1] == 4066 and list(s0.IP.sigmas) == [0.1, 1.0]
    assert s0.prior.OP.isCtrl.sum() == 16 and s0.prior.OP.use.sum() == 48 and s0.OP.val.shape[1] == 1182
    assert np.isfinite(s0.EO.val).all() and np.isfinite(s0.OP.val).all()
    s, ok, it, sig0, E = obundle(copy.deepcopy(s0), 'gna', 20)
    assert ok and it == 3 and (E.numParams, E.numObs) == (3576, 8180)
    s, lines = bundle_result_file(s, E, None, cov=ocov)
    assert report_diff(lines, os.path.join(root, 'psprojects', 'sxb-dbatreport.txt')) == []


@pytest.mark.gpu
def test_photoscan_project_on_the_device():
    """The same on the device: forward (computer vision) lens model -1, fixed camera, weighted control points."""
    import dbat_b200
    root = os.path.join(GOLD, 'prague2016sxb')
    s0, psz = ps_postproc(os.path.join(root, 'psprojects', 'sxb.psz'))
    s, ok, it, sig0, E = dbat_b200.bundle(s0, 'gna', 20)
    assert ok and it == 3
    s, lines = dbat_b200.bundle_result_file(s, E)
    assert report_diff(lines, os.path.join(root, 'psprojects', 'sxb-dbatreport.txt'), rtol=1e-5) == []


def test_prague2016_ps_demo_from_archive_to_result_file():
    """code/demo/prague2016_ps.m ('s5'): the PhotoScan project with the control points of
    `ref/ctrlpts-weighted-raw.txt` (matched by PhotoScan's marker ids), EO / OP cleared and recomputed by
    resection and intersection, legacy backward model: `dbatexports/sxb-dbatreport.txt`, exact but for the
    start-value dependent first error (resection at 1e6 m coordinates: 2e-4)."""
    root = os.path.join(GOLD, 'prague2016sxb')
    psz = ingest.loadpsz(os.path.join(root, 'psprojects', 'sxb.psz'))
    ctrlPts = ingest.loadcpt(os.path.join(root, 'ref', 'ctrlpts-weighted-raw.txt'))
    s0 = ingest.prob2dbatstruct(ingest.ps2pmstruct(psz))
    s0 = ingest.setcamvals(s0, 'loaded')
    s0 = ingest.setcamest(s0, 'not', 'all')
    i, j = ingest.matchcpt(s0, ctrlPts)
    assert len(i) == 15
    s0 = ingest.setcpt(s0, ctrlPts, i, j)
    s0 = ingest.clearop(ingest.cleareo(s0))
    lines = _solve_and_report(s0)
    rep = os.path.join(root, 'dbatexports', 'sxb-dbatreport.txt')
    exact = report_diff(lines, rep)
    assert len(exact) == 1 and 'First error' in exact[0][1]
    assert report_diff(lines, rep, first_error_rtol=5e-4) == []


def test_photoscan_project_statistics_file():
    """ps_postproc.m:70-76 -> writestats.m: the 1101-line pre-bundle statistics of the PhotoScan project
    (image / control point / object point ray counts and ray angles, MATLAB-style histograms, worst cases)
    against `psprojects/sxb-psstats-prefilt.txt`; only the path and time stamp lines differ."""
    from dbat_b200.report import writestats
    root = os.path.join(GOLD, 'prague2016sxb')
    s0, psz = ps_postproc(os.path.join(root, 'psprojects', 'sxb.psz'))
    s0, lines = writestats(s0, None, 'Initial, unfiltered statitistics')
    gold = [l.rstrip('\n') for l in open(os.path.join(root, 'psprojects', 'sxb-psstats-prefilt.txt'))]
    assert len(lines) == len(gold) == 1101
    bad = [(n + 1, a, b) for n, (a, b) in enumerate(zip(gold, lines)) if a != b]
    assert [n for n, _, _ in bad] == [3, 5, 14, 15, 16, 17, 18]          # project file, time stamp, image paths
    assert all(a.split(', ')[:2] == b.split(', ')[:2] for n, a, b in bad if n >= 14)
    assert s0.camRayAng.shape == (5,) and s0.rayAng.shape == (1182,)


def test_point_filter_of_the_photoscan_demo():
    """loadplotpsz.m:58-90 (ps_postproc's minRays / minAngle arguments): filtered points disappear with their
    measurements, control points stay whatever their ray count."""
    root = os.path.join(GOLD, 'prague2016sxb')
    psz = ingest.loadpsz(os.path.join(root, 'psprojects', 'sxb.psz'))
    prob = ingest.ps2pmstruct(psz)
    s0 = ingest.prob2dbatstruct(prob)
    rays = np.bincount(s0.IP.op, minlength=s0.OP.val.shape[1])
    prob4, gone = ingest.filterprob(prob, s0, minRays=4)
    s4 = ingest.prob2dbatstruct(prob4)
    assert len(gone) == np.count_nonzero((rays < 4) & ~s0.prior.OP.isCtrl) > 0
    assert s4.OP.val.shape[1] == s0.OP.val.shape[1] - len(gone) and s4.prior.OP.isCtrl.sum() == 16
    r4 = np.bincount(s4.IP.op, minlength=s4.OP.val.shape[1])
    assert r4[~s4.prior.OP.isCtrl].min() >= 4 and r4[s4.prior.OP.isCtrl].min() == 1
    assert s4.IP.val.shape[1] == s0.IP.val.shape[1] - rays[np.isin(s0.OP.id, gone)].sum()
    from dbat_b200.report import angles
    probA, goneA = ingest.filterprob(prob, s0, minAngle=15.0)
    sA = ingest.prob2dbatstruct(probA)
    assert len(goneA) > 0 and np.rad2deg(angles(sA))[~sA.prior.OP.isCtrl].min() >= 15.0
    same, none = ingest.filterprob(prob, s0)
    assert len(none) == 0 and same.objPts.shape == prob.objPts.shape
