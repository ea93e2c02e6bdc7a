"""Pin the CPU oracle against the reference's own golden result files (camcal XML project,
data/script/camcaldemo/result/* copied into tests/golden/camcaldemo by make_golden.py)."""
import os
import re

import numpy as np
import pytest

from camcal_fixture import GOLD, camcal_struct, golden_camera_xml, golden_eo
from oracle import loaders
from oracle.bundle import bundle, bundle_cov
from oracle.cameramodel import brown_euler_cam4
from oracle.dbatstruct import buildserialindices, serialize


@pytest.fixture(scope='module')
def gna_run():
    s = camcal_struct('default', seed=1)
    s, ok, iters, s0, E = bundle(s, 'gna')
    assert ok
    return s, iters, s0, E


def test_report_numbers(gna_run):
    """report.txt: 423 params (9 IO, 126 EO, 288 OP), 4148 observations, sigma0 1.6148,
    last error 98.556, redundancy 3725."""
    s, iters, s0, E = gna_run
    rep = open(os.path.join(GOLD, 'result', 'report.txt')).read()
    assert E.numParams == int(re.search(r'Number of params:\s+(\d+)', rep).group(1)) == 423
    assert E.numObs == int(re.search(r'Number of observations:\s+(\d+)', rep).group(1)) == 4148
    assert E.redundancy == int(re.search(r'Redundancy\s+(\d+)', rep).group(1))
    assert len(s.bundle.serial.IO.dest) == 9 and len(s.bundle.serial.EO.dest) == 126
    sig = float(re.search(r'Sigma0:\s+([\d.]+)', rep).group(1))
    assert abs(s0 - sig) < 5e-5
    last = float(re.search(r'Last error:\s+([\d.]+)', rep).group(1))
    assert abs(E.res[-1] - last) < 5e-4


def test_io_18_digits(gna_run):
    """result/c4040z.xml: cc, pp, K1-3, P1-2, aspect printed with 18 digits.  The golden run
    itself stopped at convTol=1e-6, so agreement is limited by its own convergence error."""
    s, _, _, _ = gna_run
    gold = loaders.load_camcal_script(GOLD, golden_camera_xml()).IO.val[:, 0]
    est = s.IO.val[:, 0]
    np.testing.assert_allclose(est[[0, 1, 2]], gold[[0, 1, 2]], rtol=2e-8)       # cc, px, py
    np.testing.assert_allclose(est[3], gold[3], rtol=1e-5)                       # aspect
    np.testing.assert_allclose(est[5:10], gold[5:10], rtol=5e-6)                 # K1-3, P1-2


def test_eo_and_posterior_std(gna_run):
    """result/camera_stations.txt: EO values and posterior std devs (bundle_cov 'CEO')."""
    s, _, _, E = gna_run
    ids, EOg, stdg = golden_eo()
    np.testing.assert_allclose(s.EO.val, EOg, atol=2e-8)
    CEO = bundle_cov(s, E, 'CEO')
    sd = np.sqrt(np.diag(CEO)).reshape(6, -1, order='F')
    # the file prints every std dev multiplied by 180/pi (header: "Unit: degrees")
    np.testing.assert_allclose(sd * 180 / np.pi, stdg, rtol=1e-6)


def test_top_residuals(gna_run):
    """result/top_residuals.txt: 50 largest image residuals (pixels), 6 printed digits."""
    s, _, _, _ = gna_run
    rows = loaders.load_table(os.path.join(GOLD, 'result', 'top_residuals.txt'))
    idmap = {(int(s.OP.id[j]), int(s.EO.id[i])): k for k, (i, j) in enumerate(zip(s.IP.img, s.IP.op))}
    for r in rows:
        k = idmap[(int(r[0]), int(r[1]))]
        np.testing.assert_allclose(s.post.res.IP[:, k], [float(r[4]), float(r[5])], atol=2e-6)


def test_known_answer_residual_at_golden_parameters():
    """residual formula at the golden IO/EO reproduces the control-point rows
    of top_residuals.txt (independent of any optimiser)."""
    s = loaders.load_camcal_script(GOLD, golden_camera_xml())
    _, EOg, _ = golden_eo()
    s.EO.val = EOg
    s.OP.val[:, ~s.prior.OP.isCtrl] = 0.0
    s.bundle.est.IO[:] = False
    s.bundle.est.OP[:] = False
    buildserialindices(s)
    f, _ = brown_euler_cam4(serialize(s), s, False)
    res_px = f.reshape(2, -1, order='F') / s.IO.sensor.pxSize[:, s.IP.cam]
    rows = loaders.load_table(os.path.join(GOLD, 'result', 'top_residuals.txt'))
    idmap = {(int(s.OP.id[j]), int(s.EO.id[i])): k for k, (i, j) in enumerate(zip(s.IP.img, s.IP.op))}
    n = 0
    for r in rows:
        if int(r[0]) > 1000:
            k = idmap[(int(r[0]), int(r[1]))]
            np.testing.assert_allclose(res_px[:, k], [float(r[4]), float(r[5])], atol=2e-6)
            n += 1
    assert n >= 30


@pytest.mark.parametrize('damping', ['lm', 'lmp'])
def test_other_dampings_reach_the_same_minimum(gna_run, damping):
    s, _, s0, E = gna_run
    s2 = camcal_struct('default', seed=1)
    s2, ok, _, s02, E2 = bundle(s2, damping)
    assert ok
    np.testing.assert_allclose(s02, s0, rtol=1e-9)
    np.testing.assert_allclose(E2.x, E.x, rtol=1e-6, atol=1e-9)


def test_analytic_jacobian_matches_central_differences():
    """The reference's own unit-test method (private/full_self_test.m, jacapprox.m h=1e-6, thr 1e-8)
    for all four modular models."""
    from oracle.cameramodel import res_euler_brown
    rng = np.random.default_rng(5)
    m = 5
    Q = 3 + rng.random((3, m)); ang = rng.random(3) * np.pi / 6; q0 = rng.random(3)
    f = 1 + rng.random(); u = rng.random((2, m)); K = rng.random(4) * 0.1; P = rng.random(3) * 0.1
    sz = rng.random() / 10; u0 = rng.random(2); b = rng.random(2) * 0.1
    for model in range(4):
        v, d = res_euler_brown(model, Q, q0, ang, f, u, sz, u0, K, P, b, True)

        def num(fun, x0):
            x0 = np.array(x0, dtype=float)
            out = []
            for k in range(x0.size):
                e = np.zeros(x0.size); e[k] = 1e-6
                out.append((fun((x0.ravel() + e).reshape(x0.shape)) - fun((x0.ravel() - e).reshape(x0.shape))) / 2e-6)
            return np.array(out)            # (nparam, 2, m)
        F = lambda **kw: res_euler_brown(model, kw.get('Q', Q), kw.get('q0', q0), kw.get('ang', ang), kw.get('f', f),
                                         u, sz, kw.get('u0', u0), kw.get('K', K), kw.get('P', P), kw.get('b', b), False)[0]
        np.testing.assert_allclose(num(lambda x: F(q0=x), q0), np.transpose(d['dQ0'], (2, 1, 0)), atol=1e-7)
        np.testing.assert_allclose(num(lambda x: F(ang=x), ang), np.transpose(d['dA'], (2, 1, 0)), atol=1e-7)
        np.testing.assert_allclose(num(lambda x: F(f=x[0]), [f])[0], d['dF'].T, atol=1e-7)
        np.testing.assert_allclose(num(lambda x: F(u0=x), u0), np.transpose(d['dU0'], (2, 1, 0)), atol=1e-7)
        np.testing.assert_allclose(num(lambda x: F(K=x), K), np.transpose(d['dK'], (2, 1, 0)), atol=1e-7)
        np.testing.assert_allclose(num(lambda x: F(P=x), P), np.transpose(d['dP'], (2, 1, 0)), atol=1e-7)
        if model > 0:
            np.testing.assert_allclose(num(lambda x: F(b=x), b), np.transpose(d['dB'], (2, 1, 0)), atol=1e-7)
        nq = num(lambda x: F(Q=x.reshape(3, m, order='F')), Q.reshape(-1, order='F'))   # (3m,2,m)
        for j in range(m):
            np.testing.assert_allclose(nq[3 * j:3 * j + 3, :, j], d['dQ'][j].T, atol=1e-7)


@pytest.mark.parametrize('model,sigma0', [(-1, 1.62168), (1, 1.68901), (2, 1.68901), (3, 1.6148),
                                          (4, 1.61247), (5, 1.6148)])
def test_all_models_golden_sigma0(model, sigma0):
    """data/dbat/dbatexports/camcal-dbatreport-model{-1,1,2,3,4,5}.txt (camcaldemo_allmodels.m):
    sigma0 and parameter count of every distortion model, including the legacy code path."""
    rep = open(os.path.join(os.path.dirname(GOLD), 'dbatexports', 'camcal-dbatreport-model%d.txt' % model)).read()
    assert abs(float(re.search(r'Sigma0:\s+([\d.]+)', rep).group(1)) - sigma0) < 1e-9
    s = camcal_struct('default', seed=1)
    s.IO.model.distModel[:] = model
    if abs(model) < 3:
        s.bundle.est.IO[3:5, :] = False            # setcamest.m:46-58
    s, ok, iters, s0, E = bundle(s, 'gna')
    assert ok
    assert abs(s0 - sigma0) < 6e-6 * (10 if sigma0 == 1.6148 else 1)
    assert E.numParams == int(re.search(r'Number of params:\s+(\d+)', rep).group(1))


PRAGUE = os.path.join(os.path.dirname(GOLD), 'prague2016cam')


@pytest.mark.parametrize('stub,sigma0,nparams,nobs,last', [('weighted', 1.60984, 426, 4160, 98.3715),
                                                           ('fixed', 1.78095, 414, 4148, 108.827)])
def test_prague2016_cam_golden(stub, sigma0, nparams, nobs, last):
    """data/prague2016/cam/dbatexports/{weighted,fixed}-no-orient-dbatreport.txt: PhotoModeler export,
    legacy model 1, fixed camera, fixed / weighted control points (12 prior OP observations)."""
    s = loaders.prague_cam_struct(PRAGUE, stub)
    s, ok, iters, s0, E = bundle(s, 'gna')
    assert ok
    assert abs(s0 - sigma0) < 6e-6
    assert (E.numParams, E.numObs) == (nparams, nobs)
    assert E.redundancy == 3734
    assert abs(E.res[-1] - last) < 6e-4


def test_stpierre_oracle_converges():
    """hamburg2017 StPierre C5_reduced export (BASELINE config 3 data; no reference golden exists for
    this input): the oracle converges with every damping to the same sigma0 ~ 1 (a-priori mark point
    sigma of 1 px, prob2dbatstruct.m:367-373) - a plausibility pin of loader + model -1 + prior
    observations on real data."""
    import copy, os
    from oracle import loaders
    from oracle.bundle import bundle as obundle
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'stpierre')
    s = loaders.stpierre_struct(root)
    assert s.EO.val.shape[1] == 28 and s.OP.val.shape[1] == 2003 and s.IP.val.shape[1] == 4331
    vals = []
    for damping, n in (('gna', 5), ('lmp', 5), ('lm', 6)):
        s2, ok, iters, s0, E = obundle(copy.deepcopy(s), damping)
        assert ok and iters == n
        vals.append(s0)
    assert max(vals) - min(vals) < 1e-9 and abs(vals[0] - 1.0282960127) < 1e-8


def test_forwintersect_oracle_known_answer():
    """Restatement of forwintersect.m / pm_multiforwintersect.m / pm_forwintersect3.m: with the true
    IO/EO and noise-free image points of a seeded synthetic block (no affine term) the intersection
    returns the true object points (the generator projects with the modular model 3, the intersection
    removes the distortion with pm_lens1 - both sides of the same Brown polynomial), residuals are ~0,
    and a point left with one ray comes back as NaN."""
    import numpy as np
    from dbat_b200 import synth
    from oracle.photogrammetry import forwintersect
    saved = synth.IO_TRUE.copy()
    synth.IO_TRUE[3] = 0.0            # pm_multilenscorr1 has no affine term ("a: not implemented yet", :155-158)
    try:
        s, truth = synth.make_scene(21, 60, rays=6, seed=3, noise_px=0.0, build_indices=False)
    finally:
        synth.IO_TRUE[:] = saved
    s.IO.val[:] = truth['IO'][:, None]
    s.EO.val[:] = truth['EO']
    k = np.flatnonzero(s.IP.op == 5)
    keep = np.ones(len(s.IP.op), bool)
    keep[k[1:]] = False
    for f in ('val', 'std'):
        setattr(s.IP, f, getattr(s.IP, f)[:, keep])
    for f in ('img', 'op', 'cam'):
        setattr(s.IP, f, getattr(s.IP, f)[keep])
    s.OP.val[:] = np.nan
    s2, ids, res = forwintersect(s, 'all')
    m = np.arange(60) != 5
    assert np.isnan(s2.OP.val[:, 5]).all() and np.isnan(res[5])
    # model 3 applies the distortion on the measured side (backward), as pm_multilenscorr1 does
    assert np.abs(s2.OP.val[:, m] - truth['OP'][:, m]).max() < 1e-6
    assert np.nanmax(res) < 1e-7


def test_camcal_script_pipeline_matches_golden_report():
    """The whole script pipeline of data/script/camcaldemo/camcaldemo.xml - default camera values,
    spatial_resection (resect.m, 3-point Grunert from the 4 control points), forward_intersection
    (forwintersect.m), bundle_adjustment (GNA) - reproduces the reference's own report.txt:
    8 iterations, first error 28805.9, last error 98.556, sigma0 1.6148.  This pins the restatements of
    resect / pm_resect_3pt / largesttriangle / forwintersect on a reference golden: the first error is a
    function of the start values only."""
    import copy, os
    import numpy as np
    from oracle.loaders import load_camcal_script
    from oracle.photogrammetry import resect, forwintersect
    from oracle.bundle import bundle as obundle
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'camcaldemo')
    s = load_camcal_script(root)
    assert np.isnan(s.EO.val).all()
    cpId = np.asarray(s.OP.id)[s.prior.OP.isCtrl]
    s1, rms, fail = resect(s, 'all', cpId, 1, 0, cpId)
    assert not fail and np.isfinite(s1.EO.val).all() and rms.max() < 0.2      # normalised image units, default camera
    s2, ids, res = forwintersect(s1, 'all', True)
    assert np.isfinite(s2.OP.val).all() and len(ids) == 96
    s3, ok, iters, sigma0, E = obundle(copy.deepcopy(s2), 'gna')
    assert ok and iters == 8
    assert abs(E.res[0] - 28805.9) < 0.06 and abs(E.res[-1] - 98.556) < 6e-4
    assert abs(sigma0 - 1.6148) < 6e-5


def test_gauss_markov_oracle_reaches_golden_solution():
    """gauss_markov.m restated (undamped Gauss-Newton): from the start values of the camcal fixture it
    reaches the same estimates as the golden GNA run."""
    import copy
    import numpy as np
    from camcal_fixture import camcal_struct
    from oracle.cameramodel import brown_euler_cam4
    from oracle.dbatstruct import buildserialindices, buildweightmatrix, serialize
    from oracle.lsa import gauss_markov
    from oracle.bundle import bundle as obundle
    s = camcal_struct('default', seed=1)
    buildserialindices(s)
    x0, W = serialize(s), buildweightmatrix(s)
    x, code, n, final, T, rr = gauss_markov(lambda xx, j: brown_euler_cam4(xx, s, j), x0, W, 20, 1e-6, False, False)
    assert code == 0 and T.shape[1] == n + 1 and len(rr) == n + 1
    s2, ok, iters, s0, E = obundle(copy.deepcopy(s), 'gna')
    assert ok
    np.testing.assert_allclose(x, E.x, rtol=1e-6, atol=1e-8)
    assert abs(rr[-1] - 98.556) < 6e-4


@pytest.mark.parametrize('pm,iters,first,last,sigma0,nparams', [
    ('camcal-pmexport.txt', 9, 30873.9, 98.556, 1.6148, 423),
    ('camcal-pmexport5.txt', 6, 254.75, 18.4099, 2.80749, 57)])
def test_camcal_pm_demo_pipeline_matches_golden_reports(pm, iters, first, last, sigma0, nparams):
    """camcaldemo.m / camcaldemo2.m end to end (loadpm, prob2dbatstruct, default camera 7.3 mm, fixed
    control points, resect, forwintersect, bundle GNA) against data/dbat/dbatexports/camcal-dbatreport.txt
    and camcal-dbatreport5.txt: iteration count, first and last error, sigma0, number of parameters."""
    import copy, os
    import numpy as np
    from oracle.loaders import camcal_pm_struct
    from oracle.photogrammetry import resect, forwintersect
    from oracle.bundle import bundle as obundle
    G = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'camcalpm')
    s = camcal_pm_struct(os.path.join(G, pm), os.path.join(G, 'camcal-fixed.txt'))
    cpId = np.asarray(s.OP.id)[s.prior.OP.isCtrl]
    s1, _, fail = resect(s, 'all', cpId, 1, 0, cpId)
    assert not fail
    s2, _, _ = forwintersect(s1, 'all', True)
    s3, ok, it, s0, E = obundle(copy.deepcopy(s2), 'gna')
    assert ok and it == iters and E.numParams == nparams
    assert abs(E.res[0] - first) < 0.6 * 10 ** (np.floor(np.log10(first)) - 5)      # 6 printed digits
    assert abs(E.res[-1] - last) < 0.6 * 10 ** (np.floor(np.log10(last)) - 5)
    assert abs(s0 - sigma0) < 0.6 * 10 ** (np.floor(np.log10(sigma0)) - 5)


@pytest.mark.parametrize('pm,code,first,sigma0', [
    ('camcal-pmexport-missing-obs.txt', -4, 30118.6, 499.142),
    ('camcal-pmexport-1ray.txt', -4, None, None)])
def test_camcal_pm_failure_demos_match_golden_reports(pm, code, first, sigma0):
    """camcaldemo_missing_obs.m / camcaldemo_1ray.m: the reference's reports record failure code -4
    (structurally rank-deficient normal matrix) at iteration 0, with first error 30118.6 / sigma0 499.142
    for the missing-observations project and NaN for the single-ray one."""
    import copy, os
    import numpy as np
    from oracle.loaders import camcal_pm_struct
    from oracle.photogrammetry import resect, forwintersect
    from oracle.bundle import bundle as obundle
    G = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'camcalpm')
    s = camcal_pm_struct(os.path.join(G, pm), os.path.join(G, 'camcal-fixed.txt'))
    cpId = np.asarray(s.OP.id)[s.prior.OP.isCtrl]
    s1, _, fail = resect(s, 'all', cpId, 1, 0, cpId)
    s2, _, _ = forwintersect(s1, 'all', True)
    s3, ok, it, s0, E = obundle(copy.deepcopy(s2), 'gna')
    assert not ok and it == 0 and E.code == code and E.numParams == 423
    if first is None:
        assert np.isnan(s0) and np.isnan(E.res[0])
    else:
        assert abs(E.res[0] - first) < 0.06 and abs(s0 - sigma0) < 6e-4


def test_camcal_no_datum_demo_matches_golden_report():
    """camcaldemo_no_datum.m: start values as loaded, no control points (435 parameters, 7-dimensional
    gauge freedom).  The reference reports code -2 (singular normal matrix: MATLAB's rcond warning) at
    iteration 0, first error 15772.8, sigma0 258.848."""
    import copy, os
    from oracle.loaders import camcal_pm_struct
    from oracle.bundle import bundle as obundle
    G = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'camcalpm')
    s = camcal_pm_struct(os.path.join(G, 'camcal-pmexport.txt'), None, keep_loaded=True)
    s3, ok, it, s0, E = obundle(copy.deepcopy(s), 'gna')
    assert not ok and it == 0 and E.code == -2 and E.numParams == 435
    assert abs(E.res[0] - 15772.8) < 0.06 and abs(s0 - 258.848) < 6e-4
    # bundle.m:373-428 — report: "Numerical rank: 428 (deficiency: 7)" and seven null-space vectors
    # with eigenvalues ~1e-17.  A basis of a 7-dimensional eigenspace is not unique, so every vector
    # the report prints (its largest entries, 3 digits) must lie in the span of ours: the least-squares
    # fit over the printed entries leaves only print rounding and needs coefficients of norm <= 1.
    import re
    import numpy as np
    w = E.weakness.numerical
    assert E.weakness.structural is None and (w.rank, w.deficiency) == (428, 7)
    assert w.V.shape == (435, 7) and np.abs(w.d).max() < 1e-12 and abs(w.trace - 435) < 1e-4
    assert np.linalg.norm(E.final.scaled.J @ w.V) < 1e-10
    vecs = []
    for line in open(os.path.join(G, 'camcal-dbatreport-no-datum.txt')):
        if re.match(r'\s+Vector \d+', line):
            vecs.append([])
        m = re.match(r'\s+\(([A-Za-z0-9]+-\d+), (-?[\d.e-]+)\)', line)
        if m and vecs:
            vecs[-1].append((m.group(1), float(m.group(2))))
        if 'Problems related' in line:
            break
    assert len(vecs) == 7
    ix = {p: i for i, p in enumerate(E.paramTypes)}
    for v in vecs:
        K = [ix[p] for p, _ in v]
        g = np.array([x for _, x in v])
        c = np.linalg.lstsq(w.V[K], g, rcond=1e-2)[0]
        assert np.abs(w.V[K] @ c - g).max() < 1e-3 and np.linalg.norm(c) < 1.001
    for sp_ in w.suspectedParams:                    # IO parameters take no part in a gauge freedom
        assert all(p[:2] in ('EX', 'EY', 'EZ', 'om', 'ph', 'ka', 'OX', 'OY', 'OZ') for p in sp_.params)
    # the package's host-side analysis (Gram-matrix route) on the same Jacobian: same rank, same space
    from dbat_b200.bundle import numerical_weakness
    wp = numerical_weakness(E.final.scaled.J, E.paramTypes)
    assert (wp.rank, wp.deficiency) == (428, 7) and abs(wp.trace - w.trace) < 1e-9
    assert np.linalg.norm(wp.V - w.V @ (w.V.T @ wp.V)) < 1e-8
    assert len(wp.suspectedParams) == 7


@pytest.mark.parametrize('stub,sigma0,last,nparams', [('fixed', 1.78095, 108.827, 414), ('weighted', 1.60984, 98.3715, 426)])
def test_prague2016_cam_demo_pipeline_matches_golden_reports(stub, sigma0, last, nparams):
    """prague2016_pm('c1'|'c2') end to end as the demo runs it (`prague2016_pm.m:100-215`): fixed
    camera, control points fixed / weighted, EO and OP cleared, start values by resect + forwintersect,
    GNA.  The reference's reports (`{fixed,weighted}-{no,with}-orient-dbatreport.txt`, all four agree)
    give 3 iterations, first error 460.624 (a function of the start values only), last error, sigma0 and
    the parameter count."""
    import copy, os
    import numpy as np
    from oracle.loaders import prague_cam_struct
    from oracle.photogrammetry import resect, forwintersect
    from oracle.bundle import bundle as obundle
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'prague2016cam')
    s = prague_cam_struct(root, stub)
    s.EO.val[:] = np.nan                                   # cleareo / clearop (prague2016_pm.m:194-195)
    s.OP.val[:, ~s.prior.OP.isCtrl] = np.nan
    cpId = np.asarray(s.OP.id)[s.prior.OP.isCtrl]
    s1, _, fail = resect(s, 'all', cpId, 1, 0, cpId)
    assert not fail
    s2, _, _ = forwintersect(s1, 'all', True)
    s3, ok, it, s0, E = obundle(copy.deepcopy(s2), 'gna')
    assert ok and it == 3 and E.numParams == nparams
    assert abs(E.res[0] - 460.624) < 6e-4 and abs(E.res[-1] - last) < 6e-4 and abs(s0 - sigma0) < 6e-6


@pytest.mark.parametrize('stub,cps,first,last,sigma0,nparams', [
    ('f-op0', 'fixed', 10.4471, 8.33522, 1.0419, 30),
    ('w-op0', 'weighted', 10.4471, 7.87923, 0.984904, 78),
    ('w-op1', 'weighted', 10.6748, 8.01901, 0.965375, 81),
    ('wsmart', 'weighted', 60.1091, 38.2456, 1.07447, 1173)])
def test_prague2016_sxb_demo_pipelines_match_golden_reports(stub, cps, first, last, sigma0, nparams):
    """prague2016_pm('s1'..'s4') (`prague2016_pm.m`, StereoBox data: 5 images; fixed or weighted control
    points; s3 adds a check point, s4 1100 "smart" points whose zero export sigmas make prob2dbatstruct
    fall back to 1 px for every image point): resect + forwintersect start, GNA.  Reports
    `data/prague2016/sxb/dbatexports/<stub>-no-orient-dbatreport.txt`: 4 iterations, first / last error,
    sigma0, parameter count."""
    import copy, os
    import numpy as np
    from oracle.loaders import prague_cam_struct
    from oracle.photogrammetry import resect, forwintersect
    from oracle.bundle import bundle as obundle
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'prague2016sxb')
    s = prague_cam_struct(root, stub, 'ctrlpts-%s.txt' % cps)
    s.EO.val[:] = np.nan
    s.OP.val[:, ~s.prior.OP.isCtrl] = np.nan
    cpId = np.asarray(s.OP.id)[s.prior.OP.isCtrl]
    s1, _, fail = resect(s, 'all', cpId, 1, 0, cpId)
    assert not fail
    s2, _, _ = forwintersect(s1, 'all', True)
    s3, ok, it, s0, E = obundle(copy.deepcopy(s2), 'gna')
    assert ok and it == 4 and E.numParams == nparams
    tol = lambda v: 0.6 * 10 ** (np.floor(np.log10(v)) - 5)
    assert abs(E.res[0] - first) < tol(first) and abs(E.res[-1] - last) < tol(last) and abs(s0 - sigma0) < tol(sigma0)


@pytest.mark.parametrize('use_prior,iters,first,last,sigma0,nobs', [
    (True, 3, 86.8008, 38.2458, 1.06942, 2452), (False, 4, 60.1122, 38.2456, 1.07447, 2440)])
def test_sxb_prior_eo_demo_matches_golden_reports(use_prior, iters, first, last, sigma0, nobs):
    """sxb_prior_eo.m (`code/demo/sxb_prior_eo.m:30-100`): the StereoBox project with smart points,
    unshifted weighted control points and - with use_prior - prior observations of four camera positions
    (`ref/fake-camera-positions.txt`, sigma 5 cm; `setprioreo.m`).  Reports `sxb-prior-eo-dbatreport.txt`
    / `sxb-no-prior-eo-dbatreport.txt`: iterations, last error, sigma0 and the observation count (12 EO
    observations) to the printed digits; the first error to 2e-5 (the resection start values of this
    project sit on an ill-conditioned quartic, cf. the device test of the camcal demo)."""
    import copy, os
    import numpy as np
    from oracle.loaders import prague_cam_struct, load_pm_export, set_prior_eo_positions
    from oracle.photogrammetry import resect, forwintersect
    from oracle.bundle import bundle as obundle
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'prague2016sxb')
    s = prague_cam_struct(root, 'wsmart', 'ctrlpts-weighted.txt', shift_cp=False, orient='with')
    if use_prior:
        prob = load_pm_export(os.path.join(root, 'pmexports', 'wsmart-with-orient-pmexport.txt'))
        set_prior_eo_positions(s, prob, os.path.join(root, 'ref', 'fake-camera-positions.txt'))
    s.EO.val[:] = np.nan
    s.OP.val[:, ~s.prior.OP.isCtrl] = np.nan
    cpId = np.asarray(s.OP.id)[s.prior.OP.isCtrl]
    s1, _, fail = resect(s, 'all', cpId, 1, 0, cpId)
    assert not fail
    s2, _, _ = forwintersect(s1, 'all', True)
    s3, ok, it, s0, E = obundle(copy.deepcopy(s2), 'gna')
    assert ok and it == iters and E.numParams == 1173 and len(E.final.weighted.r) == nobs
    assert abs(E.res[0] - first) < 2e-5 * first
    assert abs(E.res[-1] - last) < 6e-4 and abs(s0 - sigma0) < 6e-6


@pytest.mark.parametrize('pm,rank,suspects', [
    ('camcal-pmexport-1ray.txt', 422, ['OZ-87/88']),
    ('camcal-pmexport-missing-obs.txt', 417, ['OX-12/13', 'OY-12/13', 'OZ-12/13', 'OX-59/60', 'OY-59/60', 'OZ-59/60'])])
def test_structural_weakness_matches_golden_reports(pm, rank, suspects):
    """bundle.m:431-446 on the two structurally deficient camcal projects: the reports list
    "Structural rank: 422 (deficiency: 1) ... OZ-87/88" and "Structural rank: 417 (deficiency: 6) ...
    OX/OY/OZ-12/13, OX/OY/OZ-59/60" (parameter types from buildparamtypes.m, unmatched columns from
    dmperm).  Checked for the restatement and for the product's host-side helpers on the same Jacobian."""
    import copy, os
    import numpy as np
    from oracle.loaders import camcal_pm_struct
    from oracle.photogrammetry import resect, forwintersect
    from oracle.bundle import bundle as obundle
    from dbat_b200.bundle import structural_weakness
    from dbat_b200.dbatstruct import paramtypes as product_paramtypes
    G = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'camcalpm')
    s = camcal_pm_struct(os.path.join(G, pm), os.path.join(G, 'camcal-fixed.txt'))
    cpId = np.asarray(s.OP.id)[s.prior.OP.isCtrl]
    s1, _, _ = resect(s, 'all', cpId, 1, 0, cpId)
    s2, _, _ = forwintersect(s1, 'all', True)
    s3, ok, it, s0, E = obundle(copy.deepcopy(s2), 'gna')
    w = E.weakness.structural
    assert E.code == -4 and w.rank == rank and w.deficiency == 423 - rank and w.suspectedParams == suspects
    assert list(E.paramTypes[:9]) == ['cc', 'px', 'py', 'as', 'K1', 'K2', 'K3', 'P1', 'P2'] and E.paramTypes[9] == 'EX-1'
    pt = product_paramtypes(s3)
    assert list(pt) == list(E.paramTypes)
    wp = structural_weakness(E.final.weighted.J, pt)
    assert wp.structural.rank == rank and wp.structural.suspectedParams == suspects
    assert np.array_equal(wp.structural.dmperm == 0, w.dmperm == 0)


def test_pointfirst_cpu_solver_equals_the_sparse_solve():
    """bench.py's CPU baseline solves (J'J + lambda I) p = -J'r by eliminating the 3 x 3 point blocks first and
    factoring the camera front densely (what CHOLMOD's supernodal factorisation does with an AMD ordering); it
    must give the step of the oracle's general sparse solve (`_solve_spd`, MATLAB's backslash)."""
    import scipy.sparse as sp
    from dbat_b200.synth import make_scene
    from oracle import lsa
    from oracle.cameramodel import brown_euler_cam4
    from oracle.dbatstruct import buildweightmatrix, serialize
    s, _ = make_scene(30, 1500, rays=6, seed=9)
    x = serialize(s)
    R = np.sqrt(buildweightmatrix(s))
    f, J = brown_euler_cam4(x, s, True)
    Jw = (sp.diags(R) @ J).tocsc()
    N = (Jw.T @ Jw).tocsc()
    g = Jw.T @ (R * f)
    n = len(x)
    nC = n - len(s.bundle.serial.OP.dest)
    A = (N + 1e-6 * N.diagonal().mean() * sp.identity(n, format='csc')).tocsc()
    p1, sing = lsa._solve_spd(A, -g)
    assert not sing
    p2 = lsa.solve_spd_pointfirst(A, -g, nC)
    np.testing.assert_allclose(p2, p1, rtol=1e-9, atol=1e-12 * np.abs(p1).max())
