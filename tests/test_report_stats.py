"""Unit tests of the statistics behind the result file (dbat_b200/report.py) on hand-made inputs."""
from types import SimpleNamespace as NS

import numpy as np
import pytest
import scipy.sparse as sp
from scipy.integrate import quad
from scipy.special import gamma

from dbat_b200 import report


def test_corrmat_contract():
    """corrmat.m:21-47: unit diagonal (or zero with nodiag), symmetric, clipped to [-1, 1], NaN kept."""
    rng = np.random.default_rng(3)
    A = rng.standard_normal((6, 6))
    C = A @ A.T
    R, sd = report.corrmat(C)
    np.testing.assert_allclose(sd, np.sqrt(np.diag(C)))
    np.testing.assert_allclose(R, C / np.outer(sd, sd), atol=1e-15)
    assert np.all(np.diag(R) == 1) and np.all(np.diag(report.corrmat(C, True)[0]) == 0)
    C2 = C.copy()
    C2[0, 1] = C2[1, 0] = 2 * sd[0] * sd[1]                       # inconsistent entry: clipped
    assert report.corrmat(C2)[0][0, 1] == 1.0
    C3 = C.copy()
    C3[2, :] = C3[:, 2] = 0                                       # fixed parameter: 0/0 stays NaN (:36,47), diagonal too
    R3 = report.corrmat(C3, True)[0]
    assert np.isnan(R3[2, 0]) and np.isnan(R3[2, 2]) and not np.isnan(R3[0, 1])
    Rs, _ = report.corrmat(sp.csc_matrix(C))
    np.testing.assert_array_equal(Rs, R)
    Rb, _ = report.corrmat(np.stack([C, 4 * C]))                  # a stack of blocks
    np.testing.assert_allclose(Rb[1], R, atol=1e-15)


@pytest.mark.parametrize('n', [1, 2, 3])
def test_cumchi2_equals_the_reference_quadrature(n):
    """cumchi2.m integrates the chi-square density with quad; the closed form used here is that integral."""
    dens = lambda x: 2 ** (-n / 2) / gamma(n / 2) * x ** (n / 2 - 1) * np.exp(-x / 2)
    for x in (0.003, 0.5, 1.0, 3.84, 9.0, 40.0, 99.0):
        assert abs(report.cumchi2(x, n) - quad(dens, 0, x)[0]) < 1e-7
    assert report.cumchi2(101.0, n) == 1.0 and report.cumchi2(0.0, n) == 0.0 and report.cumchi2(-1.0, n) == 0.0
    assert np.isnan(report.cumchi2(np.nan, n))
    assert abs(report.cumchi2(3.841458820694124, 1) - 0.95) < 1e-12


def _blocks_to_cov(blocks):
    return sp.block_diag([sp.csc_matrix(b) for b in blocks]).tocsc()


def test_high_correlation_scans_report_matlab_find_order():
    """Pairs come in the column-major order of find(abs(tril(R))>thres) on the big matrix: by block, then by
    column, then by row - the order the result file lists them in."""
    def block(pairs, k):
        C = np.eye(k)
        for (i, j), v in pairs.items():
            C[i, j] = C[j, i] = v
        return C
    eo = [np.eye(6), block({(5, 3): 0.99, (4, 0): -0.97, (2, 0): 0.96}, 6)]
    op = [block({(2, 0): 0.999}, 3), np.eye(3), block({(1, 0): -0.96, (2, 1): 0.98}, 3)]
    s = NS(EO=NS(val=np.zeros((6, 2)), struct=NS(block=np.tile([1, 2], (6, 1)))), OP=NS(val=np.zeros((3, 3))))
    cov = lambda s_, e_, w: _blocks_to_cov(eo if w == 'CEO' else op)
    i, j, k, v, _ = report.high_eo_correlations(s, None, 0.95, cov)
    assert list(zip(i, j, k)) == [(2, 0, 1), (4, 0, 1), (5, 3, 1)]
    np.testing.assert_allclose(v, [0.96, -0.97, 0.99])
    i, j, k, v = report.high_op_correlations(s, None, 0.95, cov=cov)
    assert list(zip(i, j, k)) == [(2, 0, 0), (1, 0, 2), (2, 1, 2)]
    i, j, k, v = report.high_op_correlations(s, None, 0.985, cov=cov)
    assert list(k) == [0]


def test_distortion_significance_including_the_reference_quirks():
    """test_distortion_params.m: K individually and cumulatively, P jointly (stored on the P1 row only,
    `P(j,:)=...` at :59), aspect / skew individually; NaN where a coefficient is not estimated."""
    NC = 10
    x = np.zeros((NC, 2))
    x[:, 0] = [7.0, 3.0, -2.0, 1e-3, 0.0, 4e-3, -2e-5, 3e-6, 1e-4, -3e-4]
    x[:, 1] = x[:, 0]
    sd = np.array([1e-2, 1e-2, 1e-2, 5e-4, 1.0, 1e-3, 2e-5, 1e-5, 1e-4, 1e-4])
    C = np.diag(sd ** 2)
    est = np.ones((NC, 2), bool)
    est[4] = False                                                # skew fixed
    est[7] = False                                                # K3 fixed
    s = NS(IO=NS(val=x, model=NS(nK=3, nP=2), struct=NS(block=np.ones((NC, 2), int))), bundle=NS(est=NS(IO=est)))
    e = NS(final=NS(factorized=None))
    K, P, B, KC = report.test_distortion_params(s, e, np.stack([C, C]))
    c = report.cumchi2
    np.testing.assert_allclose(K[:, 0], [c(16.0, 1), c(1.0, 1), np.nan])
    np.testing.assert_allclose(KC[:, 0], [c(16.0, 1), c(17.0, 2), np.nan])
    np.testing.assert_allclose(P[:, 0], [c(1.0 + 9.0, 2), np.nan])          # joint value on the P1 row
    np.testing.assert_allclose(B[:, 0], [c(4.0, 1), np.nan])
    assert np.isnan(K[:, 1]).all()                                 # second image shares the camera: not unique
    e.final.factorized = NS(fail=True)
    assert all(np.isnan(a).all() for a in report.test_distortion_params(s, e, np.stack([C, C])))


def test_angles_and_coverage_on_a_toy_block():
    """angles.m / coverage.m: a point seen from two stations at right angles; one ray -> 0; none -> NaN;
    coverage of four corner points in a 100 x 50 image."""
    s = NS()
    s.EO = NS(val=np.array([[0.0, 10.0, 5.0], [0, 0, 0], [0, 0, 0], [0, 0, 0], [0, 0, 0], [0, 0, 0]]), name=['a', 'b', 'c'])
    s.OP = NS(val=np.array([[0.0, 5.0, 0.0, 1.0], [10.0, 5.0, 1.0, 1.0], [0.0, 0.0, 0.0, 0.0]]))
    s.IP = NS(op=np.array([0, 0, 1, 1, 1, 2]), img=np.array([0, 1, 0, 1, 2, 0]),
              val=np.array([[10.0, 90.0, 10.0, 90.0, 50.0, 50.0], [10.0, 10.0, 40.0, 40.0, 25.0, 25.0]]))
    a = report.angles(s)
    np.testing.assert_allclose(a[0], np.pi / 4)                   # rays (0,10,0) and (-10,10,0)
    np.testing.assert_allclose(a[1], np.pi / 2)                   # (5,5,0) and (-5,5,0) are orthogonal
    assert a[2] == 0 and np.isnan(a[3])
    s.IO = NS(val=np.array([[50.0] * 3, [50.0] * 3, [-25.0] * 3]), sensor=NS(pxSize=np.ones((2, 3)), imSize=np.tile([[100.0], [50.0]], (1, 3))))
    s.IP.img = np.array([0, 0, 0, 0, 1, 1])
    c, cr, crr = report.coverage(s, np.arange(3))
    np.testing.assert_allclose([c[0], cr[0]], [80 * 30 / 5000, 80 * 30 / 5000])
    assert c[1] == 0 and cr[1] == 0 and np.isnan(c[2])            # two points: no area; no points: NaN
    np.testing.assert_allclose(crr[0], np.hypot(40, 15) / np.hypot(50.5, 25.5))
    uc, ucr, ucrr = report.coverage(s, np.arange(3), True)
    np.testing.assert_allclose([uc, ucr], [80 * 30 / 5000, 80 * 30 / 5000])


def test_result_file_branches_the_goldens_do_not_reach():
    """A small synthetic project with a control point nobody observes, a labelled check point and an object
    point seen in one image only: the report must come out (no exception) with the reference's wording for
    those cases (`bundle_result_file.m:580-586,735-738,768-790`)."""
    import copy
    from dbat_b200.synth import make_scene
    from oracle.bundle import bundle as obundle, bundle_cov as ocov
    s, truth = make_scene(6, 40, rays=4, seed=11, build_indices=False)
    nOP = s.OP.val.shape[1]
    s.IO.val[:] = truth['IO'][:, None]
    s.EO.val[:] = truth['EO']
    s.OP.val[:] = truth['OP']
    s.OP.id = np.arange(1, nOP + 1)
    s.OP.label = [''] * nOP
    s.EO.name = ['img%d.jpg' % i for i in range(6)]
    s.proj = NS(title='toy', UUID='', fileName='', cptFile='', EOfile='', objUnit='m', x0desc='')
    s.IO.model.camUnit = 'mm'
    s.IP.sigmas = np.array([1.0])
    s.bundle.est.IO[:] = False
    s.prior.OP.isCtrl = np.zeros(nOP, bool)
    s.prior.OP.isCheck = np.zeros(nOP, bool)
    for j in range(5):                                            # five control points, fixed at the truth
        s.prior.OP.isCtrl[j] = True
        s.prior.OP.val[:, j] = truth['OP'][:, j]
        s.prior.OP.std[:, j] = 0.0
        s.bundle.est.OP[:, j] = False
        s.OP.label[j] = 'CP%d' % j
    drop = np.asarray(s.IP.op) == 4                               # control point 5 is never observed
    lone = np.flatnonzero(np.asarray(s.IP.op) == 20)[1:]          # object point 21 keeps a single ray
    keep = ~drop
    keep[lone] = False
    for k in ('val', 'std'):
        setattr(s.IP, k, getattr(s.IP, k)[:, keep])
    for k in ('img', 'op', 'cam'):
        setattr(s.IP, k, np.asarray(getattr(s.IP, k))[keep])
    s.bundle.est.OP[:, 20] = False                                # a one-ray point cannot be estimated
    s.prior.OP.isCheck[7] = True                                  # a check point with a label
    s.prior.OP.val[:, 7] = truth['OP'][:, 7] + 0.01
    s.prior.OP.std[:, 7] = 0.02
    s.OP.label[7] = 'CHK'
    s, ok, it, s0, E = obundle(copy.deepcopy(s), 'gna')
    assert ok
    s, lines = report.bundle_result_file(s, E, None, cov=ocov)
    text = '\n'.join(lines)
    assert 'CP ray count: 1x0, ' in text and '1 points with 0 rays.' in text
    assert 'Ignoring 1 CP with 0 rays.' in text
    assert 'Number of check pts: 1' in text and 'Check point delta' in text and '(CHK, pt 8)' in text
    assert 'CCP ray count: ' in text and ', label CHK)' in text
    assert '1 points with 1 rays.' in text and lines[-1] == 'End of result file'
    s, stats = report.writestats(s, None, 'toy')
    assert any(l.startswith('CP with lowest ray count') for l in stats)
