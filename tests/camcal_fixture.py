"""Shared helpers for the camcal golden fixture (tests only)."""
import os

import numpy as np

from oracle import loaders
from oracle.cameramodel import eulerrotmat123

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'camcaldemo')


def golden_camera_xml():
    return open(os.path.join(GOLD, 'result', 'c4040z.xml')).read()


def golden_eo():
    return loaders.load_camera_stations(os.path.join(GOLD, 'result', 'camera_stations.txt'))


def triangulate(s, IOcol, EO):
    """Linear forward intersection of every non-fixed OP from (ideal pinhole) rays.

    Test-side start-value helper, NOT a restatement of forwintersect.m: it only has to
    land inside the convergence basin.
    """
    nOP = s.OP.val.shape[1]
    A = np.zeros((nOP, 3, 3))
    bb = np.zeros((nOP, 3))
    f, px, py = IOcol[0], IOcol[1], IOcol[2]
    sz = s.IO.sensor.pxSize[0, 0]
    for k in range(len(s.IP.img)):
        i, j = s.IP.img[k], s.IP.op[k]
        M, _ = eulerrotmat123(EO[3:6, i])
        u = s.IP.val[:, k]
        xm = np.array([sz * u[0] - px, -sz * u[1] - py, -f])   # lhs = -f*h  ⇒ ray ∝ (x, y, -f)
        d = M @ xm
        d /= np.linalg.norm(d)
        P = np.eye(3) - np.outer(d, d)
        A[j] += P
        bb[j] += P @ EO[0:3, i]
    OP = s.OP.val.copy()
    for j in range(nOP):
        if s.bundle.est.OP[:, j].all():
            OP[:, j] = np.linalg.solve(A[j], bb[j])
    return OP


def camcal_struct(start='golden', seed=0, noise=1.0):
    """camcal project with start values near the golden solution.

    start='golden': calibrated IO + golden EO (then perturbed by `noise`), OP triangulated.
    start='default': default IO (focal 7.5, pp at centre, no distortion) + perturbed golden EO.
    """
    s = loaders.load_camcal_script(GOLD, golden_camera_xml() if start == 'golden' else None)
    ids, EO, _ = golden_eo()
    assert np.array_equal(ids, s.EO.id)
    rng = np.random.default_rng(seed)
    EO = EO.copy()
    EO[0:3] += noise * 0.01 * rng.standard_normal(EO[0:3].shape)
    EO[3:6] += noise * 0.002 * rng.standard_normal(EO[3:6].shape)
    s.EO.val = EO
    s.OP.val = triangulate(s, s.IO.val[:, 0], EO)
    return s
