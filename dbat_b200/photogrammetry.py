"""Host mirrors of the reference's start-value steps `resect` (code/photogrammetry/resect.m) and
`forwintersect` (code/photogrammetry/forwintersect.m); the numerical work runs on the device through
`dbat_resect3` / `dbat_forwintersect` (csrc/startval.cu).  No CPU fallback."""
import copy
import ctypes as C

import numpy as np

from . import _lib


def forwintersect(s0, ids='all', skipPrior=False, return_ms=False):
    """[s,id,res]=forwintersect(s0,ids,skipPrior) (forwintersect.m:1-46): OP coordinates of the listed
    points by forward intersection from the current IO/EO; with skipPrior, points with fixed coordinates
    or prior observations are left alone; points seen in fewer than two images get NaN."""
    if not np.isfinite(s0.EO.val).all():
        raise ValueError('Bad or uninitialized EO data')            # forwintersect.m:19
    if not np.isfinite(s0.IO.val).all():
        raise ValueError('Bad or uninitialized IO data')            # forwintersect.m:20
    nOP, nImg = s0.OP.val.shape[1], s0.EO.val.shape[1]
    allids = np.asarray(s0.OP.id) if getattr(s0.OP, 'id', None) is not None else np.arange(nOP)
    if isinstance(ids, str) and ids == 'all':
        ids = allids
    doEst = np.ones(nOP, bool)
    if skipPrior:
        doEst = s0.bundle.est.OP.all(axis=0) & ~s0.prior.OP.use.any(axis=0)
    idx = np.flatnonzero(np.isin(allids, ids) & doEst)
    IO = np.asfortranarray(s0.IO.val, dtype=np.float64)
    EO = np.asfortranarray(s0.EO.val, dtype=np.float64)
    px = np.asfortranarray(np.broadcast_to(s0.IO.sensor.pxSize, (2, nImg)), dtype=np.float64)
    IP = np.asfortranarray(s0.IP.val, dtype=np.float64)
    img = _lib.i64(np.asarray(s0.IP.img) + 1)
    op = _lib.i64(np.asarray(s0.IP.op) + 1)
    pts = _lib.i64(idx + 1)
    d = _lib.FwiDesc(nImg, nOP, IP.shape[1], IO.shape[0], int(s0.IO.model.nK), int(s0.IO.model.nP),
                     IO.ctypes.data_as(_lib.c_dp), EO.ctypes.data_as(_lib.c_dp), px.ctypes.data_as(_lib.c_dp),
                     IP.ctypes.data_as(_lib.c_dp), _lib.iptr(img), _lib.iptr(op), _lib.iptr(pts), len(idx))
    OP = np.empty((3, len(idx)), order='F')
    res = np.empty(len(idx))
    ms = C.c_double()
    L = _lib.lib()
    rc = L.dbat_forwintersect(C.byref(d), OP.ctypes.data_as(_lib.c_dp), _lib.dptr(res), C.byref(ms))
    if rc != 0:
        raise _lib.DbatError(rc, L.dbat_forwintersect_error().decode())
    s = copy.deepcopy(s0)
    s.OP.val[:, idx] = OP
    out = (s, allids[idx], res)
    return out + (ms.value,) if return_ms else out


def _lenscorr(s0, cols):
    """Lens-corrected measurements (mm) of the image points `cols` (pm_multilenscorr1.m:36-69,
    pm_lens1.m:38-72; only the few control/check point measurements a resection needs)."""
    cam = np.asarray(s0.IP.cam)[cols]
    IO = s0.IO.val[:, cam]
    nK, nP = int(s0.IO.model.nK), int(s0.IO.model.nP)
    px = np.broadcast_to(s0.IO.sensor.pxSize, (2, s0.IO.val.shape[1]))[:, cam]
    q = px * (np.array([[1.0], [-1.0]]) * s0.IP.val[:, cols])
    xb, yb = q[0] - IO[1], q[1] - IO[2]
    r2 = xb ** 2 + yb ** 2
    Kr = np.zeros_like(r2)
    pw = np.ones_like(r2)
    for k in range(nK):
        pw = pw * r2
        Kr = Kr + IO[5 + k] * pw
    dx, dy = xb * Kr, yb * Kr
    if nP > 0:
        P1 = IO[5 + nK]
        P2 = IO[6 + nK] if nP > 1 else 0.0
        P3 = IO[7 + nK] if nP > 2 else 0.0
        dx = dx + (P1 * (r2 + 2 * xb ** 2) + 2 * P2 * xb * yb) * (1 + P3)
        dy = dy + (P2 * (r2 + 2 * yb ** 2) + 2 * P1 * xb * yb) * (1 + P3)
    return np.vstack([q[0] - dx, q[1] - dy])


def _largest_triangles(pts):
    """largesttriangle.m:21-41: triangles among the convex-hull points, largest area first."""
    from itertools import combinations
    from scipy.spatial import ConvexHull
    hull = np.unique(ConvexHull(pts.T).simplices.ravel())
    T = np.array(list(combinations(hull.tolist(), 3)), dtype=int)
    x, y = pts[0][T], pts[1][T]
    A = 0.5 * np.abs(x[:, 0] * (y[:, 1] - y[:, 2]) + x[:, 1] * (y[:, 2] - y[:, 0]) + x[:, 2] * (y[:, 0] - y[:, 1]))
    i = np.argsort(-A, kind='stable')
    return T[i], A[i]


def resect(s0, cams='all', cpId=None, n=1, v=0.0, chkId=None):
    """[s,rms,fail]=resect(s0,cams,cpId,n,v,chkId) (resect.m:1-131): EO of the listed cameras by 3-point
    resection from the control points cpId: the n triangles of largest measured area (at least v times
    the largest) are tried, the best solution is the one with the smallest reprojection residual over
    the visible points in cpId U chkId.  Cameras that cannot be resected get NaN and fail=True."""
    nImg, nOP = s0.EO.val.shape[1], s0.OP.val.shape[1]
    ids = np.asarray(s0.OP.id) if getattr(s0.OP, 'id', None) is not None else np.arange(nOP)
    if isinstance(cams, str) and cams == 'all':
        cams = np.arange(nImg)
    cams = np.asarray(cams)
    if chkId is None:
        chkId = ids
    isCp = np.isin(ids, cpId)
    isTest = np.isin(ids, np.union1d(cpId, chkId))
    ip_img, ip_op = np.asarray(s0.IP.img), np.asarray(s0.IP.op)
    cols = np.flatnonzero(isTest[ip_op] & np.isin(ip_img, cams))
    cols = cols[np.lexsort((ip_op[cols], ip_img[cols]))]             # by image, then object point index
    xy = _lenscorr(s0, cols)
    IO = s0.IO.val[:, ip_img[cols]]
    xN = np.vstack([(xy[0] - IO[1]) / -IO[0], (xy[1] - IO[2]) / -IO[0]])   # K \ [x;y;1]
    X3, x3, tstart, XT, xT, owner = [], [], [0], [], [], []
    for ci, cam in enumerate(cams):
        sel = np.flatnonzero(ip_img[cols] == cam)
        opi = ip_op[cols[sel]]
        cp = np.flatnonzero(isCp[opi])
        if len(cp) > 3:
            T, A = _largest_triangles(xy[:, sel[cp]])
            tri = T[(np.arange(1, len(A) + 1) <= n) & (A >= v * A[0])]
        elif len(cp) == 3:
            tri = np.arange(3)[None, :]
        else:
            tri = np.zeros((0, 3), dtype=int)
        for t3 in tri:
            use = np.sort(cp[t3])                                    # ismember() keeps the visible order
            X3.append(s0.OP.val[:, opi[use]].T.reshape(-1))
            x3.append(xN[:, sel[use]].T.reshape(-1))
            XT.append(s0.OP.val[:, opi].T.reshape(-1))
            xT.append(xN[:, sel].T.reshape(-1))
            tstart.append(tstart[-1] + len(sel))
            owner.append(ci)
    s = copy.deepcopy(s0)
    rms = np.full(len(cams), np.nan)
    best = np.full(len(cams), np.inf)
    s.EO.val[:, cams] = np.nan
    if owner:
        X3a, x3a = _lib.f64(np.concatenate(X3)), _lib.f64(np.concatenate(x3))
        XTa, xTa = _lib.f64(np.concatenate(XT)), _lib.f64(np.concatenate(xT))
        ts = _lib.i64(tstart)
        d = _lib.ResectDesc(len(owner), _lib.dptr(X3a), _lib.dptr(x3a), _lib.iptr(ts), _lib.dptr(XTa), _lib.dptr(xTa), 1)
        EO = np.empty((len(owner), 6))
        res = np.empty(len(owner))
        L = _lib.lib()
        rc = L.dbat_resect3(C.byref(d), _lib.dptr(EO), _lib.dptr(res))
        if rc != 0:
            raise _lib.DbatError(rc, L.dbat_forwintersect_error().decode())
        for k, ci in enumerate(owner):
            if np.isfinite(res[k]) and res[k] < best[ci]:
                best[ci] = res[k]
                s.EO.val[:, cams[ci]] = EO[k]
    rms[:] = best
    fail = bool(np.isinf(best).any())
    return s, rms, fail
