"""Start-value step that precedes `bundle` in every demo: forward intersection (test infrastructure).

Restates, literally, `code/photogrammetry/forwintersect.m:19-46`, `pm_multiforwintersect.m:15-51`,
`pm_forwintersect3.m:11-73` and the no-derivative branch of `bundle/cammodel/pm_multilenscorr1.m:36-69`
with `pm_lens1.m:38-72`.  IO columns are `[f; pp(2); b(2); K(nK); P(nP)]`
(`cammodel/private/unpackio.m`), i.e. the DBAT struct's own storage.
"""
import numpy as np

from .cammodel_legacy import pm_eulerrotmat


def pm_lens1(p, p0, K, P):
    """pm_lens1.m:38-72 (values only): radial + tangential Brown distortion at p (2xn, mm)."""
    xBar, yBar = p[0] - p0[0], p[1] - p0[1]
    r2 = xBar ** 2 + yBar ** 2
    nK = len(K)
    if nK == 0 and len(P) == 0:
        return np.zeros_like(p)
    Kr = np.zeros_like(r2)
    for k in range(nK):
        Kr = Kr + K[k] * r2 ** (k + 1)
    dx, dy = xBar * Kr, yBar * Kr
    if len(P) > 0:
        P3 = P[2] if len(P) > 2 else 0.0
        dx = dx + (P[0] * (r2 + 2 * xBar ** 2) + 2 * P[1] * xBar * yBar) * (1 + P3)
        dy = dy + (P[1] * (r2 + 2 * yBar ** 2) + 2 * P[0] * xBar * yBar) * (1 + P3)
    return np.vstack([dx, dy])


def pm_multilenscorr1(p, IO, nK, nP, pxSize, cams):
    """pm_multilenscorr1.m:36-69: q = U p (mm), xy = q - lens(q) per camera; returns 2xn."""
    xy = np.zeros_like(p, dtype=float)
    for i in range(IO.shape[1]):
        ix = cams == i
        if not ix.any():
            continue
        pp = IO[1:3, i]
        K = IO[5:5 + nK, i]
        P = IO[5 + nK:5 + nK + nP, i]
        q = pxSize[:, i][:, None] * p[:, ix]
        xy[:, ix] = q - pm_lens1(q, pp, K, P)
    return xy


def pm_forwintersect3(Pm, xy):
    """pm_forwintersect3.m:11-73.  Pm: n x 3 x 4 camera matrices, xy: 2 x n x k.  Returns OP (3 x k), r (k)."""
    n, k = Pm.shape[0], xy.shape[2]
    C = np.zeros((3, n))
    for i in range(n):
        _, _, Vt = np.linalg.svd(Pm[i])
        h = Vt[-1]                                        # null(P)
        C[:, i] = h[:3] / h[3]
    t = np.zeros((3, k, n))
    for i in range(n):
        xy1 = np.vstack([xy[:, i, :], np.ones((1, k))])
        Ppx = np.linalg.pinv(Pm[i]) @ xy1                 # second point on every ray
        far = np.abs(Ppx[3]) < 1e-8
        Ppx[:, far] += np.append(C[:, i], 1.0)[:, None]
        t[:, :, i] = Ppx[:3] / Ppx[3] - C[:, i][:, None]
    t = t / np.sqrt((t ** 2).sum(axis=0, keepdims=True))
    OP = np.full((3, k), np.nan)
    r = np.full(k, np.nan)
    b = C.T.reshape(-1)                                   # [C1; C2; ...]
    for j in range(k):
        A = np.zeros((3 * n, 3 + n))
        for i in range(n):
            A[3 * i:3 * i + 3, :3] = np.eye(3)
            A[3 * i:3 * i + 3, 3 + i] = t[:, j, i]
        x = np.linalg.lstsq(A, b, rcond=None)[0]
        OP[:, j] = x[:3]
        r[j] = np.linalg.norm(b - A @ x) / n
    return OP, r


def pm_multiforwintersect(IO, EO, colPos, pts, idx):
    """pm_multiforwintersect.m:15-51.  colPos: nOP x nImg array of 1-based column numbers into pts (0 = not
    seen); idx: 0-based point indices.  One pm_forwintersect3 call per distinct camera combination."""
    nImg = EO.shape[1]
    Pm = np.full((nImg, 3, 4), np.nan)
    for j in np.flatnonzero((colPos[idx] != 0).any(axis=0)):
        RR = pm_eulerrotmat(EO[3:6, j])[0]
        CC = EO[0:3, j]
        Kc = np.array([[-IO[0, j], 0, IO[1, j]], [0, -IO[0, j], IO[2, j]], [0, 0, 1.0]])
        Pm[j] = Kc @ RR @ np.hstack([np.eye(3), -CC[:, None]])
    OP = np.full((3, len(idx)), np.nan)
    R = np.full(len(idx), np.nan)
    vis = colPos[idx] != 0
    combs, ui = np.unique(vis, axis=0, return_inverse=True)
    ui = np.asarray(ui).reshape(-1)
    for ii in range(combs.shape[0]):
        camIx = np.flatnonzero(combs[ii])
        if len(camIx) > 1:
            sel = np.flatnonzero(ui == ii)
            cols = colPos[np.asarray(idx)[sel]][:, camIx] - 1          # (points, cams)
            xy = pts[:, cols.T]                                        # 2 x cams x points
            OP[:, sel], R[sel] = pm_forwintersect3(Pm[camIx], xy)
    return OP, R


def forwintersect(s0, ids='all', skipPrior=False):
    """forwintersect.m:19-46: OP coordinates of the listed points by forward intersection; with skipPrior,
    points with fixed coordinates or prior observations are left alone.  Returns (s, id, res)."""
    import copy
    assert np.isfinite(s0.EO.val).all() and np.isfinite(s0.IO.val).all()
    nOP, nImg = s0.OP.val.shape[1], s0.EO.val.shape[1]
    allids = np.asarray(s0.OP.id) if getattr(s0.OP, 'id', None) is not None else np.arange(nOP)
    if isinstance(ids, str) and ids == 'all':
        ids = allids
    p = np.diag([1.0, -1.0]) @ s0.IP.val
    xy = pm_multilenscorr1(p, s0.IO.val, s0.IO.model.nK, s0.IO.model.nP, s0.IO.sensor.pxSize, np.asarray(s0.IP.cam))
    doEst = np.ones(nOP, bool)
    if skipPrior:
        doEst = s0.bundle.est.OP.all(axis=0) & ~s0.prior.OP.use.any(axis=0)
    idx = np.flatnonzero(np.isin(allids, ids) & doEst)
    colPos = np.zeros((nOP, nImg), dtype=np.int64)
    colPos[np.asarray(s0.IP.op), np.asarray(s0.IP.img)] = np.arange(1, s0.IP.val.shape[1] + 1)
    OP, res = pm_multiforwintersect(s0.IO.val, s0.EO.val, colPos, xy, idx)
    s = copy.deepcopy(s0)
    s.OP.val[:, idx] = OP
    return s, allids[idx], res


# ------------------------------------------------------------------------------------------------
# spatial resection (code/photogrammetry/resect.m, pm_resect_3pt.m, misc/largesttriangle.m,
# photogrammetry/derotmat3d.m)
# ------------------------------------------------------------------------------------------------
def largesttriangle(pts):
    """largesttriangle.m:21-41 (cHull=true): all triangles among the convex-hull points, largest area
    first.  pts 2xN.  Returns (T, A): index triples (0-based, rows in nchoosek order before the stable
    sort) and areas."""
    from itertools import combinations
    from scipy.spatial import ConvexHull
    hull = np.unique(ConvexHull(pts.T).simplices.ravel())
    T = np.array(list(combinations(hull.tolist(), 3)), dtype=int)
    x, y = pts[0][T], pts[1][T]
    A = 0.5 * np.abs(x[:, 0] * (y[:, 1] - y[:, 2]) + x[:, 1] * (y[:, 2] - y[:, 0]) + x[:, 2] * (y[:, 0] - y[:, 1]))
    i = np.argsort(-A, kind='stable')
    return T[i], A[i]


def _subspace(a, b):
    """MATLAB subspace() for two vectors: the acute angle between the lines they span."""
    a = a / np.linalg.norm(a)
    b = b / np.linalg.norm(b)
    if abs(a @ b) < np.sqrt(0.5):
        return np.arccos(min(1.0, abs(a @ b)))
    return np.arcsin(min(1.0, np.linalg.norm(b - a * (a @ b))))


def derotmat3d(M):
    """derotmat3d.m:19-21: (omega, phi, kappa) of a rotation matrix."""
    return np.array([np.arctan2(-M[2, 1], M[2, 2]), np.arcsin(M[2, 0]), np.arctan2(-M[1, 0], M[0, 0])])


def pm_resect_3pt(X, x, use, behind=False):
    """pm_resect_3pt.m:38-147 (relax=false): Grunert's 3-point resection (quartic in v), absolute
    orientation for every admissible root, mean reprojection residual over all test points.
    X 3xN object points, x 2xN normalised image points, use: boolean mask of exactly 3 points.
    Returns (P of the best solution or None, list of all P, residuals)."""
    assert np.count_nonzero(use) == 3
    XT, xT = X, x
    X3 = X[:, use]
    x3 = np.vstack([x[:, use], np.ones(3)])
    x3 = x3 / np.sqrt((x3 ** 2).sum(axis=0))
    a = np.linalg.norm(X3[:, 2] - X3[:, 1])
    b = np.linalg.norm(X3[:, 2] - X3[:, 0])
    c = np.linalg.norm(X3[:, 1] - X3[:, 0])
    ca, cb, cg = np.cos(_subspace(x3[:, 1], x3[:, 2])), np.cos(_subspace(x3[:, 0], x3[:, 2])), np.cos(_subspace(x3[:, 0], x3[:, 1]))
    q1, q2 = (a * a - c * c) / (b * b), (a * a + c * c) / (b * b)
    q3, q4 = (b * b - c * c) / (b * b), (b * b - a * a) / (b * b)
    A4 = (q1 - 1) ** 2 - 4 * c * c / (b * b) * ca ** 2
    A3 = 4 * (q1 * (1 - q1) * cb + 2 * c * c / (b * b) * ca ** 2 * cb - (1 - q2) * ca * cg)
    A2 = 2 * (q1 ** 2 + 2 * q1 ** 2 * cb ** 2 + 2 * q3 * ca ** 2 + 2 * q4 * cg ** 2 - 4 * q2 * ca * cb * cg - 1)
    A1 = 4 * (-q1 * (1 + q1) * cb + 2 * a * a / (b * b) * cg ** 2 * cb - (1 - q2) * ca * cg)
    A0 = (1 + q1) ** 2 - 4 * a * a / (b * b) * cg ** 2
    v = np.roots([A4, A3, A2, A1, A0])
    v = np.real(v[np.abs(np.imag(v) / np.abs(v)) < 1e-3])
    u = ((-1 + q1) * v ** 2 - 2 * q1 * cb * v + 1 + q1) / (2 * (cg - v * ca))
    s1 = np.sqrt(b * b / (1 + v ** 2 - 2 * v * cb))
    s3, s2 = v * s1, u * s1
    ok = (s1 >= 0) & (s2 >= 0) & (s3 >= 0)
    s123 = np.unique(np.vstack([s1[ok], s2[ok], s3[ok]]).T, axis=0)
    PP, res = [], []
    for s in s123:
        cx = s[None, :] * x3
        if behind:
            cx = -cx
        def frame(p):
            ob, oc = p[:, 2] - p[:, 0], p[:, 1] - p[:, 0]
            r1 = ob / np.linalg.norm(ob)
            r2 = np.cross(ob, oc); r2 = r2 / np.linalg.norm(r2)
            r3 = np.cross(ob, np.cross(ob, oc)); r3 = r3 / np.linalg.norm(r3)
            return np.column_stack([r1, r2, r3])
        cRo = frame(cx) @ frame(X3).T
        oxO = X3[:, 0] - cRo.T @ cx[:, 0]
        P = cRo @ np.hstack([np.eye(3), -oxO[:, None]])
        pr = P @ np.vstack([XT, np.ones(XT.shape[1])])
        PP.append(P)
        res.append(np.sqrt(np.mean(((pr[:2] / pr[2] - xT) ** 2).sum(axis=0))))
    res = np.array(res)
    return (PP[int(np.argmin(res))] if len(res) else None), PP, res


def resect(s0, cams='all', cpId=None, n=1, v=0.0, chkId=None):
    """resect.m:52-131: 3-point spatial resection of the listed cameras from control points cpId (the
    n triangles of largest measured area above v times the largest), best solution by the reprojection
    residual of the visible points in cpId U chkId.  Returns (s, rms, fail)."""
    import copy
    nImg, nOP = s0.EO.val.shape[1], s0.OP.val.shape[1]
    ids = np.asarray(s0.OP.id)
    if isinstance(cams, str) and cams == 'all':
        cams = np.arange(nImg)
    if chkId is None:
        chkId = ids
    xy = pm_multilenscorr1(np.diag([1.0, -1.0]) @ s0.IP.val, s0.IO.val, s0.IO.model.nK, s0.IO.model.nP,
                           s0.IO.sensor.pxSize, np.asarray(s0.IP.cam))
    colPos = -np.ones((nOP, nImg), dtype=np.int64)
    colPos[np.asarray(s0.IP.op), np.asarray(s0.IP.img)] = np.arange(s0.IP.val.shape[1])
    s = copy.deepcopy(s0)
    rms = np.full(len(cams), np.nan)
    fail = False
    for i, cam in enumerate(cams):
        IO = s0.IO.val[:, cam]
        K = np.array([[-IO[0], 0, IO[1]], [0, -IO[0], IO[2]], [0, 0, 1.0]])
        visPts = np.flatnonzero(colPos[:, cam] >= 0)                  # object points seen, ascending index
        isCp = np.isin(ids[visPts], cpId)
        meaIx = visPts[isCp]
        if len(meaIx) > 3:
            T, A = largesttriangle(xy[:, colPos[meaIx, cam]])
            sel = (np.arange(1, len(A) + 1) <= n) & (A >= v * A[0])
            tryId = ids[meaIx[T[sel]]]
        elif len(meaIx) == 3:
            tryId = ids[meaIx][None, :]
        else:
            tryId = np.zeros((0, 3), dtype=ids.dtype)
        bestRes, bestP = np.inf, None
        for useId in tryId:
            pt2 = xy[:, colPos[visPts, cam]]
            pt2N = np.linalg.solve(K, np.vstack([pt2, np.ones(pt2.shape[1])]))
            pt3 = s0.OP.val[:, visPts]
            visId = ids[visPts]
            keep = np.isin(visId, np.union1d(cpId, chkId))
            P, _, res = pm_resect_3pt(pt3[:, keep], pt2N[:2, keep] / pt2N[2, keep], np.isin(visId[keep], useId), True)
            if len(res) and res.min() < bestRes:
                bestRes, bestP = res.min(), P
        rms[i] = bestRes
        if bestP is not None:
            _, _, Vt = np.linalg.svd(bestP)
            h = Vt[-1]
            s.EO.val[0:3, cam] = h[:3] / h[3]                         # euclidean(null(P))
            s.EO.val[3:6, cam] = derotmat3d(bestP[:, :3])
        else:
            fail = True
            s.EO.val[:, cam] = np.nan
    return s, rms, fail
