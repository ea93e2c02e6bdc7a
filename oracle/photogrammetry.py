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
