"""Legacy distortion models 1 (backward, PhotoModeler) and -1 (forward, PhotoScan/CV) — oracle
restatement (test infrastructure).

Follows `code/bundle/cameramodel/brown_euler_cam4.m:36-121` (model 1) and `:184-288` (model -1),
`code/bundle/cammodel/multieulerpinhole.m:53-227`, `eulerpinhole.m:28-56`, `roteuler.m:22-51`,
`pm_eulerrotmat.m:23-36`, `proj.m:20-49`, `multilensdist.m:54-175`, `browndist.m:103-253`,
`multiscalepts.m:8-19`, `private/createiocolumnindices.m:10-16`.  Both models force a single shared
camera (`:39-41`, `:186-188`): only IO column 1 is used and estimated.
"""
import numpy as np
import scipy.sparse as sp

from .cameramodel import _unpackio, dest_cols, prior_obs
from .dbatstruct import vis_matrix


def pm_eulerrotmat(ang):
    """pm_eulerrotmat.m:23-36 ('xyz'): M = R3(-k) R2(p) R1(-w) and its three derivatives."""
    def r2d(axis, phi):
        c, s = np.cos(phi), np.sin(phi)
        M = np.eye(3)
        ix = {1: [1, 2], 2: [0, 2], 3: [0, 1]}[axis]
        M[np.ix_(ix, ix)] = [[c, -s], [s, c]]
        return M
    M1, M2, M3 = r2d(1, -ang[0]), r2d(2, ang[1]), r2d(3, -ang[2])
    M = M3 @ M2 @ M1
    Px = np.array([[0, 0, 0], [0, 0, 1], [0, -1, 0.]])
    Py = np.array([[0, 0, -1], [0, 0, 0], [1, 0, 0.]])
    Pz = np.array([[0, 1, 0], [-1, 0, 0], [0, 0, 0.]])
    return M, M @ Px, M3 @ M2 @ Py @ M1, Pz @ M


def eulerpinhole(pp, f, P, C, ang, want_jac):
    """eulerpinhole.m:28-56 = proj(roteuler(...)): p = pp - f*[U/W;V/W]."""
    M, da1, da2, da3 = pm_eulerrotmat(ang)
    dP_ = P - C[:, None]
    T = M @ dP_                                              # roteuler.m:33-37
    UVW = T[0:2] / T[2]
    p = pp[:, None] - f * UVW                                # proj.m:27-30
    if not want_jac:
        return p, None
    N = P.shape[1]
    dpdT = np.zeros((N, 2, 3))                               # proj.m:40-48
    dpdT[:, 0, 0] = -f / T[2]
    dpdT[:, 1, 1] = -f / T[2]
    dpdT[:, 0, 2] = f * UVW[0] / T[2]
    dpdT[:, 1, 2] = f * UVW[1] / T[2]
    d = {'df': -UVW.T.copy(),                                # proj.m:36-38
         'dO': dpdT @ M,                                     # roteuler.m:46-53
         'dC': dpdT @ (-M),                                  # roteuler.m:55-57
         'dAng': np.stack([np.einsum('nij,jn->ni', dpdT, da @ dP_) for da in (da1, da2, da3)], axis=2)}
    return p, d


def browndist(s, pp, K, P, want_jac):
    """browndist.m:103-253: d = w*(sum K_k r^2k) + tangential(P); dK, dP, and the 2x2 blocks dg/dw."""
    w = s - pp[:, None]
    x, y = w
    r2 = x * x + y * y
    N = s.shape[1]
    nK, nP = len(K), len(P)
    k = np.arange(1, nK + 1)
    r2k = np.power(r2[:, None], k[None, :]) if nK else np.zeros((N, 0))
    Kr = r2k @ K if nK else np.zeros(N)
    d = w * Kr
    Aw = np.zeros((N, 2, 2))
    Aw[:, 0, 0] = r2 + 2 * x * x; Aw[:, 1, 0] = 2 * x * y
    Aw[:, 0, 1] = 2 * x * y;      Aw[:, 1, 1] = r2 + 2 * y * y
    Sr = np.zeros(N)
    if nP:
        AwQ = Aw @ P[0:2]                                     # (N,2)
        dt = AwQ.T
        if nP > 2:
            ks = np.arange(1, nP - 1)
            r2s = np.power(r2[:, None], ks[None, :])
            Sr = r2s @ P[2:]
            dt = dt * (1 + Sr)
        d = d + dt
    if not want_jac:
        return d, None
    out = {}
    out['dK'] = w.T[:, :, None] * r2k[:, None, :]             # :134-138
    if nP == 0:
        out['dP'] = np.zeros((N, 2, 0))
    elif nP == 2:
        out['dP'] = Aw
    else:
        out['dP'] = np.concatenate([Aw * (1 + Sr)[:, None, None],
                                    AwQ[:, :, None] * r2s[:, None, :]], axis=2)    # :186-190
    G = np.zeros((N, 2, 2))                                   # dg/dw
    if nK:
        Kdr = (np.power(r2[:, None], k[None, :] - 1) * k[None, :]) @ K
        G[:, 0, 0] += Kr + 2 * Kdr * x * x; G[:, 0, 1] += 2 * Kdr * x * y       # :148-157
        G[:, 1, 0] += 2 * Kdr * x * y;      G[:, 1, 1] += Kr + 2 * Kdr * y * y
    if nP:
        T = np.zeros((N, 2, 2))                               # :207-209
        T[:, 0, 0] = P[0] * 6 * x + P[1] * 2 * y; T[:, 1, 0] = P[0] * 2 * y + P[1] * 2 * x
        T[:, 0, 1] = P[0] * 2 * y + P[1] * 2 * x; T[:, 1, 1] = P[0] * 2 * x + P[1] * 6 * y
        if nP > 2:
            Sdr = (np.power(r2[:, None], ks[None, :] - 1) * ks[None, :]) @ P[2:]
            T = T * (1 + Sr)[:, None, None] + AwQ[:, :, None] * (w * 2 * Sdr).T[:, None, :]   # :211-224
        G += T
    out['dw'] = G
    out['dpp'] = -G                                           # :238-243
    return d, out


def brown_euler_cam4_legacy(x, s, IO, EO, OP, dm, want_jac):
    nK, nP = s.IO.model.nK, s.IO.model.nP
    ixm = vis_matrix(s)
    nIP = len(s.IP.img)
    pp, f, K, P, _ = _unpackio(IO[:, 0], nK, nP)              # single camera (:39-41,186-188)
    estIO = s.bundle.est.IO[:, 0]
    cpp, cf, cK, cP, _ = _unpackio(estIO, nK, nP)
    # measured points in mm, y flipped (:50, :197): per-axis pixel size of camera 1
    m = np.array([[1.0], [-1.0]]) * (s.IO.sensor.pxSize[:, [0]] * s.IP.val)
    xy = np.full((2, nIP), np.nan)
    blocks = {}
    for i in range(EO.shape[1]):                              # multieulerpinhole.m:110-176
        lo, hi = ixm.indptr[i], ixm.indptr[i + 1]
        if hi == lo:
            continue
        v = ixm.indices[lo:hi]
        cp = ixm.data[lo:hi] - 1
        p, d = eulerpinhole(pp, f, OP[:, v], EO[0:3, i], EO[3:6, i], want_jac)
        xy[:, cp] = p
        if want_jac:
            blocks[i] = (v, cp, d)
    fPre, JPre = prior_obs(x, s, want_jac)
    if dm == 1:
        ld, dl = browndist(m, pp, K, P, want_jac)             # :53-59
        fObs = xy - (m - ld)
    else:
        ld, dl = browndist(xy, pp, K, P, want_jac)            # :200-204
        fObs = xy + ld - m
    f_all = np.concatenate([fObs.reshape(-1, order='F')] + fPre)
    if not want_jac:
        return f_all, None
    dIOc, dEOc, dOPc = dest_cols(s)
    ioix = dIOc[:, 0]
    ppIx, fIx, Kix, Pix, _ = _unpackio(ioix, nK, nP)
    rows, cols, vals = [], [], []

    def pack(blk, rr, colix):
        if blk.shape[2] == 0:
            return
        R = np.broadcast_to(rr[:, :, None], blk.shape)
        C = np.broadcast_to(np.asarray(colix)[None, None, :], blk.shape)
        mnz = (blk != 0) & (C >= 0)
        rows.append(R[mnz]); cols.append(C[mnz]); vals.append(blk[mnz])

    rr_all = np.arange(2 * nIP).reshape(nIP, 2)
    I2 = np.broadcast_to(np.eye(2), (nIP, 2, 2))
    if dm == 1:
        # dIO = dIO1 (projection: pp, f) + dIO2 (distortion of the measured points: pp, K, P) (:106)
        if cpp.any():
            pack((I2 + dl['dpp'])[:, :, cpp], rr_all, ppIx[cpp])
        if cK.any():
            pack(dl['dK'][:, :, cK[:dl['dK'].shape[2]]], rr_all, Kix[cK])
        if cP.any():
            pack(dl['dP'][:, :, cP[:dl['dP'].shape[2]]], rr_all, Pix[cP])
        IG = None
    else:
        # forward model (:238-281): pp column from the projection only, K/P from the distortion,
        # f, EO, OP through (I + G)
        IG = I2 + dl['dw']
        if cpp.any():
            pack(np.array(I2)[:, :, cpp], rr_all, ppIx[cpp])
        if cK.any():
            pack(dl['dK'][:, :, cK[:dl['dK'].shape[2]]], rr_all, Kix[cK])
        if cP.any():
            pack(dl['dP'][:, :, cP[:dl['dP'].shape[2]]], rr_all, Pix[cP])
    for i, (v, cp, d) in blocks.items():
        rr = rr_all[cp]
        A = IG[cp] if IG is not None else None
        app = (lambda B: A @ B) if A is not None else (lambda B: B)
        if cf:
            pack(app(d['df'][:, :, None]), rr, [fIx])
        cEO = s.bundle.est.EO[0:6, i]
        if cEO.any():
            blk = np.concatenate([app(d['dC'])[:, :, cEO[0:3]], app(d['dAng'])[:, :, cEO[3:6]]], axis=2)
            pack(blk, rr, dEOc[0:6, i][cEO])
        cOP = s.bundle.est.OP[:, v]
        if cOP.any():
            blk = app(d['dO'])
            R = np.broadcast_to(rr[:, :, None], blk.shape)
            C = np.broadcast_to(dOPc[:, v].T[:, None, :], blk.shape)
            mnz = (blk != 0) & np.broadcast_to(cOP.T[:, None, :], blk.shape)
            rows.append(R[mnz]); cols.append(C[mnz]); vals.append(blk[mnz])
    ii = np.concatenate(rows) if rows else np.zeros(0, int)
    jj = np.concatenate(cols) if rows else np.zeros(0, int)
    vv = np.concatenate(vals) if rows else np.zeros(0)
    J = sp.coo_matrix((vv, (ii, jj)), shape=(2 * nIP, s.bundle.serial.n)).tocsc()
    J.sum_duplicates()
    J.eliminate_zeros()
    J = sp.vstack([J] + JPre, format='csc')
    J.sort_indices()
    return f_all, J
