"""Residual + analytic Jacobian of the modular camera models 2-5 — oracle restatement.

TEST INFRASTRUCTURE.  Follows `code/bundle/cameramodel/`:
`brown_euler_cam4.m:122-183`, `multi_res.m:14-347`, `res_euler_brown_{0,1,2,3}.m`,
and the building blocks `eulerpinhole2.m:50-108`, `eulerrotmat.m:57-124`,
`world2cam.m:45-84`, `pinhole.m:39-67`, `scale2.m:41`, `aniscale2.m:43`,
`aniscale2b.m:41`, `xlat2.m:41`, `affine2.m:42`, `affine2mat.m`, `skew.m:41`,
`brown_dist.m:50-91`, `brown_rad.m:46-95`, `brown_tang.m:58-138`, `rad_scale.m:43-77`,
`tang_scale.m:43-89`, `lens_rad2.m:39`, `power_vec.m:42-68`, `private/unpackio.m:4-8`.

The reference builds 2N x 2N sparse block-diagonal Jacobians per camera and multiplies
them; here each block-diagonal is held as an (N,2,k) dense array and the same chain-rule
products are taken per observation (identical arithmetic, no sparse bookkeeping).
"""
import numpy as np
import scipy.sparse as sp

from .dbatstruct import deserialize, vis_matrix


# ----------------------------------------------------------------------------- rotation
def _R1(a):
    c, s = np.cos(a), np.sin(a)
    return (np.array([[1, 0, 0], [0, c, -s], [0, s, c]]),
            np.array([[0, 0, 0], [0, 0, -1], [0, 1, 0.]]))


def _R2(a):
    c, s = np.cos(a), np.sin(a)
    return (np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]]),
            np.array([[0, 0, 1], [0, 0, 0], [-1, 0, 0.]]))


def _R3(a):
    c, s = np.cos(a), np.sin(a)
    return (np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]]),
            np.array([[0, -1, 0], [1, 0, 0], [0, 0, 0.]]))


def eulerrotmat123(ang):
    """eulerrotmat.m:57-124 with seq=123, fixed=false: M=R1*R2*R3 and dM/d(ang_k)."""
    M1, P1 = _R1(ang[0])
    M2, P2 = _R2(ang[1])
    M3, P3 = _R3(ang[2])
    M = M1 @ M2 @ M3
    dA = [P1 @ M, M1 @ M2 @ P2 @ M3, M @ P3]          # :109-113
    return M, dA


# ----------------------------------------------------------------------------- projection
def eulerpinhole2(Pw, p0, ang, f, want_jac):
    """eulerpinhole2.m:50-108: Q = f*pinhole(M'(P-p0)); Jacobians w.r.t. P, p0, ang, f."""
    M, dM = eulerrotmat123(ang)
    MT = M.T                                           # :57 (transpose = row permutation)
    X = Pw - p0[:, None]                               # xlat3.m:41
    W2C = MT @ X                                       # lin3.m:41
    PH = W2C[0:2] / W2C[2]                             # pinhole.m:39
    Q = f * PH
    if not want_jac:
        return Q, None
    N = Pw.shape[1]
    zi = 1.0 / W2C[2]
    dPH = np.zeros((N, 2, 3))                          # pinhole.m:53-67
    dPH[:, 0, 0] = zi
    dPH[:, 1, 1] = zi
    dPH[:, 0, 2] = -W2C[0] * zi ** 2
    dPH[:, 1, 2] = -W2C[1] * zi ** 2
    d = {}
    d['dF'] = PH.T.copy()                              # :95  (N,2)
    d['dP'] = f * (dPH @ MT)                           # :103 (N,2,3)
    d['dP0'] = f * (dPH @ (-MT))                       # :106, world2cam.m:83
    dA = np.empty((N, 2, 3))                           # :98-100: f*dPH*kron(X',I)*dMT.dA
    for k in range(3):
        dA[:, :, k] = f * np.einsum('nij,jn->ni', dPH, dM[k].T @ X)
    d['dA'] = dA
    return Q, d


# ----------------------------------------------------------------------------- distortion
def power_vec(x, nn):
    """power_vec.m:42,56-68: rows x.^k, k=1..nn (uses pow like MATLAB `.^`)."""
    k = np.arange(1, nn + 1)[:, None]
    v = np.power(x[None, :], k) if nn else np.zeros((0, len(x)))
    dv = k * np.power(x[None, :], k - 1) if nn else np.zeros((0, len(x)))
    return v, dv


def rad_scale(u, c):
    """rad_scale.m:43-77: v=sum_k c_k r^(2k); dC = r^(2k); dU = (sum k c_k r^(2k-2)) * 2u'."""
    r2 = np.sum(u ** 2, axis=0)                        # lens_rad2.m:39
    pv, dpv = power_vec(r2, len(c))
    v = c @ pv if len(c) else np.zeros(u.shape[1])
    dC = pv.T                                          # (N,nC)
    g = c @ dpv if len(c) else np.zeros(u.shape[1])
    dU = (g[None, :] * 2 * u).T                        # (N,2) row vector per point
    return v, dC, dU


def tang_scale(u, p):
    """tang_scale.m:43-89."""
    uTu = np.sum(u ** 2, axis=0)
    pTu = p @ u
    v = p[:, None] * uTu + 2 * pTu * u
    N = u.shape[1]
    dP = np.empty((N, 2, 2))                           # :66-73
    dP[:, 0, 0] = uTu + 2 * u[0] ** 2
    dP[:, 1, 0] = 2 * u[0] * u[1]
    dP[:, 0, 1] = 2 * u[0] * u[1]
    dP[:, 1, 1] = uTu + 2 * u[1] ** 2
    dU = np.empty((N, 2, 2))                           # :76-87
    dU[:, 0, 0] = 2 * (2 * p[0] * u[0] + pTu)
    dU[:, 1, 0] = 2 * (p[0] * u[1] + p[1] * u[0])
    dU[:, 0, 1] = 2 * (p[0] * u[1] + p[1] * u[0])
    dU[:, 1, 1] = 2 * (2 * p[1] * u[1] + pTu)
    return v, dP, dU


def brown_rad(u, K):
    """brown_rad.m:46-95."""
    rs, dC, drsU = rad_scale(u, K)
    v = u * rs
    dK = u.T[:, :, None] * dC[:, None, :]              # :74-78 (N,2,nK)
    dU = u.T[:, :, None] * drsU[:, None, :]            # :82-93 u*drs.dU + rs*I
    dU[:, 0, 0] += rs
    dU[:, 1, 1] += rs
    return v, dK, dU


def brown_tang(u, P):
    """brown_tang.m:58-138."""
    N = u.shape[1]
    if len(P) == 0:
        return np.zeros((2, N)), np.zeros((N, 2, 0)), np.zeros((N, 2, 2))
    ts, dtsP, dtsU = tang_scale(u, P[0:2])
    if len(P) <= 2:
        return ts, dtsP, dtsU
    rs, dC, drsU = rad_scale(u, P[2:])
    v = ts * (1 + rs)
    dPt = (1 + rs)[:, None, None] * dtsP               # :96-97
    dPr = ts.T[:, :, None] * dC[:, None, :]            # :98-100
    dP = np.concatenate([dPt, dPr], axis=2)
    dU = dtsU * (1 + rs)[:, None, None] + ts.T[:, :, None] * drsU[:, None, :]   # :108-134
    return v, dP, dU


def brown_dist(u, K, P):
    """brown_dist.m:50-91: v = u + brown_rad + brown_tang; dU = I + ..."""
    br, dK, dbrU = brown_rad(u, K)
    bt, dP, dbtU = brown_tang(u, P)
    v = u + br + bt
    dU = np.eye(2)[None] + dbrU + dbtU
    return v, dK, dP, dU


# ----------------------------------------------------------------------------- residual functions
def res_euler_brown(model, Q, q0, ang, f, u, sz, u0, K, P, b, want_jac=True):
    """res_euler_brown_{0,1,2,3}.m (distModel = model+2): v (2,N) and dict of per-obs blocks.

    Bodies: `_0:78-91,138-167`, `_1:81-97,147-179`, `_2`, `_3` (same positions).
    """
    lhs, dl = eulerpinhole2(Q, q0, ang, -f, want_jac)
    N = Q.shape[1]
    s = sz * u                                          # scale2.m:41
    y = np.array([1.0, -1.0])[:, None] * s              # aniscale2.m:43
    I2 = np.eye(2)
    if model == 0:
        x = y - u0[:, None]                             # xlat2.m:41 with c=-u0
        l, dLK, dLP, D = brown_dist(x, -K, -P)
        v = lhs - l
        if not want_jac:
            return v, None
        dU0 = D                                         # _0:155 dL.dU*dX.dC
        dK, dP, dB = dLK, dLP, np.zeros((N, 2, 0))
    elif model == 1:
        x = y - u0[:, None]
        A = np.array([[1 + b[0], b[1]], [0, 1.0]])      # affine2mat.m
        a = A @ x                                       # affine2.m:42
        l, dLK, dLP, D = brown_dist(a, -K, -P)
        v = lhs - l
        if not want_jac:
            return v, None
        dU0 = D @ A                                     # _1:168
        dK, dP = dLK, dLP
        dAB = np.zeros((N, 2, 2))                       # affine2.m dB(1:2:end,:)=U'
        dAB[:, 0, :] = x.T
        dB = -(D @ dAB)                                 # _1:177
    elif model == 2:
        x = y - u0[:, None]
        l, dLK, dLP, D = brown_dist(x, -K, -P)
        A = np.array([[1 + b[0], b[1]], [0, 1.0]])
        a = A @ l
        v = lhs - a
        if not want_jac:
            return v, None
        dU0 = A @ D                                     # _2 dA.dU*dL.dU*dX.dC
        dK = A @ dLK
        dP = A @ dLP
        dB = np.zeros((N, 2, 2))
        dB[:, 0, :] = -l.T                              # -dA.dB
    elif model == 3:
        ys = np.array([1 + b[0], 1.0])[:, None] * y     # aniscale2b.m:41
        x = ys - u0[:, None]
        l, dLK, dLP, D = brown_dist(x, -K, -P)
        SK = np.array([[1.0, b[1]], [0, 1.0]])          # skew.m:41
        sk = SK @ l
        v = lhs - sk
        if not want_jac:
            return v, None
        dU0 = SK @ D
        dK = SK @ dLK
        dP = SK @ dLP
        dASK = np.zeros((N, 2, 1))                      # aniscale2b dK(1:2:end)=U(1,:)'
        dASK[:, 0, 0] = y[0]
        dSKK = np.zeros((N, 2, 1))                      # skew dK(1:2:end)=U(2,:)'
        dSKK[:, 0, 0] = l[1]
        dB = -np.concatenate([SK @ D @ dASK, dSKK], axis=2)
    else:
        raise ValueError('bad model')
    d = {'dQ': dl['dP'], 'dQ0': dl['dP0'], 'dA': dl['dA'], 'dF': -dl['dF'],
         'dU0': np.broadcast_to(dU0, (N, 2, 2)), 'dK': dK, 'dP': dP, 'dB': dB}
    return v, d


# ----------------------------------------------------------------------------- multi_res
def _trimkp(K, first_is_pair):
    """multi_res.m:318-340."""
    nz = np.flatnonzero(K)
    if len(nz) == 0:
        return K[:0]
    i = nz[-1] + 1
    if first_is_pair and i == 1:
        i = 2
    return K[:i]


def _unpackio(col, nK, nP):
    """private/unpackio.m:4-8: rows [f; pp(2); b(2); K(nK); P(nP)]."""
    return col[1:3], col[0], col[5:5 + nK], col[5 + nK:5 + nK + nP], col[3:5]


def dest_cols(s):
    """multi_res.m:58-63: column in x of every IO/EO/OP element (-1 = fixed)."""
    def mk(shape, des):
        d = np.full(shape[0] * shape[1], -1, dtype=np.int64)
        d[des.dest] = des.src
        return d.reshape(shape, order='F')
    return (mk(s.IO.val.shape, s.bundle.deserial.IO), mk(s.EO.val.shape, s.bundle.deserial.EO),
            mk(s.OP.val.shape, s.bundle.deserial.OP))


def multi_res(s, IO, EO, OP, model, want_jac):
    """multi_res.m:14-315.  Returns r (2*nProj,) and (if asked) CSC J (2*nProj x n).

    Triplets are packed exactly as the reference does, including `find()` dropping
    exact zeros (:150-294) and `sparse()` summing duplicates/dropping zeros (:313).
    """
    nK, nP = s.IO.model.nK, s.IO.model.nP
    nIP = len(s.IP.img)
    xy = np.full((2, nIP), np.nan)
    ixm = vis_matrix(s)
    if want_jac:
        dIO, dEO, dOP = dest_cols(s)
        rows, cols, vals = [], [], []
        oprows, opcols, opvals = [], [], []
        xyJ = np.full(2 * nIP, np.nan)
        jacLast = 0
    for i in range(EO.shape[1]):
        lo, hi = ixm.indptr[i], ixm.indptr[i + 1]
        if hi == lo:
            continue                                    # find(any(s.IP.vis))
        v = ixm.indices[lo:hi]                          # OP rows seen, ascending
        cp = ixm.data[lo:hi] - 1                        # IP columns
        pp, f, K, P, b = _unpackio(IO[:, i], nK, nP)
        sz = s.IO.sensor.pxSize[0, i]                   # :138 uses sz(1)
        cpp, cf, cK, cP, cb = _unpackio(s.bundle.est.IO[:, i], nK, nP)
        if not (want_jac and cK.any()):
            K = _trimkp(K, False)                       # :105-111 / :37-38
        if not (want_jac and cP.any()):
            P = _trimkp(P, True)
        res, d = res_euler_brown(model, OP[:, v], EO[0:3, i], EO[3:6, i], f,
                                 s.IP.val[:, cp], sz, pp, K, P, b, want_jac)
        xy[:, cp] = res                                 # :52
        if not want_jac:
            continue
        N = len(cp)
        ppIx, fIx, Kix, Pix, bIx = _unpackio(dIO[:, i], nK, nP)
        xyJ[jacLast:jacLast + 2 * N] = res.T.ravel()    # :143-144
        rr = jacLast + np.arange(2 * N).reshape(N, 2)   # row of (obs, xy)

        def pack(blk, colix, rl, cl, vl):
            # blk (N,2,k) dense; emulate [ii,jj,vv]=find(blk2d) (drops zeros)
            k = blk.shape[2]
            if k == 0:
                return
            R = np.broadcast_to(rr[:, :, None], blk.shape)
            C = np.broadcast_to(np.asarray(colix)[None, None, :], blk.shape)
            m = blk != 0
            rl.append(R[m]); cl.append(C[m]); vl.append(blk[m])

        if cpp.any():                                   # :148-166
            pack(d['dU0'][:, :, cpp], ppIx[cpp], rows, cols, vals)
        if cf:                                          # :168-179
            pack(d['dF'][:, :, None], [fIx], rows, cols, vals)
        if cK.any():                                    # :181-203
            pack(d['dK'][:, :, cK[:d['dK'].shape[2]]], Kix[cK], rows, cols, vals)
        if cP.any():                                    # :205-226
            pack(d['dP'][:, :, cP[:d['dP'].shape[2]]], Pix[cP], rows, cols, vals)
        if cb.any():                                    # :228-242
            nb = d['dB'].shape[2]
            pack(d['dB'][:, :, cb[:nb]], bIx[:nb][cb[:nb]], rows, cols, vals)
        cEO = s.bundle.est.EO[0:6, i]
        if cEO.any():                                   # :245-273
            blk = np.concatenate([d['dQ0'][:, :, cEO[0:3]], d['dA'][:, :, cEO[3:6]]], axis=2)
            pack(blk, dEO[0:6, i][cEO], rows, cols, vals)
        cOP = s.bundle.est.OP[:, v]                     # (3,N)
        if cOP.any():                                   # :275-294 per-obs 2x3 blocks
            blk = d['dQ']                               # (N,2,3)
            R = np.broadcast_to(rr[:, :, None], blk.shape)
            C = np.broadcast_to(dOP[:, v].T[:, None, :], blk.shape)
            m = (blk != 0) & np.broadcast_to(cOP.T[:, None, :], blk.shape)
            oprows.append(R[m]); opcols.append(C[m]); opvals.append(blk[m])
        jacLast += 2 * N
    r = xy.reshape(-1, order='F')
    if not want_jac:
        return r, None
    ii = np.concatenate(rows + oprows) if rows or oprows else np.zeros(0, int)
    jj = np.concatenate(cols + opcols) if rows or oprows else np.zeros(0, int)
    vv = np.concatenate(vals + opvals) if rows or oprows else np.zeros(0)
    J = sp.coo_matrix((vv, (ii, jj)), shape=(2 * nIP, s.bundle.serial.n)).tocsc()
    J.sum_duplicates()
    J.eliminate_zeros()
    J.sort_indices()
    return xyJ, J


def prior_obs(x, s, want_jac):
    """lsa/prior_obs.m:28-65: f = x(dest(obs)) - prior.val(src(obs)); J rows = unit vectors."""
    out_f, out_J = [], []
    for ser, pri in ((s.bundle.serial.IO, s.prior.IO), (s.bundle.serial.EO, s.prior.EO),
                     (s.bundle.serial.OP, s.prior.OP)):
        dest = ser.dest[ser.obs]
        f = x[dest] - pri.val.reshape(-1, order='F')[ser.src[ser.obs]]
        out_f.append(f)
        if want_jac:
            out_J.append(sp.csc_matrix((np.ones(len(f)), (np.arange(len(f)), dest)),
                                       shape=(len(f), len(x))))
    return out_f, out_J


def brown_euler_cam4(x, s, want_jac=True):
    """brown_euler_cam4.m:22-183 for distModel 2..5: stitch image and prior rows."""
    IO, EO, OP = deserialize(s, x)
    dm = np.unique(s.IO.model.distModel)
    if len(dm) != 1:
        raise ValueError('Mixed lens distortion models not implemented.')   # :30-33
    dm = int(dm[0])
    if dm in (1, -1):
        from .cammodel_legacy import brown_euler_cam4_legacy
        return brown_euler_cam4_legacy(x, s, IO, EO, OP, dm, want_jac)
    if dm not in (2, 3, 4, 5):
        raise ValueError('Bad distortion model %d' % dm)
    fObs, JObs = multi_res(s, IO, EO, OP, dm - 2, want_jac)
    fPre, JPre = prior_obs(x, s, want_jac)
    f = np.concatenate([fObs] + fPre)                   # ix.IP, ix.IO, ix.EO, ix.OP are consecutive
    if not want_jac:
        return f, None
    J = sp.vstack([JObs] + JPre, format='csc')
    J.sort_indices()
    return f, J
