"""bundle() driver and bundle_cov() — oracle restatement (test infrastructure).

Follows `code/bundle/bundle.m:78-192,267-358,449-491` and
`code/bundle/bundle_cov.m:63-117,135-214,221-478`, `private/invblock.m:29-41,86-90`,
`misc/mkblkdiag.m:49-60`.
"""
import time
from types import SimpleNamespace as NS

import numpy as np
import scipy.linalg as sla
import scipy.sparse as sp

from . import lsa
from .cameramodel import brown_euler_cam4
from .dbatstruct import (buildserialindices, buildweightmatrix, deserialize, dmperm_match, paramtypes,
                         serialize)


def bundle(s, damping='gna', maxIter=20, convTol=1e-6, absTerm=False, doTrace=False,
           singularTest=True, pmDof=False):
    """bundle.m:1-491 → (s, ok, iters, s0, E)."""
    for pri, est in ((s.prior.IO, s.bundle.est.IO), (s.prior.EO, s.bundle.est.EO),
                     (s.prior.OP, s.bundle.est.OP)):                 # :137-154
        pri.use[~est] = False
    if s.bundle.serial is None or s.bundle.deserial is None:         # :156-159
        buildserialindices(s)
    x0 = serialize(s)                                                # :162
    nOPx = len(s.bundle.serial.OP.dest)                              # ordering hint, see lsa._PERM
    lsa.set_ordering(np.concatenate([np.arange(len(x0) - nOPx, len(x0)),
                                     np.arange(len(x0) - nOPx)[::-1]]))
    resFun = lambda x, want_jac=True: brown_euler_cam4(x, s, want_jac)   # :165
    vetoFun = None
    W = buildweightmatrix(s)                                         # :175
    if absTerm:                                                      # :186-192
        termFun = lambda Jp, r: np.linalg.norm(r) <= convTol
    else:
        termFun = lambda Jp, r: np.linalg.norm(Jp) <= convTol * np.linalg.norm(r)
    E = NS(maxIter=maxIter, convTol=convTol, absTerm=absTerm, singularTest=singularTest)
    t0 = time.process_time()
    damping = damping.lower()
    if damping == 'gna':                                             # :275-295
        mu, alphaMin = 0.1, 1e-9
        x, code, iters, final, X, res, alpha = lsa.gauss_newton_armijo(
            resFun, vetoFun, x0, W, maxIter, termFun, doTrace, singularTest, mu, alphaMin)
        E.damping = NS(name='gna', alpha=alpha, mu=mu, alphaMin=alphaMin)
    elif damping == 'lm':                                            # :296-315
        lambda0 = -1e-10
        lambdaMin = lambda0
        x, code, iters, final, X, res, lam = lsa.levenberg_marquardt(
            resFun, vetoFun, x0, W, maxIter, termFun, doTrace, lambda0, lambdaMin)
        E.damping = NS(name='lm', **{'lambda': lam}, lambda0=lam[0], lambdaMin=lam[0])
    elif damping == 'lmp':                                           # :316-335
        rhoBad, rhoGood = 0.25, 0.75
        delta0 = np.linalg.norm(x0)
        x, code, iters, final, X, res, delta, rho, step = lsa.levenberg_marquardt_powell(
            resFun, vetoFun, x0, W, maxIter, termFun, doTrace, delta0, rhoBad, rhoGood)
        E.damping = NS(name='lmp', delta=delta, rho=rho, delta0=delta0, rhoBad=rhoBad,
                       rhoGood=rhoGood, step=step)
    else:
        raise ValueError('Unknown damping')
    E.time = time.process_time() - t0
    E.res, E.trace, E.code, E.usedIters, E.final = res, X, code, iters, final   # :341-348
    E.final.factorized = None
    E.x = x
    ok = code == 0
    if ok:                                                           # :356-358
        IO, EO, OP = deserialize(s, x)
        s.IO.val, s.EO.val, s.OP.val = IO, EO, OP
    # residuals (:449-462): IP residuals mm -> px
    resIP = final.unweighted.r[s.post.res.ix.IP].reshape(2, -1, order='F')
    s.post.res.IP = resIP / s.IO.sensor.pxSize[:, s.IP.cam]
    p = 0
    if pmDof:                                                        # :470-472
        seen_op = np.zeros(s.OP.val.shape[1], bool); seen_op[s.IP.op] = True
        seen_im = np.zeros(s.EO.val.shape[1], bool); seen_im[s.IP.img] = True
        p = np.count_nonzero(~s.bundle.est.OP[:, seen_op]) + \
            np.count_nonzero(~s.bundle.est.EO[0:6, seen_im])
    r = E.final.weighted.r                                           # :476-491
    dof = len(r) + p - len(x)
    s0 = np.sqrt((r @ r) / dof)
    s.post.sigmas = s0 * s.IP.sigmas
    E.numObs, E.numParams, E.redundancy, E.s0 = len(r), len(x), dof, s0
    # bundle.m:368-446: parameter type of every unknown; on code -4 the structural analysis
    E.paramTypes = paramtypes(s)
    E.weakness = NS(structural=None, numerical=None)
    if code == -4:
        dm = dmperm_match(E.final.weighted.J)                        # :433
        rank = int(np.count_nonzero(dm))
        E.weakness.structural = NS(dmperm=dm, rank=rank, deficiency=len(x) - rank,
                                   suspectedParams=list(E.paramTypes[dm == 0]))
        E.weakness.numerical = NS(rank=np.nan, deficiency=np.nan)    # :444-445
    elif code == -2 and hasattr(E.final, 'scaled'):                  # :373-428
        E.weakness.numerical = numerical_weakness(E.final.scaled.J, E.paramTypes)
    else:                                                            # :429-432
        E.weakness.numerical = NS(rank=len(x), deficiency=0)
    return s, ok, iters, s0, E


def numerical_weakness(Js, paramTypes):
    """bundle.m:373-428: numerical rank of the column-scaled Jacobian and the parameters that carry
    its null-space.  The rank follows spnrank.m:166-178 (singular values above
    max(size)*eps(smax)); dense SVD / eigh stand in for the sparse eigs calls, so the null-space
    basis V is one orthonormal basis of the same invariant subspace, ordered by |eigenvalue|."""
    A = Js.toarray() if sp.issparse(Js) else np.asarray(Js)
    n = A.shape[1]
    sv = np.linalg.svd(A, compute_uv=False)
    tol = max(A.shape) * np.spacing(sv[0])                           # spnrank.m:177
    rank = int(np.count_nonzero(sv > tol))
    W = NS(rank=rank, deficiency=n - rank)
    if W.deficiency > 0:
        JTJ = A.T @ A
        d, V = np.linalg.eigh(JTJ)                                   # :398-402 (shift cancels)
        k = np.argsort(np.abs(d), kind='stable')[:W.deficiency]      # :405-407
        d, V = d[k], V[:, k]
        W.V, W.d = V, d
        W.trace = float(np.trace(JTJ) + np.sqrt(np.finfo(float).eps) * n)   # :410 (shifted JTJ)
        W.suspectedParams = []
        pt = np.asarray(paramTypes, dtype=object)
        avg = np.sqrt(1.0 / n)
        for j in range(V.shape[1]):                                  # :412-422
            o = np.argsort(-np.abs(V[:, j]), kind='stable')
            v = V[o, j]
            keep = np.abs(v) > 0.5 * (avg + abs(v[0]))
            W.suspectedParams.append(NS(values=v[keep], indices=o[keep], params=list(pt[o[keep]])))
    return W


# --------------------------------------------------------------------------- bundle_cov
def _prepare(s, e):
    """bundle_cov.m:63-117: permute N to [OP;EO;IO], Cholesky, split L=[LA 0;LB LC]."""
    J = e.final.weighted.J
    JTJ = (J.T @ J).toarray()
    def bix(ser, shape):
        b = np.full(shape[0] * shape[1], -1, dtype=np.int64)
        b[ser.src] = ser.dest
        return b
    bOP = bix(s.bundle.serial.OP, s.bundle.est.OP.shape)
    bEO = bix(s.bundle.serial.EO, s.bundle.est.EO.shape)
    bIO = bix(s.bundle.serial.IO, s.bundle.est.IO.shape)
    p = np.concatenate([bOP, bEO, bIO])
    p = p[p >= 0]                                                    # :83-84
    nOP = np.count_nonzero(bOP >= 0)
    try:
        if not np.all(np.isfinite(JTJ)):
            raise np.linalg.LinAlgError
        L = np.linalg.cholesky(JTJ[np.ix_(p, p)])                    # :87
        fail = False
    except np.linalg.LinAlgError:
        L = np.full(JTJ.shape, np.nan)
        fail = True
    e.final.factorized = NS(p=p, L=L, nOP=nOP, fail=fail)
    return e


def _invblock_sqrt(L, p, ix):
    """invblock.m:29-41 ('sqrt'): v2 = L\\Ip(:,ix); x = v2'*v2."""
    invP = np.empty(len(p), dtype=np.int64)
    invP[p] = np.arange(len(p))
    n = L.shape[0]
    rhs = np.zeros((n, len(ix)))
    rhs[invP[ix], np.arange(len(ix))] = 1.0
    v2 = sla.solve_triangular(L, rhs, lower=True, check_finite=False)
    return v2.T @ v2


BLOCK_PATH_ABOVE = 6000        # unknowns; beyond it the dense n x n factor does not fit


def _prepare_blocks(s, e):
    """bundle_cov.m:87-101 and VectorizedCOP (:330-478) for problems too large for a dense factor: the
    same elimination order [OP; EO; IO], carried out on the blocks.  With N = [A B'; B D] (A the
    block-diagonal OP part), L = [LA 0; LB LC] gives LC*LC' = D - B*inv(A)*B' and
    inv(N) = [inv(A) + Y'*Cc*Y, -Y'*Cc; -Cc*Y, Cc],  Y = B*inv(A), Cc = inv(LC*LC')."""
    J = e.final.weighted.J.tocsc()
    N = (J.T @ J).tocsc()
    opx = np.asarray(s.bundle.serial.OP.dest)
    src = np.asarray(s.bundle.serial.OP.src)
    pt, comp = src // 3, src % 3
    camx = np.concatenate([np.asarray(s.bundle.serial.EO.dest), np.asarray(s.bundle.serial.IO.dest)])
    nOP = s.OP.val.shape[1]
    A = N[opx][:, opx].tocoo()
    assert np.all(pt[A.row] == pt[A.col])
    Ab = np.tile(np.eye(3), (nOP, 1, 1))
    Ab[pt, comp, comp] = 0.0
    np.add.at(Ab, (pt[A.row], comp[A.row], comp[A.col]), A.data)
    try:
        Ai = np.linalg.inv(Ab)
        Aix = sp.csc_matrix(_gather_blockdiag(Ai, pt, comp), shape=(len(opx), len(opx)))
        B = N[camx][:, opx].tocsc()
        Y = (B @ Aix).tocsc()
        S = N[camx][:, camx].toarray() - (Y @ B.T).toarray()
        Lc = np.linalg.cholesky(S)
        Li = sla.solve_triangular(Lc, np.eye(len(camx)), lower=True)
        Cc = Li.T @ Li
        fail = False
    except np.linalg.LinAlgError:
        Ai, Y, Cc, fail = None, None, None, True
    e.final.factorized = NS(blocks=True, fail=fail, Ai=Ai, Y=Y, Cc=Cc, camx=camx, opx=opx, pt=pt, comp=comp)
    return e


def _gather_blockdiag(Ai, pt, comp):
    """(data, (row, col)) of the block-diagonal matrix whose x-ordered rows are (pt, comp)."""
    order = np.argsort(pt, kind='stable')
    start = np.searchsorted(pt[order], np.arange(Ai.shape[0] + 1))
    rows, cols = [], []
    size = np.diff(start)
    for k in (1, 2, 3):
        for j0 in [np.flatnonzero(size == k)]:
            if len(j0) == 0:
                continue
            idx = order[start[j0][:, None] + np.arange(k)[None, :]]          # (n, k) x-positions
            rows.append(np.repeat(idx, k, axis=1).ravel())
            cols.append(np.tile(idx, (1, k)).ravel())
    r, c = np.concatenate(rows), np.concatenate(cols)
    return Ai[pt[r], comp[r], comp[c]], (r, c)


def _cov_blocks(s, e, w):
    F = e.final.factorized
    key = w[1:3].upper()
    shape = getattr(s.bundle.est, key).shape
    Nel = shape[0] * shape[1]
    if key in ('IO', 'EO'):
        des = getattr(s.bundle.deserial, key)
        C = np.zeros((Nel, Nel))
        if F.fail:
            C[np.ix_(des.dest, des.dest)] = np.nan
        else:
            pos = np.full(int(max(F.camx.max(), F.opx.max())) + 1, -1)
            pos[F.camx] = np.arange(len(F.camx))
            ix = pos[des.src]
            C[np.ix_(des.dest, des.dest)] = F.Cc[np.ix_(ix, ix)]
        if len(w) == 3:
            C = C * np.kron(np.eye(shape[1]), np.ones((shape[0], shape[0])))
        return C
    if w != 'cop':
        raise ValueError("bundle_cov: '%s' needs the dense factor (n > %d)" % (w, BLOCK_PATH_ABOVE))
    nOP = shape[1]
    Cb = np.zeros((nOP, 3, 3))
    if F.fail:
        Cb[F.pt[:, None], F.comp[:, None], F.comp[None, :]] = np.nan
    else:
        Cb[:] = F.Ai
        edges = np.searchsorted(F.pt, np.arange(0, nOP + 3000, 3000))   # chunks end on point boundaries
        for c0, c1 in zip(edges[:-1], edges[1:]):                    # Y'*Cc*Y, diagonal blocks only
            if c1 == c0:
                continue
            cols = np.arange(c0, c1)
            Yc = F.Y[:, cols].toarray()
            T = F.Cc @ Yc
            for d in (-2, -1, 0, 1, 2):                              # neighbours in x order share a point
                a = cols[max(0, -d):len(cols) - max(0, d)]
                b = a + d
                same = F.pt[a] == F.pt[b]
                v = np.einsum('ij,ij->j', Yc[:, a[same] - c0], T[:, b[same] - c0])
                Cb[F.pt[a[same]], F.comp[a[same]], F.comp[b[same]]] += v
        fixed = np.ones((nOP, 3), bool)
        fixed[F.pt, F.comp] = False
        Cb[fixed] = 0.0
        Cb[np.broadcast_to(fixed[:, None, :], Cb.shape)] = 0.0
    base = np.repeat(np.arange(nOP) * 3, 9)
    rr = base + np.tile(np.repeat(np.arange(3), 3), nOP)
    cc = base + np.tile(np.tile(np.arange(3), 3), nOP)
    return sp.csc_matrix((Cb.ravel(), (rr, cc)), shape=(3 * nOP, 3 * nOP))


def bundle_cov(s, e, *which, blocks=None):
    """bundle_cov.m:1-214.  which ∈ {'CXX','CIO','CEO','COP','CIOF','CEOF','COPF'}.

    Returns dense arrays (the reference returns sparse matrices of the same shape):
    CIOF (NC*nImg)^2, CEOF (6*nImg)^2, COPF (3*nOP)^2, CIO/CEO/COP block-diagonal of
    those, CXX n x n — each scaled by s0^2 (:213).  Above BLOCK_PATH_ABOVE unknowns (or with
    blocks=True) the factorisation is done on the blocks (`_prepare_blocks`): same numbers, COP comes
    back as a sparse block-diagonal matrix, COPF / CXX are not available.
    """
    if blocks is None:
        blocks = len(e.x) > BLOCK_PATH_ABOVE
    if blocks:
        if e.final.factorized is None or not getattr(e.final.factorized, 'blocks', False):
            _prepare_blocks(s, e)
        out = [e.s0 ** 2 * _cov_blocks(s, e, w.lower()) for w in which]
        return out[0] if len(out) == 1 else out
    if e.final.factorized is None or getattr(e.final.factorized, 'blocks', False):
        _prepare(s, e)
    F = e.final.factorized
    out = []
    for w in which:
        w = w.lower()
        if w == 'cxx':                                               # :138-145, invblock 'direct'
            C = _invblock_sqrt(F.L, F.p, np.arange(len(F.p)))
        elif w in ('ciof', 'ceof', 'copf', 'cio', 'ceo', 'cop'):
            key = w[1:3].upper()
            des = getattr(s.bundle.deserial, key)
            shape = getattr(s.bundle.est, key).shape
            N = shape[0] * shape[1]
            C = np.zeros((N, N))
            if F.fail:
                C[np.ix_(des.dest, des.dest)] = np.nan
            else:
                C[np.ix_(des.dest, des.dest)] = _invblock_sqrt(F.L, F.p, des.src)   # :148-196
            if len(w) == 3:                                          # block-diagonal (:198-211)
                mask = np.kron(np.eye(shape[1]), np.ones((shape[0], shape[0])))     # mkblkdiag.m
                C = C * mask
        else:
            raise ValueError(w)
        out.append(e.s0 ** 2 * C)
    return out[0] if len(out) == 1 else out
