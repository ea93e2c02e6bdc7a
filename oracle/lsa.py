"""Least-squares optimisers — oracle restatement (test infrastructure).

Follows `code/bundle/lsa/levenberg_marquardt.m:54-247`,
`levenberg_marquardt_powell.m:60-335`, `gauss_newton_armijo.m:75-290`.
Control flow, quirks included, is copied literally (LM's stale-Jp termination test
`:217`, duplicated first lambda `:106,136`; LMP's trace trimming `:226`).

Linear algebra that MATLAB delegates to CHOLMOD (`\\` on sparse SPD) is done with
SciPy SuperLU on the same explicitly formed sparse normal matrix; results agree with the
reference to solver rounding only.
"""
import warnings
from types import SimpleNamespace as NS

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla
from scipy.sparse.csgraph import structural_rank


# Fill-reducing ordering handed to the sparse factorisation.  MATLAB's `\\` lets CHOLMOD pick
# an AMD ordering; SuperLU's own orderings do poorly on this arrow structure, so bundle()
# installs the ordering the reference itself uses for its explicit Cholesky,
# p=[OP;EO;IO] (bundle_cov.m:73-84), which confines the fill to the camera block.
_PERM = None


def set_ordering(perm):
    global _PERM
    _PERM = None if perm is None else np.asarray(perm)


def _solve_spd(A, b):
    """MATLAB `A\\b` for sparse symmetric A.  Returns (x, singular_flag)."""
    A = sp.csc_matrix(A)
    perm = _PERM if (_PERM is not None and len(_PERM) == A.shape[0]) else None
    with warnings.catch_warnings():
        warnings.simplefilter('error', spla.MatrixRankWarning)
        try:
            if perm is not None:
                lu = spla.splu(A[perm][:, perm].tocsc(), permc_spec='NATURAL', diag_pivot_thresh=0.0,
                               options=dict(SymmetricMode=True))
                x = np.empty(len(b))
                x[perm] = lu.solve(b[perm])
            else:
                lu = spla.splu(A, permc_spec='MMD_AT_PLUS_A', diag_pivot_thresh=0.0,
                               options=dict(SymmetricMode=True))
                x = lu.solve(b)
        except (RuntimeError, spla.MatrixRankWarning):
            return np.full(len(b), np.nan), True
    if not np.all(np.isfinite(x)):
        return x, True
    # MATLAB warns (singular / nearlySingular) when its reciprocal condition estimate is below eps.
    # Cheap screen on U's diagonal first; in the suspicious band the 1-norm estimate itself
    # (Hager/Higham, what rcond/condest do) decides.  Pinned by camcal-dbatreport-no-datum.txt
    # (datum-free network: code -2 at iteration 0).
    eps = np.finfo(float).eps
    d = np.abs(lu.U.diagonal())
    ratio = d.min() / d.max()
    if ratio <= eps * 1e-2:
        return x, True
    if ratio < 1e-12:
        n = A.shape[0]
        if perm is not None:
            def solve(v):
                out = np.empty(n)
                out[perm] = lu.solve(np.asarray(v, dtype=float).reshape(-1)[perm])
                return out
        else:
            def solve(v):
                return lu.solve(np.asarray(v, dtype=float).reshape(-1))
        Ainv = spla.LinearOperator((n, n), matvec=solve, rmatvec=solve, dtype=float)
        rcond = 1.0 / (spla.onenormest(A) * spla.onenormest(Ainv))
        if rcond < eps:
            return x, True
    return x, False


def solve_spd_pointfirst(A, b, nC):
    """`A\\b` (levenberg_marquardt.m:119) for the damped normal matrix of a bundle with x = [IO; EO; OP] and
    every object point fully estimated: what MATLAB's mldivide -> CHOLMOD (supernodal, AMD ordering) does on this
    matrix, spelled out - the 3 x 3 point blocks are eliminated first (AMD puts the low-degree point columns
    first, as the reference itself orders them in bundle_cov.m:73-84), the camera front is a dense LAPACK
    Cholesky.  Used by bench.py's CPU baseline at sizes where SuperLU (no supernodal BLAS-3 front) needs tens of
    minutes; checked against `_solve_spd` in tests/test_oracle_golden.py."""
    import scipy.linalg as sla
    A = sp.csc_matrix(A)
    n = A.shape[0]
    m3 = n - nC
    assert m3 % 3 == 0
    nP = m3 // 3
    D = A[:nC, :nC].toarray()
    B = A[:nC, nC:].tocsc()                                   # camera x point
    V = A[nC:, nC:].tocoo()
    assert np.all(V.row // 3 == V.col // 3), 'point block is not 3 x 3 block diagonal'
    Vb = np.zeros((nP, 3, 3))
    np.add.at(Vb, (V.row // 3, V.row % 3, V.col % 3), V.data)
    Vi = np.linalg.inv(Vb)
    ii = (3 * np.arange(nP)[:, None, None] + np.arange(3)[None, :, None]) + np.zeros((1, 1, 3), dtype=np.int64)
    jj = (3 * np.arange(nP)[:, None, None] + np.arange(3)[None, None, :]) + np.zeros((1, 3, 1), dtype=np.int64)
    Vis = sp.csc_matrix((Vi.ravel(), (ii.ravel(), jj.ravel())), shape=(m3, m3))
    Y = (B @ Vis).tocsr()                                     # B V^-1
    S = D - (Y @ B.T).toarray()
    rhs = b[:nC] - Y @ b[nC:]
    c, low = sla.cho_factor(S, lower=True, check_finite=False)
    pc = sla.cho_solve((c, low), rhs, check_finite=False)
    pp = Vis @ (b[nC:] - B.T @ pc)
    return np.concatenate([pc, pp])


def _weighted(resFun, R, x, want_jac):
    s, K = resFun(x, want_jac)
    r = R * s
    J = sp.diags(R) @ K if want_jac else None
    return s, K, r, (sp.csc_matrix(J) if want_jac else None)


def levenberg_marquardt(resFun, vetoFun, x0, W, maxIter, termFun, doTrace, lambda0, lambdaMin):
    """levenberg_marquardt.m:54-247.  W is diag(W) (vector); resFun(x, want_jac)->(f,J)."""
    x = x0.copy()
    T = []
    n = 0
    code = 0
    R = np.sqrt(W)                                               # :70 chol(W), W diagonal
    s, K, r, J = _weighted(resFun, R, x, True)                   # :76-79
    f = 0.5 * r @ r
    JTJ = (J.T @ J).tocsc()
    JTr = J.T @ r
    rr = []
    nn = J.shape[1]
    if lambda0 < 0:                                              # :88-95
        lambda0 = abs(lambda0) * JTJ.diagonal().sum() / nn
    if lambdaMin < 0:
        lambdaMin = abs(lambdaMin) * JTJ.diagonal().sum() / nn
    lam = lambda0
    if lam < lambdaMin:
        lam = 0.0
    lambdas = [lam]                                              # :106
    prevLambda = np.nan
    I = sp.identity(nn, format='csc')
    Jp = None
    p = None
    while True:
        while n <= maxIter:                                      # :117
            p, _ = _solve_spd(JTJ + lam * I, -JTr)               # :119
            rr.append(np.sqrt(r @ r))                            # :122
            if n == 0:                                           # :126-135
                if structural_rank(J) < nn:
                    code = -4
                    p = np.full(len(x), np.nan)
                    break
            lambdas.append(lam)                                  # :136
            if doTrace:
                print('Levenberg-Marquardt: iteration %d, residual norm=%.2g, lambda=%.2g'
                      % (n, rr[-1], lam))
            T.append(x.copy())                                   # :149-156
            n += 1                                               # :159
            Jp = J @ p                                           # :162
            t = x + p
            rNew = R * resFun(t, False)[0]                       # :166-167
            fNew = 0.5 * rNew @ rNew
            fail = bool(vetoFun(t)) if (fNew < f and vetoFun) else False
            if fNew < f and not fail:                            # :177
                x = t
                lam = lam / 10                                   # :181
                if lam < lambdaMin:
                    lam = 0.0
                s, K, r, J = _weighted(resFun, R, x, True)       # :188-194
                f = 0.5 * r @ r
                JTJ = (J.T @ J).tocsc()
                JTr = J.T @ r
                break
            else:                                                # :198-206
                lam = lambdaMin if lam == 0 else lam * 10
        if code != 0:
            break
        if prevLambda == 0 and termFun(Jp, r):                   # :217
            break
        prevLambda = lam                                         # :222
        if n > maxIter:
            code = -1
            break
    final = NS(unweighted=NS(r=s, J=K), weighted=NS(r=r, J=J), p=p)
    T.append(x.copy())                                           # :238-240
    rr.append(np.sqrt(r @ r))                                    # :242
    T = np.array(T).T[:, :n + 1]
    return x, code, n, final, T, np.array(rr), np.array(lambdas)


def _scaled_gn(J, r):
    """gauss_newton_armijo.m:166-174 / levenberg_marquardt_powell.m:267-279."""
    Jn2 = np.asarray(J.multiply(J).sum(axis=0)).ravel()
    Jn = np.sqrt(Jn2)
    with np.errstate(divide='ignore'):
        D = 1.0 / Jn
    Js = (J @ sp.diags(D)).tocsc()
    Hs = (Js.T @ Js).tocsc()
    gs = Js.T @ r
    q, sing = _solve_spd(Hs, -gs)
    return D * q, sing, D, Js, Hs, gs, Jn, Jn2


def gauss_markov(resFun, x0, W, maxIter, convTol, trace, sTest):
    """gauss_markov.m:35-121: undamped Gauss-Newton.  (bundle.m:273 passes a function handle as convTol,
    so the method is only usable by a direct call; it is restated as written.)"""
    x = x0.copy()
    T = [x0.copy()]
    n = 0
    code = 0
    rr = []
    R = np.sqrt(W)
    while True:
        s, K, r, J = _weighted(resFun, R, x, True)               # :58-61
        rr.append(np.sqrt(r @ r))                                # :63
        if trace:
            print('Gauss-Markov: iteration %d, residual norm=%.2g' % (n, rr[-1]))
        p, sing = _solve_spd((J.T @ J).tocsc(), -(J.T @ r))      # :69
        if sTest and sing:                                       # :71-79
            code = -2
            break
        if np.linalg.norm(J @ p) <= convTol * np.linalg.norm(r): # :86
            break
        n += 1                                                   # :92
        x = x + p                                                # :95
        T.append(x.copy())                                       # :97-104
        if n > maxIter:                                          # :107-110
            code = -1
            break
    final = NS(unweighted=NS(r=s, J=K), weighted=NS(r=r, J=J))   # :113-116
    return x, code, n, final, np.array(T).T[:, :n + 1], np.array(rr)


def gauss_newton_armijo(resFun, vetoFun, x0, W, maxIter, termFun, trace, sTest, mu, alphaMin):
    """gauss_newton_armijo.m:75-290."""
    x = x0.copy()
    T = [x0.copy()]
    n = 0
    code = 0
    rr = []
    alphas = []
    R = np.sqrt(W)
    wResFun = lambda xx: R * resFun(xx, False)[0]
    D = Js = p = None
    while True:
        s, K, r, J = _weighted(resFun, R, x, True)               # :112-116
        rr.append(np.sqrt(r @ r))
        if trace:
            print('Gauss-Newton-Armijo: iteration %d, residual norm=%.2g' % (n, rr[-1]))
        if n == 0:                                               # :132-143
            if structural_rank(J) < J.shape[1]:
                code = -4
                p = np.full(len(x), np.nan)
                D = Js = None
                break
        p, sing, D, Js, _, _, _, _ = _scaled_gn(J, r)            # :166-174
        if sTest and sing:                                       # :176-184
            code = -2
            break
        Jp = J @ p                                               # :187
        if termFun(Jp, r):                                       # :191
            break
        n += 1
        alpha, xNew, rNew = _linesearch(wResFun, vetoFun, x, p, alphaMin, r, r @ Jp, mu)
        x = xNew
        r = rNew
        alphas.append(alpha)
        T.append(x.copy())
        if alpha == 0:                                           # :217-223
            code = -3
            rr.append(rr[-1])
            break
        if n > maxIter:                                          # :225-231
            code = -1
            rr.append(np.sqrt(r @ r))
            break
    final = NS(unweighted=NS(r=s, J=K), weighted=NS(r=r, J=J), scaled=NS(D=D, J=Js), p=p)
    T = np.array(T).T[:, :n + 1]
    return x, code, n, final, T, np.array(rr), np.array(alphas)


def _linesearch(fun, veto, x, p, alphaMin, r0, fp0, mu):
    """gauss_newton_armijo.m:249-290 Armijo backtracking."""
    f0 = 0.5 * (r0 @ r0)
    alpha = 1.0
    while alpha >= alphaMin:
        t = x + alpha * p
        r = fun(t)
        f = 0.5 * (r @ r)
        redOK = f < f0 + mu * alpha * fp0
        fail = bool(veto(t)) if (redOK and veto) else False
        if redOK and not fail:
            return alpha, t, r
        alpha = alpha / 2
    return 0.0, x, r0


def levenberg_marquardt_powell(resFun, vetoFun, x0, W, maxIter, termFun, doTrace,
                               delta0, mu, eta):
    """levenberg_marquardt_powell.m:60-230."""
    x = x0.copy()
    T = {0: x0.copy()}
    n = 0
    code = 0
    delta = delta0
    deltas, rhos, steps = [], [], []
    R = np.sqrt(W)
    s, K, r, J = _weighted(resFun, R, x, True)
    f = 0.5 * r @ r
    rr = []
    p = None
    while True:
        rr.append(np.sqrt(r @ r))                                # :109
        if n == 0:
            if structural_rank(J) < J.shape[1]:                  # :113-122
                code = -4
                p = np.full(len(x), np.nan)
                break
        p, pGN, step = _dogleg(r, J, delta)                      # :125
        deltas.append(delta)
        steps.append(step)
        JpGN = J @ pGN
        Jp = J @ p
        if step == 0 and termFun(JpGN, r):                       # :134
            break
        t = x + p
        rt = R * resFun(t, False)[0]
        ft = 0.5 * rt @ rt
        veto = bool(vetoFun(t)) if vetoFun else False
        predicted = -r @ Jp - 0.5 * Jp @ Jp                      # :153
        actual = f - ft
        rho = actual / predicted
        rhos.append(rho)
        if doTrace:
            print('Levenberg-Marquardt-Powell: iteration %d, residual norm=%.2g, delta=%.2g, '
                  'step=%s, rho=%.1f' % (n, rr[-1], delta, ['GN', 'IP', 'CP'][step], rho))
        if veto or rho <= mu:                                    # :166-179
            delta = delta / 2
            npGN = np.linalg.norm(pGN)
            if delta > npGN:
                delta = delta / 2.0 ** np.ceil(np.log2(delta / npGN))
        else:                                                    # :180-195
            x = t
            s, K, r, J = _weighted(resFun, R, x, True)
            f = 0.5 * r @ r
            if rho >= eta:
                delta = delta * 2
        T[n] = x.copy()                                          # :197-204 T(:,n+1)=x
        n += 1
        if n > maxIter:
            code = -1
            break
    T[n] = x.copy()                                              # :215-218
    final = NS(unweighted=NS(r=s, J=K), weighted=NS(r=r, J=J), p=p)
    Tm = np.array([T[k] for k in range(n)]).T if n > 0 else np.zeros((len(x), 0))  # :226 T(:,1:n)
    return (x, code, n, final, Tm, np.array(rr), np.array(deltas), np.array(rhos),
            np.array(steps))


def _dogleg(r, J, delta):
    """levenberg_marquardt_powell.m:232-335 (Powell single dogleg)."""
    pGN, _, D, Js, Hs, gs, Jn, Jn2 = _scaled_gn(J, r)
    if np.linalg.norm(pGN) <= delta:                             # :281-286
        return pGN, pGN, 0
    invD2gs = Jn2 * gs                                           # :304-307
    g = Jn * gs
    lambdaStar = (g @ g) / (invD2gs @ (Hs @ invD2gs))            # :309
    CP = -lambdaStar * g
    if np.linalg.norm(CP) > delta:                               # :313-318
        return -g / np.linalg.norm(g) * delta, pGN, 2
    A = np.sum((CP - pGN) ** 2)                                  # :324-332
    B = np.sum(2 * CP * (pGN - CP))
    C = np.sum(CP ** 2) - delta ** 2
    k = (-B + np.sqrt(B ** 2 - 4 * A * C)) / (2 * A)
    return CP + k * (pGN - CP), pGN, 1
