"""One session through mex/dbat_mex.c's mexFunction (tests/mexharness.py) against the same calls through the ctypes
binding of dbat_b200 - both end in the same libdbatgpu.so, so everything except the optimiser's wall time must agree
exactly.  Run by tests/test_mex_gateway.py in its own process (needs a CUDA device)."""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    import dbat_b200
    import mexharness
    from dbat_b200 import build, photogrammetry
    from dbat_b200.dbatstruct import buildserialindices
    from dbat_b200.synth import make_scene

    lib = build.build()
    mex = mexharness.Harness(mexharness.build(tempfile.mkdtemp(prefix='mexh'), lib))
    s, _ = make_scene(21, 100, rays=10, seed=7)
    if s.bundle.serial is None:
        buildserialindices(s)
    x0 = dbat_b200.serialize(s)
    nImg = s.EO.val.shape[1]

    P = dbat_b200.Problem(s)
    h = mex.call('create', mexharness.desc_struct(s))
    assert h.dtype == np.uint64 and h.shape == (1, 1) and mex.H.hs_lock_count() == 1
    h = np.uint64(h[0, 0])

    # [f] = dbat_mex('eval', h, x, weighted)
    for w in (0.0, 1.0):
        f = mex.call('eval', h, x0, w)
        assert f.shape == (P.m, 1) and np.array_equal(f[:, 0], P(x0, weighted=bool(w)))
    # J = dbat_mex('jacobian', h, weighted): MATLAB-owned sparse m-by-n, filled in place
    J = mex.call('jacobian', h, 1.0)
    Jp = P.jacobian(True)
    assert J.shape == Jp.shape and np.array_equal(J.indptr, Jp.indptr) and np.array_equal(J.indices, Jp.indices)
    assert np.array_equal(J.data, Jp.data)
    # wrong-size x is refused by the gateway before the library sees it
    try:
        mex.call('eval', h, x0[:-1], 1.0)
        raise AssertionError('short x accepted')
    except mexharness.MexError as e:
        assert e.id == 'DBAT:dbat_mex:badSize'

    # r = dbat_mex('solve', h, method, opts, x0)
    opts = dict(maxIter=20, convTol=1e-6, absTerm=0, singularTest=1, doTrace=0)
    for method, name in ((1, 'gna'), (2, 'lm'), (3, 'lmp')):
        r = mex.call('solve', h, float(method), opts, x0)
        q = P.solve(name, x0)
        assert int(r['code'][0, 0]) == q.code and int(r['n'][0, 0]) == q.n, (name, r['code'], r['n'], q.code, q.n)
        assert name != 'gna' or q.code == 0
        assert np.array_equal(r['x'][:, 0], q.x) and np.array_equal(r['p'][:, 0], q.p), name
        assert r['T'].shape == q.T.shape and np.array_equal(r['T'], q.T), name
        assert np.array_equal(r['rr'].ravel(), q.rr) and np.array_equal(r['damping'].ravel(), q.damping), name
        assert np.array_equal(r['r_w'][:, 0], q.r_w) and np.array_equal(r['r_u'][:, 0], q.r_u), name
        if name == 'lmp':
            assert np.array_equal(r['rhos'].ravel(), q.rhos) and np.array_equal(r['steps'].ravel(), q.steps)
    s0 = float(np.sqrt(q.r_w @ q.r_w / (P.m - P.n)))

    # C = dbat_mex('cov', h, which, s0, dims)
    Cm = mex.call('cov', h, 2.0, s0, np.array([6.0, 6.0, nImg]))
    Cp = P.cov('ceo', s0)
    assert Cm.shape == (6, 6, nImg) and np.array_equal(Cm.transpose(2, 1, 0), Cp)
    Cm = mex.call('cov', h, 3.0, s0, np.array([3.0, 3.0, s.OP.val.shape[1]]))
    assert np.array_equal(Cm.transpose(2, 1, 0), P.cov('cop', s0))

    # [sd, eo, op, io] = dbat_mex('covstats', h, s0, thres)
    sd, eo, op, io = mex.call('covstats', h, s0, 0.5, nlhs=4)
    st = P.cov_stats(s0, 0.5)
    assert np.array_equal(sd[:, 0], st['std'])
    for got, key in ((eo, 'eo'), (op, 'op'), (io, 'io')):
        r_, c_, b_, v_ = st[key]
        assert got.shape == (len(v_), 4), (key, got.shape, len(v_))
        assert np.array_equal(got[:, 0], b_ + 1) and np.array_equal(got[:, 1], r_ + 1) and np.array_equal(got[:, 2], c_ + 1)
        assert np.array_equal(got[:, 3], v_)
    print('covstats pairs above 0.5: eo %d, op %d, io %d' % (len(st['eo'][3]), len(st['op'][3]), len(st['io'][3])))

    # [OP, res] = dbat_mex('forwintersect', IO, EO, pxSize, IPval, im, op, pts, nK, nP)
    s2, ids, res = photogrammetry.forwintersect(s)
    i64 = lambda a: np.asarray(a, dtype=np.int64)
    OP, rs = mex.call('forwintersect', s.IO.val, s.EO.val[0:6], np.broadcast_to(s.IO.sensor.pxSize, (2, nImg)), s.IP.val,
                      i64(s.IP.img) + 1, i64(s.IP.op) + 1, i64(np.arange(s.OP.val.shape[1])) + 1,
                      float(s.IO.model.nK), float(s.IO.model.nP), nlhs=2)
    assert np.array_equal(OP, s2.OP.val[0:3]) and np.array_equal(rs.ravel(), res)

    # [EO, res] = dbat_mex('resect3', X3, x3, int64(testStart), XT, xT, behind): the batch the host mirror of
    # resect.m:52-95 prepares (captured from dbat_b200.photogrammetry.resect), through the gateway and through ctypes
    import ctypes as C
    from dbat_b200 import _lib
    _lib.lib()                                  # argtypes are bound to the real ResectDesc before it is wrapped
    cap = {}
    orig = _lib.ResectDesc

    def capture(nCam, X3, x3, ts, XT, xT, behind):
        t = np.ctypeslib.as_array(ts, (nCam + 1,)).copy()
        cap.update(nCam=nCam, ts=t, behind=behind, X3=np.ctypeslib.as_array(X3, (9 * nCam,)).copy(),
                   x3=np.ctypeslib.as_array(x3, (6 * nCam,)).copy(), XT=np.ctypeslib.as_array(XT, (3 * int(t[-1]),)).copy(),
                   xT=np.ctypeslib.as_array(xT, (2 * int(t[-1]),)).copy())
        return orig(nCam, X3, x3, ts, XT, xT, behind)

    _lib.ResectDesc = capture
    try:
        photogrammetry.resect(s, 'all', cpId=np.asarray(s.OP.id) if getattr(s.OP, 'id', None) is not None
                              else np.arange(s.OP.val.shape[1]), n=2)
    finally:
        _lib.ResectDesc = orig
    nC = cap['nCam']
    assert nC >= nImg, nC
    d = orig(nC, _lib.dptr(cap['X3']), _lib.dptr(cap['x3']), _lib.iptr(cap['ts']), _lib.dptr(cap['XT']), _lib.dptr(cap['xT']), 1)
    EOc, resc = np.empty((nC, 6)), np.empty(nC)
    assert _lib.lib().dbat_resect3(C.byref(d), _lib.dptr(EOc), _lib.dptr(resc)) == 0
    EOm, resm = mex.call('resect3', cap['X3'], cap['x3'], cap['ts'], cap['XT'], cap['xT'], True, nlhs=2)
    assert EOm.shape == (6, nC) and np.array_equal(EOm.T, EOc, equal_nan=True) and np.array_equal(resm.ravel(), resc, equal_nan=True)
    assert np.isfinite(resc).any()
    print('resect3: %d candidates, %d with a solution' % (nC, int(np.isfinite(resc).sum())))

    mex.call('destroy', h, nlhs=0)
    assert mex.H.hs_lock_count() == 0
    P.close()
    print('MEX SESSION OK')


if __name__ == '__main__':
    main()
