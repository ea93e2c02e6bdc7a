"""Host-side DBAT problem struct and (de)serialisation indices for the B200 path.

Mirrors the reference's struct layout (`code/misc/emptydbatstruct.m:8-182`) and index
contract (`code/misc/buildserialindices.m:57-221`, `serialize.m:14-18`,
`deserialize.m:28-30`, `buildweightmatrix.m:13-43`) so that the same struct goes in and
out of `bundle()`.  Vectorised NumPy; indices are 0-based here and converted to MATLAB's
1-based int64 only at the C ABI (`dbat_b200/_lib.py`).
"""
from types import SimpleNamespace as NS

import numpy as np


def new_struct(IOval, EOval, OPval, IPval, ip_img, ip_op, pxSize, imSize,
               distModel=3, nK=3, nP=2, IPstd=1.0, IOblock=None, EOblock=None):
    """Assemble a DBAT struct from flat arrays; image points are sorted by (image, OP)
    as `prob2dbatstruct.m:349-365` does."""
    IOval = np.array(IOval, dtype=float, order='F')
    EOval = np.array(EOval, dtype=float, order='F')
    OPval = np.array(OPval, dtype=float, order='F')
    NC, nImg = IOval.shape
    nOP = OPval.shape[1]
    ip_img = np.asarray(ip_img, dtype=np.int64)
    ip_op = np.asarray(ip_op, dtype=np.int64)
    order = np.lexsort((ip_op, ip_img))
    nIP = len(order)
    if np.ndim(IPstd) == 0:
        IPstd = np.full((2, nIP), float(IPstd))
    else:
        IPstd = np.array(np.broadcast_to(np.asarray(IPstd, dtype=float), (2, nIP))[:, order])
    s = NS()
    s.IO = NS(val=IOval,
              model=NS(distModel=np.full(nImg, distModel, dtype=int), nK=nK, nP=nP),
              sensor=NS(pxSize=np.array(np.broadcast_to(pxSize, (2, nImg)), dtype=float),
                        imSize=np.array(np.broadcast_to(imSize, (2, nImg)), dtype=float)),
              struct=NS(block=np.ones((NC, nImg), dtype=int) if IOblock is None
                        else np.array(IOblock, dtype=int), leading=None))
    s.EO = NS(val=EOval, cam=np.zeros(nImg, dtype=int),
              struct=NS(block=np.tile(np.arange(1, nImg + 1), (6, 1)) if EOblock is None
                        else np.array(EOblock, dtype=int), leading=None))
    s.OP = NS(val=OPval, id=np.arange(nOP))
    s.IP = NS(val=np.array(IPval, dtype=float)[:, order], std=IPstd, img=ip_img[order],
              op=ip_op[order], cam=ip_img[order].copy(), sigmas=np.array([1.0]))
    s.prior = NS(
        IO=NS(use=np.zeros((NC, nImg), bool), val=np.full((NC, nImg), np.nan),
              std=np.full((NC, nImg), np.nan)),
        EO=NS(use=np.zeros((6, nImg), bool), val=np.full((6, nImg), np.nan),
              std=np.full((6, nImg), np.nan)),
        OP=NS(use=np.zeros((3, nOP), bool), val=np.full((3, nOP), np.nan),
              std=np.full((3, nOP), np.nan)))
    s.bundle = NS(est=NS(IO=np.zeros((NC, nImg), bool), EO=np.ones((6, nImg), bool),
                         OP=np.ones((3, nOP), bool)), serial=None, deserial=None)
    s.post = NS(res=NS(ix=None), cov=NS(CEO=None, COP=None), std=NS())
    return s


def _serializeblock(block, est, useObs, distinct=False):
    """buildserialindices.m:162-221: leading element of every estimated block per row,
    serial (matrix -> x) and deserial (x -> matrix, with fan-out over repeated ids)."""
    nr, ncol = block.shape
    if distinct:                                     # every column its own block (OP)
        leading = est.copy()
        dist = np.full(nr * ncol, -1, dtype=np.int64)
        lead_lin = np.flatnonzero(leading.ravel(order='F'))
        dist[lead_lin] = np.arange(len(lead_lin))
    else:
        blk = np.where(est, block, 0)
        leading = np.zeros(blk.shape, bool)
        rep = np.full(blk.shape, -1, dtype=np.int64)  # column of the leading element of (row, col)
        for i in range(nr):
            row = blk[i]
            nz = np.flatnonzero(row)
            if len(nz) == 0:
                continue
            _, first, inv = np.unique(row[nz], return_index=True, return_inverse=True)
            leading[i, nz[first]] = True
            rep[i, nz] = nz[first][inv]
        lead_lin = np.flatnonzero(leading.ravel(order='F'))
        pos = np.full(nr * ncol, -1, dtype=np.int64)
        pos[lead_lin] = np.arange(len(lead_lin))
        rows = np.broadcast_to(np.arange(nr)[:, None], blk.shape)
        lin_rep = rows + nr * rep                      # linear index of the leading element
        dist2 = np.where(rep >= 0, pos[np.where(rep >= 0, lin_rep, 0)], -1)
        dist = dist2.ravel(order='F')
    serial = NS(src=lead_lin, dest=np.arange(len(lead_lin)))
    serial.obs = np.flatnonzero(useObs.ravel(order='F')[lead_lin])
    dest = np.flatnonzero(dist >= 0)
    deserial = NS(dest=dest, src=dist[dest])
    return leading, serial, deserial, np.flatnonzero(leading.any(axis=0))


def buildserialindices(s):
    """buildserialindices.m:57-159 with the default order [IO; EO; OP]."""
    IOlead, IOser, IOdes, blockIx = _serializeblock(
        np.asarray(s.IO.struct.block), s.bundle.est.IO, s.prior.IO.use)
    nImg = s.EO.val.shape[1]
    if len(blockIx) == 0:
        s.EO.cam = np.arange(nImg)
    elif len(blockIx) == 1:
        s.EO.cam = np.full(nImg, blockIx[0])
    else:
        s.EO.cam = np.full(nImg, -1)
    EOlead, EOser, EOdes, _ = _serializeblock(
        np.asarray(s.EO.struct.block), s.bundle.est.EO, s.prior.EO.use)
    _, OPser, OPdes, _ = _serializeblock(np.zeros((3, s.OP.val.shape[1]), int),
                                         s.bundle.est.OP, s.prior.OP.use, distinct=True)
    n = 0
    for ser, des in ((IOser, IOdes), (EOser, EOdes), (OPser, OPdes)):
        ser.dest = ser.dest + n
        des.src = des.src + n
        n += len(ser.dest)
    s.IO.struct.leading = IOlead
    s.EO.struct.leading = EOlead
    s.prior.IO.use = s.prior.IO.use & IOlead
    s.prior.EO.use = s.prior.EO.use & EOlead
    s.bundle.serial = NS(IO=IOser, EO=EOser, OP=OPser, n=n)
    s.bundle.deserial = NS(IO=IOdes, EO=EOdes, OP=OPdes, n=n)
    base = 0
    ixs = []
    for k in (2 * len(s.IP.img), len(IOser.obs), len(EOser.obs), len(OPser.obs)):
        ixs.append(np.arange(base, base + k))
        base += k
    s.post.res.ix = NS(IP=ixs[0], IO=ixs[1], EO=ixs[2], OP=ixs[3], n=base)
    return s


def _lin(a):
    return np.asarray(a).reshape(-1, order='F')


def serialize(s):
    """serialize.m:14-18."""
    x = np.full(s.bundle.serial.n, np.nan)
    x[s.bundle.serial.IO.dest] = _lin(s.IO.val)[s.bundle.serial.IO.src]
    x[s.bundle.serial.EO.dest] = _lin(s.EO.val)[s.bundle.serial.EO.src]
    x[s.bundle.serial.OP.dest] = _lin(s.OP.val)[s.bundle.serial.OP.src]
    return x


def deserialize(s, x, i=None, what=None):
    """deserialize.m:20-100.
    deserialize(s, x): the struct updated in place from the vector of unknowns x (:28-30).
    deserialize(s, E, i): the same from column i of the iteration trace E.trace (0 = start values,
    np.inf = last), i.e. a rewind to that iteration (:31-46).
    deserialize(s, E, v, 'IO'|'EO'|'OP'): an array with one layer per iteration in v ('all' = every
    iteration) of that parameter group; s is left alone (:47-100)."""
    if i is None:
        cols = {None: np.asarray(x)}
    else:
        E = x
        if E is None or getattr(E, 'trace', None) is None:
            raise ValueError('Empty bundle iteration struct')
        T = np.asarray(E.trace)
        if what is None:
            cols = {None: T[:, T.shape[1] - 1 if i == np.inf else int(i)]}
        else:
            v = np.arange(T.shape[1]) if isinstance(i, str) and i == 'all' else np.atleast_1d(i).astype(int)
            fld = getattr(s, what)
            des = getattr(s.bundle.deserial, what)
            out = np.repeat(fld.val[:, :, None], len(v), axis=2)
            r, c = np.unravel_index(des.dest, fld.val.shape, order='F')
            for k, it in enumerate(v):
                out[r, c, k] = T[des.src, it]
            return out
    xv = cols[None]
    for name in ('IO', 'EO', 'OP'):
        fld = getattr(s, name)
        des = getattr(s.bundle.deserial, name)
        v = _lin(fld.val).copy()
        v[des.dest] = xv[des.src]
        fld.val = v.reshape(fld.val.shape, order='F')
    return s


def buildweightmatrix(s):
    """buildweightmatrix.m:13-43 → diag(W) (length m)."""
    stdIPmm = s.IP.std * s.IO.sensor.pxSize[:, s.IP.cam]
    d = np.full(s.post.res.ix.n, np.nan)
    d[s.post.res.ix.IP] = _lin(stdIPmm) ** 2
    d[s.post.res.ix.IO] = _lin(s.prior.IO.std)[_lin(s.prior.IO.use)] ** 2
    d[s.post.res.ix.EO] = _lin(s.prior.EO.std)[_lin(s.prior.EO.use)] ** 2
    d[s.post.res.ix.OP] = _lin(s.prior.OP.std)[_lin(s.prior.OP.use)] ** 2
    return 1.0 / d


def seteoest_depend(s, camNo=0):
    """seteoest.m:90-128: 'depend' datum."""
    offset = s.EO.val[0:3, :] - s.EO.val[0:3, camNo][:, None]
    i, j = np.argwhere(offset == offset.max())[0]
    s.bundle.est.EO[:] = True
    s.bundle.est.EO[:, camNo] = False
    s.bundle.est.EO[i, j] = False
    return s


def buildparamtypes(s):
    """buildparamtypes.m:22-104 for one IO block: the type string of every IO, EO and OP element
    ('cc', 'K1', 'EX-3', 'om-21(7)', 'OX-12/13', control points 'CX-..', check points 'HX-..')."""
    NC, nImg = s.IO.val.shape
    nK, nP = int(s.IO.model.nK), int(s.IO.model.nP)
    io = ['cc', 'px', 'py', 'as', 'sk'] + ['K%d' % (k + 1) for k in range(nK)] + ['P%d' % (k + 1) for k in range(nP)]
    IOt = np.array([io] * nImg, dtype=object).T
    eo_id = getattr(s.EO, 'id', None)
    seq = np.arange(1, nImg + 1)
    with_id = eo_id is not None and nImg > 1 and bool(np.any(np.asarray(eo_id) != seq))
    EOt = np.empty((6, nImg), dtype=object)
    for i in range(nImg):
        tail = '' if nImg == 1 else '-%d' % (i + 1) + ('(%d)' % eo_id[i] if with_id else '')
        for k, nm in enumerate(('EX', 'EY', 'EZ', 'om', 'ph', 'ka')):
            EOt[k, i] = nm + tail
    nOP = s.OP.val.shape[1]
    op_id = getattr(s.OP, 'id', None)
    ctrl = getattr(s.prior.OP, 'isCtrl', None)
    check = getattr(s.prior.OP, 'isCheck', None)
    OPt = np.empty((3, nOP), dtype=object)
    for j in range(nOP):
        lead = 'H' if (check is not None and check[j]) else ('C' if (ctrl is not None and ctrl[j]) else 'O')
        tail = ''
        if nOP > 1 and op_id is not None:
            tail = '-%d' % (j + 1) + ('/%d' % op_id[j] if op_id[j] != j + 1 else '')
        for k, ax in enumerate('XYZ'):
            OPt[k, j] = lead + ax + tail
    return IOt, EOt, OPt


def paramtypes(s):
    """[x,t]=serialize(s) (serialize.m:20-25): the type string of every unknown, in x order."""
    IOt, EOt, OPt = buildparamtypes(s)
    ser = s.bundle.serial
    t = np.empty(ser.n, dtype=object)
    for blk, T in ((ser.IO, IOt), (ser.EO, EOt), (ser.OP, OPt)):
        t[blk.dest] = T.reshape(-1, order='F')[blk.src]
    return t


def column_matching(J):
    """p = dmperm(J) for a tall sparse J: p[j] = 1-based row matched to column j, 0 if unmatched.
    Columns are taken in order (cheap assignment, then augmenting depth-first search), as MATLAB's
    dmperm / CSparse cs_maxtrans does, so an over-determined column set loses its last columns."""
    import scipy.sparse as sps
    A = sps.csc_matrix(J)
    A.eliminate_zeros()
    A.sort_indices()
    m, n = A.shape
    Ap, Ai = A.indptr, A.indices
    row_of = np.full(n, -1, dtype=np.int64)
    col_of = np.full(m, -1, dtype=np.int64)
    nxt = Ap[:-1].copy()                                   # cheap-assignment cursor of every column
    for k in range(n):
        path = [k]                                         # columns on the current DFS path
        via = {}                                           # column -> row through which it was reached
        cur = {k: Ap[k]}
        seen = {k}
        free_row = -1
        while path:
            j = path[-1]
            while nxt[j] < Ap[j + 1] and free_row < 0:
                i = Ai[nxt[j]]
                nxt[j] += 1
                if col_of[i] < 0:
                    free_row = i
            if free_row >= 0:
                break
            moved = False
            while cur[j] < Ap[j + 1]:
                i = Ai[cur[j]]
                cur[j] += 1
                j2 = col_of[i]
                if j2 >= 0 and j2 not in seen:
                    seen.add(j2)
                    via[j2] = i
                    cur[j2] = Ap[j2]
                    path.append(j2)
                    moved = True
                    break
            if not moved:
                path.pop()
        if free_row >= 0:                                  # flip the matching along the path
            i = free_row
            for j in reversed(path):
                row_of[j] = i
                col_of[i] = j
                i = via.get(j, -1)
    return np.where(row_of >= 0, row_of + 1, 0)
