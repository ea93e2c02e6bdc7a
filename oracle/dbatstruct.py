"""DBAT problem struct and (de)serialisation — oracle restatement (test infrastructure).

Follows `code/misc/emptydbatstruct.m:8-182` (field layout),
`code/misc/buildserialindices.m:57-221`, `indvec.m:29`, `serialize.m:14-18`,
`deserialize.m:28-30`, `buildweightmatrix.m:13-43`, `seteoest.m:90-128`.
All indices are 0-based; "linear" indices are column-major (MATLAB order).
"""
from types import SimpleNamespace as NS

import numpy as np
import scipy.sparse as sp


def new_struct(IOval, EOval, OPval, IPval, ip_img, ip_op, pxSize, imSize,
               distModel=3, nK=3, nP=2, IPstd=None, IOblock=None, EOblock=None):
    """Build a DBAT struct (subset of emptydbatstruct.m:8-182 that the hot path reads).

    IOval NC x nImg, EOval 6 x nImg, OPval 3 x nOP, IPval 2 x nIP (pixels),
    ip_img/ip_op: image column and OP column of every image point.
    Image points are re-sorted by (image, OP column) as `prob2dbatstruct.m:349-365`
    / `setdbatpts.m` do, so that IP.ix is monotone inside every image.
    """
    IOval = np.array(IOval, dtype=float, order='F')
    EOval = np.array(EOval, dtype=float, order='F')
    OPval = np.array(OPval, dtype=float, order='F')
    NC, nImg = IOval.shape
    nOP = OPval.shape[1]
    ip_img = np.asarray(ip_img, dtype=np.int64)
    ip_op = np.asarray(ip_op, dtype=np.int64)
    order = np.lexsort((ip_op, ip_img))
    ip_img, ip_op = ip_img[order], ip_op[order]
    IPval = np.array(IPval, dtype=float)[:, order]
    nIP = IPval.shape[1]
    if IPstd is None:
        IPstd = np.ones((2, nIP))
    elif np.ndim(IPstd) == 0:
        IPstd = np.full((2, nIP), float(IPstd))
    else:
        IPstd = np.broadcast_to(np.asarray(IPstd, dtype=float), (2, nIP))[:, order]
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
    s.IP = NS(val=IPval, std=np.array(IPstd, dtype=float), img=ip_img, op=ip_op,
              cam=ip_img.copy(), sigmas=np.array([1.0]))
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


def vis_matrix(s):
    """IP.vis / IP.ix (emptydbatstruct.m:66-72) as CSC nOP x nImg; ix stores IP column + 1."""
    nOP, nImg = s.OP.val.shape[1], s.EO.val.shape[1]
    ix = sp.csc_matrix((np.arange(1, len(s.IP.img) + 1), (s.IP.op, s.IP.img)),
                       shape=(nOP, nImg))
    return ix


def _serializeblock(block, est, useObs):
    """buildserialindices.m:162-221 (serializeblock)."""
    block = np.array(block, dtype=int)
    nr, ncol = block.shape
    block[~est] = 0                                   # :173-174
    leading = np.zeros(block.shape, dtype=int)
    simple = True
    for i in range(nr):                               # :180-187
        seen = set()
        for j in range(ncol):
            b = block[i, j]
            if b == 0:
                continue
            if b in seen:
                simple = False
            else:
                seen.add(b)
                leading[i, j] = 1                    # first occurrence of each block id
    lead_lin = np.flatnonzero(leading.ravel(order='F'))   # find(leading), column-major
    serial = NS(src=lead_lin, dest=np.arange(len(lead_lin)))
    serial.obs = np.flatnonzero(useObs.ravel(order='F')[lead_lin])   # :198
    dist = np.full(block.shape, -1, dtype=int)        # :202-203 (−1 = not in x)
    dist.ravel(order='F')  # no-op, clarity
    dist_f = dist.reshape(-1, order='F').copy()
    dist_f[lead_lin] = serial.dest
    dist = dist_f.reshape(block.shape, order='F')
    if not simple:                                    # :206-214 fan-out over repeated ids
        for k in range(len(serial.dest)):
            i, j = np.argwhere(dist == k)[0]
            in_block = block[i, :] == block[i, j]
            dist[i, in_block] = k
    dist_f = dist.reshape(-1, order='F')
    dest = np.flatnonzero(dist_f >= 0)                # :217-219
    deserial = NS(dest=dest, src=dist_f[dest])
    blockIx = np.flatnonzero(leading.any(axis=0))
    return leading.astype(bool), serial, deserial, blockIx


def buildserialindices(s):
    """buildserialindices.m:57-159; x order = [IO; EO; OP]."""
    IOlead, IOser, IOdes, blockIx = _serializeblock(
        s.IO.struct.block, s.bundle.est.IO, s.prior.IO.use)
    nImg = s.EO.val.shape[1]
    if len(blockIx) == 0:                             # :69-81
        s.EO.cam = np.arange(nImg)
    elif len(blockIx) == 1:
        s.EO.cam = np.full(nImg, blockIx[0])
    else:
        s.EO.cam = np.full(nImg, -1)
    EOlead, EOser, EOdes, _ = _serializeblock(
        s.EO.struct.block, s.bundle.est.EO, s.prior.EO.use)
    nOP = s.OP.val.shape[1]
    _, OPser, OPdes, _ = _serializeblock(
        np.tile(np.arange(1, nOP + 1), (3, 1)), s.bundle.est.OP, s.prior.OP.use)
    n = 0                                             # :109-128
    for ser, des in ((IOser, IOdes), (EOser, EOdes), (OPser, OPdes)):
        ser.dest = ser.dest + n
        des.src = des.src + n
        n += len(ser.dest)
    s.IO.struct.leading = IOlead
    s.EO.struct.leading = EOlead
    s.prior.IO.use = s.prior.IO.use & IOlead          # :135-136
    s.prior.EO.use = s.prior.EO.use & EOlead
    s.bundle.serial = NS(IO=IOser, EO=EOser, OP=OPser, n=n)
    s.bundle.deserial = NS(IO=IOdes, EO=EOdes, OP=OPdes, n=n)
    numObs = [2 * len(s.IP.img), len(IOser.obs), len(EOser.obs), len(OPser.obs)]
    base = 0                                          # indvec.m:29
    ixs = []
    for k in numObs:
        ixs.append(np.arange(base, base + k))
        base += k
    s.post.res.ix = NS(IP=ixs[0], IO=ixs[1], EO=ixs[2], OP=ixs[3], n=base)
    return s


def _lin(a):
    return a.reshape(-1, order='F')


def serialize(s):
    """serialize.m:14-18."""
    x = np.full(s.bundle.serial.n, np.nan)
    x[s.bundle.serial.IO.dest] = _lin(s.IO.val)[s.bundle.serial.IO.src]
    x[s.bundle.serial.EO.dest] = _lin(s.EO.val)[s.bundle.serial.EO.src]
    x[s.bundle.serial.OP.dest] = _lin(s.OP.val)[s.bundle.serial.OP.src]
    return x


def deserialize(s, x):
    """deserialize.m:28-30.  Returns (IO, EO, OP) value arrays updated from x (s untouched)."""
    IO = _lin(s.IO.val).copy()
    EO = _lin(s.EO.val).copy()
    OP = _lin(s.OP.val).copy()
    IO[s.bundle.deserial.IO.dest] = x[s.bundle.deserial.IO.src]
    EO[s.bundle.deserial.EO.dest] = x[s.bundle.deserial.EO.src]
    OP[s.bundle.deserial.OP.dest] = x[s.bundle.deserial.OP.src]
    return (IO.reshape(s.IO.val.shape, order='F'), EO.reshape(s.EO.val.shape, order='F'),
            OP.reshape(s.OP.val.shape, order='F'))


def buildweightmatrix(s):
    """buildweightmatrix.m:13-43.  Returns diag(W) = 1/sigma^2 (length m)."""
    stdIPmm = s.IP.std * s.IO.sensor.pxSize[:, s.IP.cam]
    d = np.full(s.post.res.ix.n, np.nan)
    d[s.post.res.ix.IP] = _lin(stdIPmm) ** 2
    d[s.post.res.ix.IO] = _lin(s.prior.IO.std)[_lin(s.prior.IO.use)] ** 2
    d[s.post.res.ix.EO] = _lin(s.prior.EO.std)[_lin(s.prior.EO.use)] ** 2
    d[s.post.res.ix.OP] = _lin(s.prior.OP.std)[_lin(s.prior.OP.use)] ** 2
    return 1.0 / d


def seteoest_depend(s, camNo=0):
    """seteoest.m:90-128 ('depend' datum): fix base camera and the largest offset coordinate."""
    basePos = s.EO.val[0:3, camNo]
    offset = s.EO.val[0:3, :] - basePos[:, None]
    i, j = np.argwhere(offset == offset.max())[0]
    s.bundle.est.EO[:] = True
    s.bundle.est.EO[:, camNo] = False
    s.bundle.est.EO[i, j] = False
    return s


def buildparamtypes(s):
    """buildparamtypes.m:22-104 (single IO block): type strings of every IO / EO / OP element, e.g.
    'cc', 'K1', 'EX-3', 'om-21', 'OX-12/13' (index/id), 'CX-97/1001' for control points."""
    NC, nImg = s.IO.val.shape
    nK, nP = int(s.IO.model.nK), int(s.IO.model.nP)
    base = ['cc', 'px', 'py', 'as', 'sk'] + ['K%d' % k for k in range(1, nK + 1)] + ['P%d' % k for k in range(1, nP + 1)]
    IOt = np.empty((NC, nImg), dtype=object)
    for i in range(nImg):
        IOt[:, i] = base
    names = ['EX', 'EY', 'EZ', 'om', 'ph', 'ka']
    eo_id = getattr(s.EO, 'id', None)
    EOt = np.empty((6, nImg), dtype=object)
    useIds = eo_id is not None and nImg > 1 and np.any(np.arange(1, nImg + 1) != np.asarray(eo_id))
    for i in range(nImg):
        suf = '' if nImg == 1 else ('-%d(%d)' % (i + 1, eo_id[i]) if useIds else '-%d' % (i + 1))
        EOt[:, i] = [n + suf for n in names]
    nOP = s.OP.val.shape[1]
    op_id = getattr(s.OP, 'id', None)
    isCtrl = getattr(s.prior.OP, 'isCtrl', None)
    isCheck = getattr(s.prior.OP, 'isCheck', None)
    OPt = np.empty((3, nOP), dtype=object)
    for j in range(nOP):
        pre = 'O'
        if isCtrl is not None and isCtrl[j]:
            pre = 'C'
        if isCheck is not None and isCheck[j]:
            pre = 'H'
        suf = ''
        if nOP > 1 and op_id is not None:
            suf = '-%d' % (j + 1)
            if op_id[j] != j + 1:
                suf += '/%d' % op_id[j]
        OPt[:, j] = [pre + c + suf for c in 'XYZ']
    return IOt, EOt, OPt


def paramtypes(s):
    """Second output of serialize.m:20-25: the type string of every element of x."""
    IOt, EOt, OPt = buildparamtypes(s)
    ser = s.bundle.serial
    t = np.empty(ser.n, dtype=object)
    t[ser.IO.dest] = IOt.reshape(-1, order='F')[ser.IO.src]
    t[ser.EO.dest] = EOt.reshape(-1, order='F')[ser.EO.src]
    t[ser.OP.dest] = OPt.reshape(-1, order='F')[ser.OP.src]
    return t


def dmperm_match(J):
    """`p = dmperm(A)` for a tall sparse matrix: p[j] = row matched to column j (1-based) or 0.
    CSparse's cs_maxtrans as MATLAB runs it: columns in order, cheap assignment first, then an
    augmenting depth-first search - so when a column set cannot be matched completely it is the LAST
    columns of the set that stay unmatched (bundle.m:433-442 reports exactly those parameters)."""
    import scipy.sparse as sp
    A = sp.csc_matrix(J)
    A.eliminate_zeros()
    A.sort_indices()
    m, n = A.shape
    Ap, Ai = A.indptr, A.indices
    rowmatch = -np.ones(m, dtype=np.int64)            # column matched to each row
    colmatch = -np.ones(n, dtype=np.int64)
    cheap = Ap[:-1].copy()
    for k in range(n):
        # depth-first search for an augmenting path starting at column k
        stack = [k]
        ptr = {k: Ap[k]}
        parent_row = {}
        found = -1
        visited = set([k])
        while stack:
            j = stack[-1]
            # cheap assignment: first unmatched row in column j
            hit = -1
            while cheap[j] < Ap[j + 1]:
                i = Ai[cheap[j]]
                cheap[j] += 1
                if rowmatch[i] < 0:
                    hit = i
                    break
            if hit >= 0:
                found = hit
                break
            advanced = False
            while ptr[j] < Ap[j + 1]:
                i = Ai[ptr[j]]
                ptr[j] += 1
                j2 = rowmatch[i]
                if j2 >= 0 and j2 not in visited:
                    visited.add(j2)
                    parent_row[j2] = i
                    ptr[j2] = Ap[j2]
                    stack.append(j2)
                    advanced = True
                    break
            if not advanced:
                stack.pop()
        if found >= 0:
            # augment along the stack
            i = found
            for j in reversed(stack):
                prev = colmatch[j]
                colmatch[j] = i
                rowmatch[i] = j
                i = prev if j != k else -1
                if j != k:
                    i = parent_row[j]
    return np.where(colmatch >= 0, colmatch + 1, 0)
