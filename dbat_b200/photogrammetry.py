"""Host mirror of the reference's start-value step `forwintersect` (code/photogrammetry/forwintersect.m),
running on the device through `dbat_forwintersect` (csrc/startval.cu).  No CPU fallback."""
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
