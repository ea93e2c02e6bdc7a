"""Drives mex/dbat_mex.c without MATLAB: the gateway is compiled against tests/mexstub/mex.h (a stand-in that
declares the documented MEX / C Matrix API subset the gateway uses), linked to libdbatgpu.so, and called through
ctypes with NumPy values converted to mxArrays the way MATLAB would pass them (column-major doubles, int64 index
vectors, 1x1 structs, char rows, a uint64 handle).  Test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MX_DOUBLE, MX_STRUCT, MX_INT32, MX_INT64, MX_UINT64 = 6, 2, 12, 14, 15


class MexError(Exception):
    def __init__(self, ident, msg):
        super().__init__('%s: %s' % (ident, msg))
        self.id, self.msg = ident, msg


def build(outdir, lib_path=None):
    """gcc -Wall -Wextra -Werror on the gateway + the stand-in; returns the path of the harness library."""
    lib_path = lib_path or os.path.join(ROOT, 'dbat_b200', 'libdbatgpu.so')
    out = os.path.join(str(outdir), 'libmexharness.so')
    cmd = ['gcc', '-std=c11', '-D_GNU_SOURCE', '-O1', '-Wall', '-Wextra', '-Werror', '-shared', '-fPIC',
           '-I', os.path.join(ROOT, 'tests', 'mexstub'), '-I', os.path.join(ROOT, 'include'),
           os.path.join(ROOT, 'mex', 'dbat_mex.c'), os.path.join(ROOT, 'tests', 'mexstub', 'mexstub.c'),
           lib_path, '-Wl,-rpath,' + os.path.dirname(lib_path), '-o', out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode:
        raise RuntimeError('gateway does not compile:\n' + r.stderr)
    return out


class Harness:
    def __init__(self, path):
        H = self.H = C.CDLL(path)
        vp = C.c_void_p
        for name, res, args in [
                ('hs_double', vp, [C.c_size_t, C.c_size_t, vp]), ('hs_int64', vp, [C.c_size_t, C.c_size_t, vp]),
                ('hs_uint64', vp, [C.c_uint64]), ('hs_logical', vp, [C.c_int]), ('hs_string', vp, [C.c_char_p]),
                ('hs_struct', vp, [C.c_int, C.POINTER(C.c_char_p)]), ('hs_class', C.c_int, [vp]),
                ('hs_dim', C.c_size_t, [vp, C.c_int]), ('hs_lock_count', C.c_int, []), ('hs_has_at_exit', C.c_int, []),
                ('hs_err_id', C.c_char_p, []), ('hs_err_msg', C.c_char_p, []),
                ('hs_call', C.c_int, [C.c_int, C.POINTER(vp), C.c_int, C.POINTER(vp)]),
                ('hs_nfields', C.c_int, [vp]), ('hs_field_name', C.c_char_p, [vp, C.c_int]), ('hs_field', vp, [vp, C.c_int]),
                ('mxSetField', None, [vp, C.c_size_t, C.c_char_p, vp]), ('mxGetData', vp, [vp]), ('mxIsSparse', C.c_bool, [vp]),
                ('mxGetIr', vp, [vp]), ('mxGetJc', vp, [vp]), ('mxDestroyArray', None, [vp])]:
            f = getattr(H, name)
            f.restype, f.argtypes = res, args

    # ---- NumPy -> mxArray
    def to_mx(self, v):
        H = self.H
        if isinstance(v, str):
            return H.hs_string(v.encode())
        if isinstance(v, dict):
            names = (C.c_char_p * len(v))(*[k.encode() for k in v])
            s = H.hs_struct(len(v), names)
            for k, x in v.items():
                H.mxSetField(s, 0, k.encode(), self.to_mx(x))
            return s
        if isinstance(v, (bool, np.bool_)):
            return H.hs_logical(int(v))
        if isinstance(v, np.uint64):
            return H.hs_uint64(int(v))
        if isinstance(v, (int, float, np.floating, np.integer)):
            v = np.array([[float(v)]])
        a = np.asarray(v)
        if a.ndim == 1:
            a = a[:, None]                                 # MATLAB column vector
        m, n = a.shape[0], int(np.prod(a.shape[1:]))
        flat = np.ascontiguousarray(a.reshape(m, n, order='F').T)          # column-major payload
        if a.dtype == np.int64:
            return H.hs_int64(m, n, flat.ctypes.data)
        flat = np.ascontiguousarray(flat, dtype=np.float64)
        return H.hs_double(m, n, flat.ctypes.data)

    # ---- mxArray -> NumPy
    def from_mx(self, p):
        H = self.H
        cls = H.hs_class(p)
        dims = [H.hs_dim(p, k) for k in range(3)]
        if cls == MX_STRUCT:
            return {H.hs_field_name(p, k).decode(): self.from_mx(H.hs_field(p, k)) for k in range(H.hs_nfields(p))}
        if H.mxIsSparse(p):
            m, n = dims[0], dims[1]
            jc = np.ctypeslib.as_array(C.cast(H.mxGetJc(p), C.POINTER(C.c_uint64)), (n + 1,)).astype(np.int64)
            nnz = int(jc[-1])
            ir = np.ctypeslib.as_array(C.cast(H.mxGetIr(p), C.POINTER(C.c_uint64)), (max(nnz, 1),))[:nnz].astype(np.int64)
            va = np.ctypeslib.as_array(C.cast(H.mxGetData(p), C.POINTER(C.c_double)), (max(nnz, 1),))[:nnz].copy()
            return sp.csc_matrix((va, ir, jc), shape=(m, n))
        ct = {MX_DOUBLE: C.c_double, MX_INT32: C.c_int32, MX_INT64: C.c_int64, MX_UINT64: C.c_uint64}[cls]
        numel = dims[0] * dims[1] * dims[2]
        if numel == 0:
            return np.zeros(dims if dims[2] != 1 else dims[:2], dtype=ct)
        a = np.ctypeslib.as_array(C.cast(H.mxGetData(p), C.POINTER(ct)), (numel,)).copy()
        a = a.reshape(dims if dims[2] != 1 else dims[:2], order='F')
        return a

    def call(self, *args, nlhs=1):
        """dbat_mex(args...) with nlhs outputs; raises MexError where MATLAB would raise the error."""
        H = self.H
        prhs = (C.c_void_p * max(1, len(args)))(*[self.to_mx(a) for a in args])
        plhs = (C.c_void_p * max(1, nlhs))()
        rc = H.hs_call(nlhs, plhs, len(args), prhs)
        try:
            if rc:
                raise MexError(H.hs_err_id().decode(), H.hs_err_msg().decode())
            outs = [self.from_mx(plhs[k]) if plhs[k] else None for k in range(nlhs)]
        finally:
            for k in range(len(args)):
                H.mxDestroyArray(prhs[k])
            for k in range(nlhs):
                if plhs[k]:
                    H.mxDestroyArray(plhs[k])
        return outs[0] if nlhs == 1 else outs


def desc_struct(s):
    """The struct argument of dbat_mex('create', d): the fields of dbat_problem_desc (include/dbat_gpu.h) taken from
    a DBAT struct `s` the way the shimmed bundle.m of INTEGRATION.md does it - 1-based int64 indices, column-major
    doubles."""
    lin = lambda a: np.asarray(a, dtype=np.float64).ravel(order='F')
    i64 = lambda a: np.asarray(a, dtype=np.int64)
    ser, des = s.bundle.serial, s.bundle.deserial
    d = dict(nImg=s.EO.val.shape[1], nOP=s.OP.val.shape[1], nIP=len(s.IP.img),
             distModel=int(np.unique(s.IO.model.distModel)[0]), nK=int(s.IO.model.nK), nP=int(s.IO.model.nP), n=int(ser.n),
             IOval=lin(s.IO.val), EOval=lin(s.EO.val[0:6]), OPval=lin(s.OP.val), IPval=lin(s.IP.val), IPstd=lin(s.IP.std),
             pxSize=lin(np.broadcast_to(s.IO.sensor.pxSize, (2, s.EO.val.shape[1]))), IPimg=i64(s.IP.img) + 1, IPop=i64(s.IP.op) + 1)
    for nm in ('IO', 'EO', 'OP'):
        dd = getattr(des, nm)
        d[nm + 'des_src'], d[nm + 'des_dest'] = i64(dd.src) + 1, i64(dd.dest) + 1
    px, pv, ps = [], [], []
    for nm in ('IO', 'EO', 'OP'):                          # prior_obs.m:28-65, buildweightmatrix.m:26-31
        sr, pr = getattr(ser, nm), getattr(s.prior, nm)
        px.append(i64(sr.dest[sr.obs]) + 1)
        pv.append(lin(pr.val)[sr.src[sr.obs]])
        ps.append(lin(pr.std)[np.asarray(pr.use).ravel(order='F')])
        d['nPrior' + nm] = len(px[-1])
    d['prior_x'], d['prior_val'], d['prior_std'] = np.concatenate(px), np.concatenate(pv), np.concatenate(ps)
    return d
