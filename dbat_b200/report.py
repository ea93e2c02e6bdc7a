"""Report-side consumers of the posterior covariance.

Host mirror of the reference's `bundle_result_file` and the statistics it is built from:

  corrmat                  `code/misc/corrmat.m:21-47`
  cumchi2                  `code/misc/cumchi2.m` (chi-square CDF; closed form, not `quad`)
  high_io_correlations     `code/bundle/private/high_io_correlations.m` (cross form, as the report calls it)
  high_eo_correlations     `code/bundle/private/high_eo_correlations.m` (block form)
  high_op_correlations     `code/bundle/private/high_op_correlations.m` (block form)
  test_distortion_params   `code/bundle/private/test_distortion_params.m`
  bundle_residuals         `code/bundle/bundle_residuals.m`
  coverage                 `code/photogrammetry/coverage.m`
  angles                   `code/photogrammetry/angles.m`
  bundle_result_file       `code/bundle/bundle_result_file.m`
  writestats, camangles    `code/file/writestats.m`, `code/photogrammetry/camangles.m` (pre-bundle statistics)

The covariances come from `bundle_cov`, i.e. from the factorisation that lives on the device; what is
done here is O(unknowns) post-processing and text layout.  Everything is written against the block
layout the device path supports (one shared IO block, one EO block per image).  The text is the
reference's, line for line, so that a result file can be diffed against one the reference wrote.
"""
import math
import os
import platform
import re
import time

import numpy as np
import scipy.sparse as sp
from scipy.spatial import ConvexHull, QhullError
from scipy.special import gammainc

from .bundle import bundle_cov as _device_cov

CORR_THRESHOLD = 0.95           # bundle_result_file.m:19-22
SIG_THRESHOLD = 0.95


# ----------------------------------------------------------------------------- statistics
def _dense(C):
    return C.toarray() if sp.issparse(C) else np.asarray(C)


def _diag_blocks(C, k):
    """(N,k,k) diagonal blocks of an (N*k)x(N*k) matrix, dense or sparse block-diagonal."""
    N = C.shape[0] // k
    if sp.issparse(C):
        M = C.tocoo()
        out = np.zeros((N, k, k))
        b = M.row // k
        on = b == M.col // k
        np.add.at(out, (b[on], M.row[on] % k, M.col[on] % k), M.data[on])
        return out
    C = np.asarray(C)
    ix = np.arange(N)[:, None, None] * k
    return C[ix + np.arange(k)[None, :, None], ix + np.arange(k)[None, None, :]]


def corrmat(C, nodiag=False):
    """corrmat.m:21-47: R(i,j) = C(i,j)/sqrt(C(i,i) C(j,j)) clipped to [-1,1], NaN kept, diagonal set to
    1 (or 0 with nodiag).  Returns (R, std).  Works on one matrix or a stack of blocks (..., k, k)."""
    C = _dense(C)
    sd = np.sqrt(np.diagonal(C, axis1=-2, axis2=-1))
    with np.errstate(divide='ignore', invalid='ignore'):
        R = C / sd[..., :, None] / sd[..., None, :]
    bad = np.isnan(R)
    R = np.clip(R, -1, 1)
    k = C.shape[-1]
    R[..., np.arange(k), np.arange(k)] = 0.0 if nodiag else 1.0
    R[bad] = np.nan
    return R, sd


def cumchi2(x, n):
    """cumchi2.m: P(chi2_n <= x); 1 for x beyond n^2 and 100, 0 for x <= 0.  The reference integrates
    the density with `quad` (1e-6); the regularised incomplete gamma function is the same number."""
    x = float(x)
    if math.isnan(x):
        return float('nan')
    if x > n * n and x > 100:
        return 1.0
    return float(gammainc(n / 2.0, x / 2.0)) if x > 0 else 0.0


def _find_lower(R, thres):
    """MATLAB `find(abs(tril(R))>thres)`: (row, col) pairs in column-major order."""
    M = np.tril(np.abs(R), -1) > thres
    c, r = np.nonzero(M.T)
    return r, c


def high_io_correlations(s, e, thres, cov=_device_cov):
    """high_io_correlations.m (cross=true): pairs of leading IO parameters whose posterior correlation
    exceeds thres.  Returns (i, j, v, CIOF): i, j are (row, camera) pairs, 0-based; v the correlations."""
    CIO = _dense(cov(s, e, 'CIOF'))
    R, _ = corrmat(CIO, True)
    lead = np.asarray(s.IO.struct.leading).ravel(order='F')
    R[~lead, :] = 0
    R[:, ~lead] = 0
    r, c = _find_lower(R, thres)
    NC = s.IO.val.shape[0]
    return (np.stack([r % NC, r // NC], 1).reshape(-1, 2), np.stack([c % NC, c // NC], 1).reshape(-1, 2),
            R[r, c], CIO)


def _block_pairs(C, k, thres):
    B = _diag_blocks(C, k)
    R, _ = corrmat(B, True)
    M = np.tril(np.abs(R), -1) > thres                    # (N,k,k)
    # column-major order of the big matrix: by block, then column, then row
    n, c, r = np.nonzero(M.transpose(0, 2, 1))
    return r, c, n, R[n, r, c], B


def high_eo_correlations(s, e, thres, cov=_device_cov):
    """high_eo_correlations.m (block form): (i, j, k, v, CEO) — EO rows i > j of image k (0-based)."""
    CEO = cov(s, e, 'CEO')
    r, c, n, v, _ = _block_pairs(CEO, s.EO.val.shape[0], thres)
    blk = np.asarray(s.EO.struct.block)
    _, first = np.unique(blk.T, axis=0, return_index=True)
    keep = np.isin(n, first)
    return r[keep], c[keep], n[keep], v[keep], CEO


def high_op_correlations(s, e, thres, COP=None, cov=_device_cov):
    """high_op_correlations.m (block form): (i, j, k, v) — coordinates i > j of object point k."""
    if COP is None:
        COP = cov(s, e, 'COP')
    r, c, n, v, _ = _block_pairs(COP, 3, thres)
    return r, c, n, v


def test_distortion_params(s, e, CIO=None, cov=_device_cov):
    """test_distortion_params.m: chi-square p-values of H0 'coefficient is zero' for each radial
    coefficient (K), the first i radial coefficients jointly (KC), the tangential pair (P) and the affine
    pair (B = aspect, skew).  P is written as the reference writes it (`P(j,:)=...`, :59): row j of a
    2 x nCams array, so for camera 1 the P1 row holds the joint value and the P2 row stays NaN."""
    x = s.IO.val
    NC, nCams = x.shape
    nK = s.IO.model.nK
    if CIO is None:
        CIO = cov(s, e, 'CIO')
    B_ = CIO if isinstance(CIO, np.ndarray) and CIO.ndim == 3 else _diag_blocks(CIO, NC)
    est = s.bundle.est.IO
    K = np.full((nK, nCams), np.nan)
    KC = np.full((nK, nCams), np.nan)
    P = np.full((2, nCams), np.nan)
    B = np.full((2, nCams), np.nan)
    fz = getattr(e.final, 'factorized', None)
    if fz is not None and getattr(fz, 'fail', False):
        return K, P, B, KC
    for j in np.flatnonzero(_io_uniq(s)):
        C = B_[j]
        for i in range(nK):
            if est[5 + i, j]:
                K[i, j] = cumchi2(x[5 + i, j] ** 2 / C[5 + i, 5 + i], 1)
            ii = np.arange(5, 6 + i)
            if est[ii, j].all():
                KC[i, j] = cumchi2(x[ii, j] @ np.linalg.solve(C[np.ix_(ii, ii)], x[ii, j]), i + 1)
        ii = 5 + nK + np.arange(2)
        if est[ii, j].all() and j < 2:
            P[j, :] = cumchi2(x[ii, j] @ np.linalg.solve(C[np.ix_(ii, ii)], x[ii, j]), 2)
        for i in range(2):
            if est[3 + i, j]:
                B[i, j] = cumchi2(x[3 + i, j] ** 2 / C[3 + i, 3 + i], 1)
    return K, P, B, KC


test_distortion_params.__test__ = False      # not a pytest test


def bundle_residuals(s, e):
    """bundle_residuals.m: (rms, ptRes) — the length of every image-point residual in pixels (IP
    order) and their RMS.  The reference also scatters ptRes into an nOP x nImg matrix; callers here
    index by (s.IP.op, s.IP.img) instead."""
    n = s.IP.val.shape[1]
    r = np.asarray(e.final.unweighted.r[:2 * n]).reshape(2, -1, order='F')
    pt = np.sqrt(((r / s.IO.sensor.pxSize[:, s.IP.cam]) ** 2).sum(axis=0))
    return float(np.sqrt(np.mean(pt ** 2))), pt


def _hull_area(pts):
    if pts.shape[1] < 3:
        return 0.0
    try:
        return float(ConvexHull(pts.T).volume)
    except QhullError:                                   # collinear points
        return 0.0


def coverage(s, ix=None, union=False):
    """coverage.m: fraction of the image area covered by the measured points of images ix — convex hull
    (c), bounding rectangle (cr) and radial reach from the principal point relative to the farthest
    corner (crr); per image, or of all points together with union=True."""
    nImg = s.EO.val.shape[1]
    ix = np.arange(nImg) if ix is None else np.atleast_1d(ix)

    def radial(pts, i):
        px = s.IO.sensor.pxSize[:, i]
        xy = np.stack([pts[0] * px[0] - s.IO.val[1, i], -pts[1] * px[1] - s.IO.val[2, i]])
        return np.sqrt((xy ** 2).sum(axis=0))

    def corners(i):
        w, h = s.IO.sensor.imSize[:, i]
        return np.array([[0.5, 0.5, w + 0.5, w + 0.5], [0.5, h + 0.5, h + 0.5, 0.5]])

    if union:
        i0 = ix[0]
        tot = np.prod(s.IO.sensor.imSize[:, i0])
        pts = s.IP.val[:, np.isin(s.IP.img, ix)]
        if pts.shape[1] == 0:
            return 0.0, 0.0, 0.0
        crr = radial(pts, i0).max() / radial(corners(i0), i0).max()
        return _hull_area(pts) / tot, float(np.prod(pts.max(axis=1) - pts.min(axis=1)) / tot), float(crr)
    c = np.full(len(ix), np.nan)
    cr = np.full(len(ix), np.nan)
    crr = np.full(len(ix), np.nan)
    order = np.argsort(s.IP.img, kind='stable')
    start = np.searchsorted(np.asarray(s.IP.img)[order], np.arange(nImg + 1))
    for n, i in enumerate(ix):
        pts = s.IP.val[:, order[start[i]:start[i + 1]]]
        if pts.shape[1] == 0:
            continue
        tot = np.prod(s.IO.sensor.imSize[:, i])
        crr[n] = radial(pts, i).max() / radial(corners(i), i).max()
        c[n] = _hull_area(pts) / tot
        cr[n] = np.prod(pts.max(axis=1) - pts.min(axis=1)) / tot
    return c, cr, crr


def angles(s):
    """angles.m: per object point the largest angle (rad, folded to [0, pi/2]) between any two of its
    rays; 0 for one ray, NaN for none.  Points are processed in groups of equal ray count instead of the
    reference's per-point loop."""
    nOP = s.OP.val.shape[1]
    a = np.full(nOP, np.nan)
    order = np.lexsort((s.IP.img, s.IP.op))
    img = np.asarray(s.IP.img)[order]
    start = np.searchsorted(np.asarray(s.IP.op)[order], np.arange(nOP + 1))
    cnt = np.diff(start)
    a[cnt == 1] = 0.0
    for k in np.unique(cnt[cnt > 1]):
        grp = np.flatnonzero(cnt == k)
        step = max(1, int(4e6 // (k * k)))
        for g0 in range(0, len(grp), step):
            pts = grp[g0:g0 + step]
            cams = img[start[pts][:, None] + np.arange(k)[None, :]]              # (n, k)
            d = s.OP.val[:, pts][:, :, None] - s.EO.val[0:3][:, cams]            # (3, n, k)
            dn = d / np.sqrt((d ** 2).sum(axis=0))
            ip = np.clip(np.einsum('cnk,cnl->nkl', dn, dn), -1, 1)
            a[pts] = np.arccos(np.abs(ip).reshape(len(pts), -1).min(axis=1))     # max acos = acos of the min
    return a


# ----------------------------------------------------------------------------- struct helpers
def _io_uniq(s):
    """IO.struct.uniq (parseblockvariant.m): first image of every distinct IO block column."""
    u = getattr(s.IO.struct, 'uniq', None)
    if u is not None:
        return np.asarray(u, bool)
    blk = np.asarray(s.IO.struct.block)
    _, first = np.unique(blk.T, axis=0, return_index=True)
    out = np.zeros(blk.shape[1], bool)
    out[first] = True
    return out


def _io_no(s):
    no = getattr(s.IO.struct, 'no', None)
    if no is not None:
        return np.asarray(no)
    blk = np.asarray(s.IO.struct.block)
    _, first, inv = np.unique(blk.T, axis=0, return_index=True, return_inverse=True)
    rank = np.argsort(np.argsort(first))               # number blocks in order of first appearance
    return rank[np.ravel(inv)] + 1


def _io_simple(s):
    v = getattr(s.IO.struct, 'isSimple', None)
    if v is not None:
        return np.asarray(v, bool)
    blk = np.asarray(s.IO.struct.block)
    return np.array([len(np.unique(blk[:, i])) == 1 for i in range(blk.shape[1])])


def _io_names(s):
    return ['cc', 'px', 'py', 'as', 'sk'] + ['K%d' % (i + 1) for i in range(s.IO.model.nK)] + \
        ['P%d' % (i + 1) for i in range(s.IO.model.nP)]


def _get(o, path, default):
    for k in path.split('.'):
        o = getattr(o, k, None)
        if o is None:
            return default
    return o


def _ray_counts(s):
    return np.bincount(s.IP.op, minlength=s.OP.val.shape[1])


# ----------------------------------------------------------------------------- the result file
_NONFINITE = re.compile(r'\b(nan|inf)\b')      # MATLAB spells them NaN / Inf


class _Out:
    """Indented line writer; `table` is the reference's pretty_print (bundle_result_file.m:928-944)."""

    def __init__(self):
        self.lines = []

    def __call__(self, level, text):
        self.lines.append('   ' * level + _NONFINITE.sub(lambda m: {'nan': 'NaN', 'inf': 'Inf'}[m.group(0)], text))

    def table(self, level, rows, minLen=math.inf, maxLen=-math.inf):
        longest = max(len(r[0]) for r in rows)
        width = max(min(minLen, longest), maxLen) + 1
        for name, text in rows:
            self(level, name + ' ' * max(int(width - len(name)), 0) + text)


def _g(v):
    return '%g' % v


def _pstats(out, level, a, ids, labels, kind, nanmean=False):
    """Minimum / Maximum / Average of the ray angles of one point class."""
    def lab(i):
        return ', label %s' % labels[i] if labels is not None and labels[i] else ''
    if np.all(np.isnan(a)):
        mni = mxi = 0
    else:
        mni, mxi = int(np.nanargmin(a)), int(np.nanargmax(a))
    out(level, 'Minimum: %.1f degrees (%s %d%s)' % (a[mni], kind, ids[mni], lab(mni)))
    out(level, 'Maximum: %.1f degrees (%s %d%s)' % (a[mxi], kind, ids[mxi], lab(mxi)))
    if nanmean:
        avg = np.nan if np.all(np.isnan(a)) else np.nanmean(a)
    else:
        avg = np.mean(a)
    out(level, 'Average: %.1f degrees' % avg)


def _nanarg(v, fn):
    v = np.asarray(v, float)
    return 0 if np.all(np.isnan(v)) else int(fn(v))


def _ray_block(out, rays, label, zero_note=False):
    if rays is None:
        out(3, '%s ray count: -' % label)
        return
    nz = rays[rays != 0] if zero_note else rays
    n0 = int(np.count_nonzero(rays == 0))
    if zero_note and n0 > 0:
        out(3, '%s ray count: %dx0, %d-%d (%.1f avg)' % (label, n0, nz.min(), nz.max(), nz.mean()))
    else:
        out(3, '%s ray count: %d-%d (%.1f avg)' % (label, nz.min(), nz.max(), nz.mean()))
    h = np.bincount(rays)
    for k in np.flatnonzero(h):
        out(4, '%d points with %d rays.' % (h[k], k))


def _point_table(out, s, cIx, kind):
    ids = np.asarray(s.OP.id)[cIx]
    labels = _get(s, 'OP.label', None)
    lab = [labels[i] if labels is not None else '' for i in cIx]
    rays = _ray_counts(s)[cIx]
    pos0, std0 = s.prior.OP.val[:, cIx], s.prior.OP.std[:, cIx]
    pos1, std1 = s.OP.val[:, cIx], s.post.std.OP[:, cIx]
    h7 = ('id', 'x', 'y', 'z', 'stdx', 'stdy', 'stdz')
    out(3, 'Prior')
    out(3, '%6s, %8s, %8s, %8s, %8s, %8s, %8s, %s' % (h7 + ('label',)))
    for n in range(len(cIx)):
        out(3, '%6d, %8.3f, %8.3f, %8.3f, %8.3g, %8.3g, %8.3g, %s'
            % ((ids[n],) + tuple(pos0[:, n]) + tuple(std0[:, n]) + (lab[n],)))
    out(3, 'Posterior')
    out(3, '%6s, %8s, %8s, %8s, %8s, %8s, %8s, %4s, %s' % (h7 + ('rays', 'label')))
    for n in range(len(cIx)):
        out(3, '%6d, %8.3f, %8.3f, %8.3f, %8.3g, %8.3g, %8.3g, %4d, %s'
            % ((ids[n],) + tuple(pos1[:, n]) + tuple(std1[:, n]) + (rays[n], lab[n])))
    out(3, 'Diff (pos=abs diff, std=rel diff)')
    out(3, '%6s, %8s, %8s, %8s, %8s, %8s, %8s, %8s, %8s, %4s, %s'
        % ('id', 'x', 'y', 'z', 'xy', 'xyz', 'stdx', 'stdy', 'stdz', 'rays', 'label'))
    posd = pos1 - pos0
    eps = np.finfo(float).eps
    stdd = ((std1 + eps) / (std0 + eps) - 1) * 100
    for n in range(len(cIx)):
        out(3, '%6d, %8.3f, %8.3f, %8.3f, %8.3f, %8.3f, %7.1f%%, %7.1f%%, %7.1f%%, %4d, %s'
            % ((ids[n],) + tuple(posd[:, n]) + (np.linalg.norm(posd[0:2, n]), np.linalg.norm(posd[:, n]))
               + tuple(stdd[:, n]) + (rays[n], lab[n])))
    out(3, '%s point delta' % kind)
    dn = np.sqrt((posd ** 2).sum(axis=0))
    i = _nanarg(dn, np.nanargmax)
    if kind == 'Ctrl':
        fmt = lambda n: (lab[n] + ', ') if lab[n] else ''
        out(4, 'Max: %.3f ou (%spt %d)' % (dn[i], fmt(i), ids[i]))
    else:
        fmt = lambda n: lab[n] + ', '
        out(4, 'Max: %.3f ou (%s, pt %d)' % (dn[i], lab[i], ids[i]))
    out(4, 'Max X,Y,Z')
    for k in range(3):
        j = _nanarg(np.abs(posd[k]), np.nanargmax)
        out(5, '%c: %.3f ou (%spt %d)' % ('XYZ'[k], abs(posd[k, j]), fmt(j), ids[j]))
    out(4, 'RMS: %.3f ou (from %d items)' % (np.sqrt(np.mean(dn ** 2)), len(dn)))


def bundle_result_file(s, e, f=None, cov=_device_cov):
    """bundle_result_file.m: write the text result file of a bundle run to path f (or only build it when
    f is None) and return (s, lines) with s.post.std / s.post.cov filled in (:139-173).  `cov` is the
    covariance provider, `bundle_cov` of the device path by default."""
    out = _Out()
    nImg, nOP = s.EO.val.shape[1], s.OP.val.shape[1]
    NC = s.IO.val.shape[0]
    nK, nP = s.IO.model.nK, s.IO.model.nP
    camUnit = _get(s, 'IO.model.camUnit', 'mm')
    out(0, 'Damped Bundle Adjustment Toolbox result file')
    out(1, 'Project')
    out(2, 'Name             : %s' % _get(s, 'proj.title', ''))
    out(2, 'Computation UUID : %s' % _get(s, 'proj.UUID', ''))
    for key, tag in (('fileName', 'Input file name  '), ('cptFile', 'Ctrl pt file     '), ('EOfile', 'EO file          ')):
        if _get(s, 'proj.' + key, ''):
            out(2, '%s: %s' % (tag, _get(s, 'proj.' + key, '')))

    # -- problems found by bundle (:57-92)
    out(1, 'Problems and suggestions:')
    out(2, 'Project Problems:')
    ws, wn = e.weakness.structural, e.weakness.numerical
    if ws is None:
        out(3, 'Structural rank: ok.')
    else:
        out(3, 'Structural rank: %d (deficiency: %d)' % (ws.rank, ws.deficiency))
        out(4, 'DMPERM suggests the following parameters have problems:')
        for p in ws.suspectedParams:
            out(5, p)
    if wn is None or wn.deficiency == 0:
        out(3, 'Numerical rank: ok.')
    elif isinstance(wn.rank, float) and math.isnan(wn.rank):
        out(3, 'Numerical rank: not tested.')
    else:
        out(3, 'Numerical rank: %d (deficiency: %d)' % (wn.rank, wn.deficiency))
        out(4, 'Null-space suggest the following parameters are part of the problem:')
        for i, sp_ in enumerate(getattr(wn, 'suspectedParams', [])):
            out(5, 'Vector %d (eigenvalue %g):' % (i + 1, wn.d[i]))
            for p, v in zip(sp_.params, sp_.values):
                out(6, '(%s, %.3g)' % (p, v))

    # -- covariances, correlations, significance (:94-178)
    t0 = time.process_time()
    iio, jio, vio, CIO = high_io_correlations(s, e, CORR_THRESHOLD, cov)
    tCIO = time.process_time() - t0
    t0 = time.process_time()
    ieo, jeo, keo, veo, CEO = high_eo_correlations(s, e, CORR_THRESHOLD, cov)
    tCEO = time.process_time() - t0
    t0 = time.process_time()
    COP = cov(s, e, 'COP')
    tCOP = time.process_time() - t0
    iop, jop, kop, vop = high_op_correlations(s, e, CORR_THRESHOLD, COP)
    bIO, bEO, bOP = _diag_blocks(CIO, NC), _diag_blocks(CEO, 6), _diag_blocks(COP, 3)
    dg = lambda B: np.sqrt(np.diagonal(B, axis1=1, axis2=2)).T.copy()
    NS = type(s.post)
    s.post.std = getattr(s.post, 'std', None) or NS()
    s.post.cov = getattr(s.post, 'cov', None) or NS()
    if getattr(s.post, 'sensor', None) is None:             # bundle.m:360-366
        aspect = np.ones((2, nImg))
        aspect[0] = 1 + s.IO.val[3]
        s.post.sensor = NS(imSize=s.IO.sensor.imSize, pxSize=s.IO.sensor.pxSize * aspect,
                           ssSize=s.IO.sensor.imSize * s.IO.sensor.pxSize * aspect)
    s.post.std.IO, s.post.std.EO, s.post.std.OP = dg(bIO), dg(bEO), dg(bOP)
    s.post.cov.CIO, s.post.cov.CEO, s.post.cov.COP = CIO, CEO, COP
    s.post.cov.IO, s.post.cov.EO, s.post.cov.OP = (np.moveaxis(B, 0, 2) for B in (bIO, bEO, bOP))
    pk, pp, pb, pkc = test_distortion_params(s, e, bIO)
    with np.errstate(invalid='ignore'):
        lowsig = bool(np.any(np.concatenate([pk, pp, pb]) < SIG_THRESHOLD))
    nprob = int(len(vio) > 0) + int(len(veo) > 0) + int(len(vop) > 0) + int(lowsig) + int(e.code != 0)
    out(2, 'Problems related to the processing: (%d)' % nprob)
    if e.code != 0:
        out(3, 'Bundle failed with code %d (see below for details).' % e.code)
    if len(iio):
        out(3, 'One or more of the camera parameter has a high correlation (see below).')
    if len(ieo):
        out(3, 'One or more of the camera station parameters has a high correlation (see below).')
    if len(iop):
        out(3, 'One or more of the object point coordinates has a high correlation.')
    if lowsig:
        out(3, 'One or more estimated lens and/or affine distortion coefficients failed significance test (see below).')

    # -- run information (:180-270)
    out(1, 'Information from last bundle')
    msgs = ['Too many iterations', 'Normal matrix is singular', 'No step length found by the line search',
            'Normal matrix is structurally rank deficient']
    if e.code == 0:
        status = 'OK'
    elif abs(e.code) <= len(msgs):
        status = 'fail (code %d: %s)' % (e.code, msgs[abs(e.code) - 1])
    else:
        status = 'fail (code %d: unknown code)' % e.code
    rix = s.post.res.ix
    out.table(2, [
        ('Last Bundle Run:', getattr(e, 'dateStamp', time.strftime('%d-%b-%Y %H:%M:%S'))),
        ('DBAT version:', getattr(e, 'version', 'dbat_b200')),
        ('MATLAB version:', 'n/a (Python %s)' % platform.python_version()),
        ('Host system:', '%s (endian=%s)' % (platform.machine(), 'L' if np.little_endian else 'B')),
        ('Host name:', os.environ.get('HOST') or platform.node() or '<unknown>'),
        ('Status:', status),
        ('Sigma0:', _g(e.s0)),
        ('Sigma0 (pixels):', _g(np.ravel(s.post.sigmas)[0])),
        ('Redundancy', '%d' % e.redundancy),
        ('Number of params:', '%d (%d IO, %d EO, %d OP)' % (
            e.numParams, np.count_nonzero(s.IO.struct.leading), np.count_nonzero(s.EO.struct.leading),
            np.count_nonzero(s.bundle.est.OP))),
        ('Number of observations:', '%d (%d IP, %d IO, %d EO, %d OP)' % (
            e.numObs, len(rix.IP), len(rix.IO), len(rix.EO), len(rix.OP)))])
    offon = ('off', 'on')
    out(2, 'Processing options:')
    out.table(3, [
        ('Orientation:', 'on'), ('Global optimization:', 'on'),
        ('Calibration:', offon[int(np.any(s.bundle.est.IO))]), ('Constraints:', 'off'),
        ('Maximum # of iterations:', '%d' % e.maxIter), ('Convergence tolerance:', _g(e.convTol)),
        ('Termination criteria:', ('relative', 'absolute')[int(bool(e.absTerm))]),
        ('Singular test:', offon[int(bool(e.singularTest))]),
        ('Chirality veto:', offon[int(bool(getattr(e, 'chirality', False)))]),
        ('Damping:', e.damping.name), ('Camera unit (cu):', camUnit),
        ('Object space unit (ou):', _get(s, 'proj.objUnit', 'm')),
        ('Initial value comment:', _get(s, 'proj.x0desc', ''))])
    out(2, 'Total error:')
    out.table(3, [('Number of stages:', '1'), ('Number of iterations:', '%d' % e.usedIters),
                  ('First error:', _g(e.res[0])), ('Last error:', _g(e.res[-1]))])
    out(2, 'Execution times (s):')
    out.table(3, [('Bundle:', '%.2f' % e.time), ('Post-cov prep:', '%.2f' % 0.0), ('Post-cov CIO:', '%.2f' % tCIO),
                  ('Post-cov CEO:', '%.2f' % tCEO), ('Post-cov COP:', '%.2f' % tCOP)])
    out(2, 'Lens distortion models:')
    dm = np.unique(s.IO.model.distModel)
    if len(dm) == 1 and dm[0] > 0:
        out(3, 'Backward (Photogrammetry) model %d' % dm[0])
    elif len(dm) == 1 and dm[0] < 0:
        out(3, 'Forward (Computer Vision) model %d' % -dm[0])
    else:
        out(3, 'Mixed Forward/Backward')

    # -- cameras (:272-443)
    corrStr = 'Correlations over %g%%:' % (CORR_THRESHOLD * 100)
    out(2, 'Cameras:')
    est = s.bundle.est.IO
    selfCal = est.any(axis=0)
    names = _io_names(s)
    if selfCal.all():
        allP, anyP = est.all(axis=1), est.any(axis=1)
        selfCalStr = 'yes (%s)' % ' '.join(n for n, a in zip(names, allP) if a) if np.array_equal(allP, anyP) \
            else 'yes (mixed parameters)'
    elif not selfCal.any():
        selfCalStr = 'no'
    else:
        selfCalStr = 'mixed'
    out(3, 'Calibration: %s' % selfCalStr)
    u = camUnit
    ssSize, imSize, pxPost = s.post.sensor.ssSize, s.post.sensor.imSize, s.post.sensor.pxSize
    # (heading, unit, source): source >= 0 is a row of IO.val, ('s', array) a derived sensor quantity
    entries = [('Camera Constant', u, 0), ('px - principal point x', u, 1), ('py - principal point y', u, 2),
               ('Format width', u, ('s', ssSize[0])), ('Format height', u, ('s', ssSize[1]))]
    entries += [('K%d - radial distortion %d' % (i + 1, i + 1), '%s^(-%d)' % (u, 2 * i + 3), 5 + i) for i in range(nK)]
    entries += [('P%d - decentering distortion %d' % (i + 1, i + 1), '%s^(-3)' % u, 5 + nK + i) for i in range(nP)]
    entries += [('as - off-unit aspect parameter', '', 3), ('sk - skew', '', 4),
                ('Image width', 'px', ('s', imSize[0])), ('Image height', 'px', ('s', imSize[1])),
                ('X resolution', 'px/' + u, ('s', imSize[0] / ssSize[0])),
                ('Y resolution', 'px/' + u, ('s', imSize[1] / ssSize[1])),
                ('Pixel width', u, ('s', pxPost[0])), ('Pixel height', u, ('s', pxPost[1]))]
    IOdata = s.IO.val.copy()
    IOdata[2] = -IOdata[2]                                  # presentation signs (:358-359)
    IOdata[5:] = -IOdata[5:]
    sig = np.full((NC, nImg), np.nan)
    cum = np.full((NC, nImg), np.nan)
    sig[3:5], sig[5:5 + nK], sig[5 + nK:5 + nK + 2] = pb, pk, pp
    cum[5:5 + nK] = pk                                      # :350-353 prints pk here, not pkc
    ioNo = _io_no(s)
    simple = _io_simple(s)
    both = (np.vstack([iio, jio]), np.vstack([jio, iio]), np.concatenate([vio, vio]))
    for i in np.flatnonzero(_io_uniq(s)):
        pad = len('Significance:') if selfCal[i] else len('Value:')
        out(3, 'Camera%d (%s)' % (ioNo[i], 'simple' if simple[i] else 'mixed'))
        out(4, 'Lens distortion model:')
        m = s.IO.model.distModel[i]
        out(5, 'Backward (Photogrammetry) model %d' % m if m > 0 else 'Forward (Computer Vision) model %d' % -m)
        for head, unit, src in entries:
            out(4, head + ':')
            if isinstance(src, tuple):
                val, sd, sg, cs = src[1][i], np.nan, np.nan, np.nan
            else:
                val, sd, sg, cs = IOdata[src, i], s.post.std.IO[src, i], sig[src, i], cum[src, i]
            rows = [('Value:', '%g %s' % (val, unit))]
            if not np.isnan(sd) and sd != 0:
                rows.append(('Deviation:', '%.3g %s' % (sd, unit)))
            if not np.isnan(sg):
                rows.append(('Significance:', 'p=%.2f' % sg))
            if not np.isnan(cs):
                rows.append(('Cumulative significance:', 'p=%.2f' % cs))
            if selfCal[i] and not isinstance(src, tuple) and len(both[2]):
                hit = np.flatnonzero((both[0][:, 1] == i) & (both[0][:, 0] == src))
                if len(hit):
                    txt = ','.join(
                        ' %s:%.1f%%' % (names[both[1][h, 0]], both[2][h] * 100) if both[1][h, 1] == i else
                        ' %s(cam%d):%.1f%%' % (names[both[1][h, 0]], both[1][h, 1] + 1, both[2][h] * 100)
                        for h in hit)
                    rows.append((corrStr, txt + '.'))
            out.table(5, rows, pad, pad)
        ss0 = _get(s, 'IO.sensor.ssSize', None)
        ss0 = s.IO.sensor.imSize[:, i] * s.IO.sensor.pxSize[:, i] if ss0 is None else ss0[:, i]
        whd = np.array([ss0[0], ss0[1], np.linalg.norm(ss0)])
        aov = 2 * np.arctan(whd / (2 * s.IO.val[0, i])) * 180 / np.pi
        out(3, 'Rated angle of view (h,v,d): (%.0f, %.0f, %.0f) deg' % tuple(aov))
        w, h = s.IO.sensor.imSize[:, i]
        cx = np.array([0.5, 0.5, w + 0.5, w + 0.5])
        cy = np.array([0.5, h + 0.5, h + 0.5, 0.5])
        xr = cx * s.IO.sensor.pxSize[0, i] - s.IO.val[1, i]
        yr = cy * s.IO.sensor.pxSize[1, i] + s.IO.val[2, i]
        r2 = xr ** 2 + yr ** 2
        Kv = np.zeros(3)
        Kv[:min(nK, 3)] = s.IO.val[5:5 + min(nK, 3), i]
        rad = Kv[0] * r2 + Kv[1] * r2 ** 2 + Kv[2] * r2 ** 3
        P1, P2 = (s.IO.val[5 + nK, i], s.IO.val[6 + nK, i]) if nP >= 2 else (0.0, 0.0)
        xc = xr * rad + P1 * (r2 + 2 * xr ** 2) + 2 * P1 * xr * yr       # as written in :428-429
        yc = yr * rad + P2 * (r2 + 2 * yr ** 2) + 2 * P2 * xr * yr
        mx = np.max(np.abs(xc) + np.abs(yc))
        out(3, 'Largest distortion: %.2g %s (%.1f px, %.1f%% of half-diagonal)'
            % (mx, camUnit, mx / s.IO.sensor.pxSize[0, i], mx / (whd[2] / 2) * 100))

    # -- camera stations (:445-500)
    out(2, 'Precisions / Standard Deviations:')
    out(3, 'Photograph Standard Deviations:')
    heads = ('Omega', 'Phi', 'Kappa', 'Xc', 'Yc', 'Zc')
    units = ('deg',) * 3 + ('ou',) * 3
    rows_eo = (3, 4, 5, 0, 1, 2)
    present = {r: n for n, r in enumerate(rows_eo)}
    scale = np.array([180 / np.pi] * 3 + [1.0] * 3)
    a_i, a_j = np.concatenate([ieo, jeo]), np.concatenate([jeo, ieo])
    a_k, a_v = np.concatenate([keo, keo]), np.concatenate([veo, veo])
    by_img = {}
    for n in range(len(a_k)):
        by_img.setdefault((int(a_k[n]), int(a_i[n])), []).append(n)
    enames = _get(s, 'EO.name', None)
    pad = len('Deviation:')
    for i in range(nImg):
        out(4, 'Photo %d: %s' % (i + 1, enames[i] if enames is not None else ''))
        vals = scale * s.EO.val[list(rows_eo), i]
        sd = scale * s.post.std.EO[list(rows_eo), i]
        for j in range(6):
            out(5, heads[j] + ':')
            rows = [('Value:', '%.6f %s' % (vals[j], units[j]))]
            if sd[j] != 0:
                rows.append(('Deviation:', '%.3g %s' % (sd[j], units[j])))
            hit = by_img.get((i, rows_eo[j]))
            if hit:
                rows.append((corrStr, ','.join(' %s:%.1f%%' % (heads[present[int(a_j[h])]], a_v[h] * 100)
                                               for h in hit) + '.'))
            out.table(6, rows, pad, pad)

    # -- quality (:502-560)
    out(1, 'Quality')
    out(2, 'Photographs')
    out.table(3, [('Total number:', '%d' % (len(enames) if enames is not None else nImg)), ('Numbers used:', '%d' % nImg)])
    out(2, 'Cameras')
    uq = _io_uniq(s)
    out(3, 'Total number: %d (%d simple, %d mixed)' % (uq.sum(), (uq & simple).sum(), (uq & ~simple).sum()))
    pc = lambda v: int(np.floor(v * 100 + 0.5))
    for i in np.flatnonzero(uq):
        out(3, 'Camera%d:' % ioNo[i])
        mine = np.flatnonzero(ioNo == ioNo[i])
        out.table(4, [('Calibration:', ('<not available>', 'yes')[int(est[:, i].any())]),
                      ('Number of photos using camera:', '%d' % len(mine))])
        c, cr, crr = coverage(s, mine)
        uc, ucr, ucrr = coverage(s, mine, True)
        out(4, 'Photo point coverage:')
        line = lambda v, uv: '%d%%-%d%% (%d%% average, %d%% union)' % (
            pc(np.nanmin(v)), pc(np.nanmax(v)), pc(np.nanmean(v)), pc(uv))
        out.table(5, [('Rectangular:', line(cr, ucr)), ('Convex hull:', line(c, uc)), ('Radial:', line(crr, ucrr))])
    out(2, 'Photo Coverage')
    out(3, 'Reference points outside calibrated region:')
    for i in np.flatnonzero(uq):
        out(4, 'Camera %d: %s' % (ioNo[i], 'none' if est[:, i].any() else '<not available>'))

    # -- point measurements (:562-630)
    isCtrl = np.asarray(_get(s, 'prior.OP.isCtrl', np.zeros(nOP, bool)), bool)
    isCheck = np.asarray(_get(s, 'prior.OP.isCheck', np.zeros(nOP, bool)), bool)
    isOP = ~isCtrl & ~isCheck
    rays = _ray_counts(s)
    out(2, 'Point Measurements')
    out(3, 'Number of control pts: %d' % isCtrl.sum())
    out(3, 'Number of check pts: %d' % isCheck.sum())
    out(3, 'Number of object pts: %d' % isOP.sum())
    _ray_block(out, rays[isCtrl] if isCtrl.any() else None, 'CP', True)
    _ray_block(out, rays[isCheck] if isCheck.any() else None, 'CCP')
    _ray_block(out, rays[~isCtrl] if (~isCtrl).any() else None, 'OP')

    # -- residuals (:632-672)
    rms, pt = bundle_residuals(s, e)
    ids = np.asarray(s.OP.id)
    out(2, 'Point Marking Residuals')
    out(3, 'Overall point RMS: %.3f pixels' % rms)
    out(3, 'Mark point residuals:')
    k = _nanarg(pt, np.nanargmax)
    out(4, 'Maximum: %.3f pixels (OP %d on photo %d)' % (pt[k], ids[s.IP.op[k]], s.IP.img[k] + 1))
    with np.errstate(divide='ignore', invalid='ignore'):
        meanOP = np.sqrt(np.bincount(s.IP.op, pt ** 2, nOP) / rays)
        nPh = np.bincount(s.IP.img, minlength=nImg)
        meanPh = np.sqrt(np.bincount(s.IP.img, pt ** 2, nImg) / nPh)
    out(3, 'Object point residuals (RMS over all images of a point):')
    a, b = _nanarg(meanOP, np.nanargmin), _nanarg(meanOP, np.nanargmax)
    out(4, 'Minimum: %.3f pixels (OP %d over %d images)' % (meanOP[a], ids[a], rays[a]))
    out(4, 'Maximum: %.3f pixels (OP %d over %d images)' % (meanOP[b], ids[b], rays[b]))
    out(3, 'Photo residuals (RMS over all points in an image):')
    a, b = _nanarg(meanPh, np.nanargmin), _nanarg(meanPh, np.nanargmax)
    out(4, 'Minimum: %.3f pixels (photo %d over %d points)' % (meanPh[a], a + 1, nPh[a]))
    out(4, 'Maximum: %.3f pixels (photo %d over %d points)' % (meanPh[b], b + 1, nPh[b]))

    # -- point precision (:674-722)
    out(2, 'Point Precision')
    v = s.post.std.OP ** 2
    v[~s.bundle.est.OP] = np.nan
    tStd = np.sqrt(v.sum(axis=0))
    out(3, 'Total standard deviation (RMS of X/Y/Z std):')
    a, b = _nanarg(tStd, np.nanargmin), _nanarg(tStd, np.nanargmax)
    out(4, 'Minimum: %.2g (OP %d)' % (tStd[a], ids[a]))
    out(4, 'Maximum: %.2g (OP %d)' % (tStd[b], ids[b]))
    for k in range(3):
        b = _nanarg(v[k], np.nanargmax)
        out(3, 'Maximum %c standard deviation: %.2g (OP %d)' % ('XYZ'[k], np.sqrt(v[k, b]), ids[b]))
    out(3, 'Points with high correlations')
    out(4, 'Points with correlation above 95%%: %d' % np.count_nonzero(np.abs(vop) > 0.95))
    out(4, 'Points with correlation above 99%%: %d' % np.count_nonzero(np.abs(vop) > 0.99))
    if np.count_nonzero(np.abs(vop) > 0.95):
        out(4, 'Points with highest correlations:')
        printed = []
        for n in np.argsort(-np.abs(vop), kind='stable'):
            if len(printed) >= 5:
                break
            if kop[n] not in printed:
                printed.append(kop[n])
                out(5, 'Points %d: %.2f' % (kop[n] + 1, 100 * vop[n]))

    # -- ray angles (:724-810)
    out(2, 'Point Angles')
    ang = angles(s) * 180 / np.pi
    labels = _get(s, 'OP.label', None)
    out(3, 'CP')
    cpIx = np.flatnonzero(isCtrl)
    if np.any(rays[cpIx] == 0):
        out(4, 'Ignoring %d CP with 0 rays.' % np.count_nonzero(rays[cpIx] == 0))
    if np.any(cpIx[rays[cpIx] > 0]):
        _pstats(out, 4, ang[cpIx], ids[cpIx], [labels[i] for i in cpIx] if labels is not None else None, 'CP', True)
    else:
        for t in ('Minimum', 'Maximum', 'Average'):
            out(4, t + ': -')
    out(3, 'CCP')
    if isCheck.any():
        ccIx = np.flatnonzero(isCheck)
        _pstats(out, 4, ang[ccIx], ids[ccIx], [labels[i] for i in ccIx] if labels is not None else None, 'CCP')
    else:
        for t in ('Minimum', 'Maximum', 'Average'):
            out(4, t + ': -')
    out(3, 'OP')
    if isOP.any():
        opIx = np.flatnonzero(isOP)
        aOP, idOP = ang[opIx], ids[opIx]
        _pstats(out, 4, aOP, idOP, None, 'OP')
        out(4, 'Smallest angles (ID, angle [deg], vis in cameras)')
        o = np.argsort(aOP, kind='stable')
        srt = aOP[o]
        lim = min(srt[min(3, len(srt)) - 1] * 1.1 + 0.1, 80)
        nPts = min(max(int(np.count_nonzero(srt < lim)), 3), len(srt))
        for n in range(nPts):
            # :805 indexes IP.vis with the position inside the OP subset, not the point's own row;
            # kept, so that a result file diffs clean against the reference's
            cams = np.sort(s.IP.img[s.IP.op == o[n]]) + 1
            out(5, '%6d: %5.2f (%s)' % (idOP[o[n]], srt[n], ' '.join('%4d' % c for c in cams)))
    else:
        for t in ('Minimum', 'Maximum', 'Average'):
            out(4, t + ': -')

    # -- control and check points (:812-924)
    out(2, 'Ctrl measurements')
    if isCtrl.any():
        _point_table(out, s, np.flatnonzero(isCtrl), 'Ctrl')
    else:
        out(3, 'none')
    out(2, 'Check measurements')
    if isCheck.any():
        _point_table(out, s, np.flatnonzero(isCheck), 'Check')
    else:
        out(3, 'none')
    out.lines.append('End of result file')
    if f is not None:
        with open(f, 'wt') as fid:
            fid.write('\n'.join(out.lines) + '\n')
    return s, out.lines


# ----------------------------------------------------------------------------- pre-bundle statistics
def camangles(s):
    """camangles.m: per image the largest angle (rad, folded to [0, pi/2]) between any two of its rays."""
    nImg = s.EO.val.shape[1]
    a = np.zeros(nImg)
    for i in range(nImg):
        pp = s.OP.val[:, np.asarray(s.IP.op)[np.asarray(s.IP.img) == i]]
        if pp.shape[1] < 2:
            continue
        d = s.EO.val[0:3, i:i + 1] - pp
        dn = d / np.sqrt((d ** 2).sum(axis=0))
        lo = 1.0
        for c0 in range(0, dn.shape[1], 2048):                   # min |cos| over all pairs, in row blocks
            lo = min(lo, float(np.abs(np.clip(dn[:, c0:c0 + 2048].T @ dn, -1, 1)).min()))
        a[i] = np.arccos(lo)
    return a


def _sturges_edges(x):
    """Bin edges of MATLAB's histcounts(x,'BinMethod','sturges'): ceil(log2(n)+1) bins widened to a 'nice'
    width (1, 2, 3, 5 or 10 times a power of ten) with the left edge on a multiple of it."""
    x = np.asarray(x, float)
    nbins = max(int(np.ceil(np.log2(len(x)) + 1)), 1)
    xmin, xmax = x.min(), x.max()
    rng = xmax - xmin
    scale = max(abs(xmin), abs(xmax))
    raw = max(rng / nbins, np.spacing(scale))
    if rng <= max(np.sqrt(np.spacing(scale)), np.finfo(float).tiny):
        return np.array([np.floor(xmin) - 0.5 * 0 + 0.0, np.floor(xmin) + 1.0])
    p10 = 10.0 ** np.floor(np.log10(raw))
    rel = raw / p10
    width = p10 * (1 if rel < 1.5 else 2 if rel < 2.5 else 3 if rel < 4 else 5 if rel < 7.5 else 10)
    left = min(width * np.floor(xmin / width), xmin)
    n = max(1, int(np.ceil((xmax - left) / width)))
    return left + width * np.arange(n + 1)


def _hist_centers(x, centers):
    """MATLAB hist(x, centers): bins split at the midpoints, open-ended at both ends."""
    mid = (centers[:-1] + centers[1:]) / 2
    return np.bincount(np.digitize(x, mid, right=True), minlength=len(centers))


def writestats(s, fName=None, desc=''):
    """writestats.m: pre-bundle statistics of a project - image ray counts and ray angles, control and
    object point ray counts and angles with their histograms and worst cases.  Returns (s, lines) with
    s.camRayAng / s.rayAng (degrees) stored like the reference does."""
    digits = lambda v: int(np.floor(np.log10(v))) + 1 if v > 0 else 1
    L = []
    w = L.append
    nImg, nOP = s.EO.val.shape[1], s.OP.val.shape[1]
    isCtrl = np.asarray(s.prior.OP.isCtrl, bool)
    nCp = int(isCtrl.sum())
    w(desc)
    w('')
    w('Project file: %s' % _get(s, 'proj.fileName', ''))
    w('')
    w('Execution time stamp: %s' % time.strftime('%Y-%m-%d %H:%M:%S'))
    w('')
    w('Total # OP          : %d' % (nOP - nCp))
    w('Total # CP          : %d' % nCp)
    w('Total # cams        : %d' % nImg)
    w('Total # image marks : %d' % s.IP.val.shape[1])
    w('Project units       : %s' % _get(s, 'proj.objUnit', 'm'))
    w('')
    w('Project images: no (id), shortened label, name:')
    labels = list(_get(s, 'EO.label', None) or s.EO.name)
    labelLen = max(len(l) for l in labels)
    if labelLen > 8:                                              # drop the common leading part of long labels
        shortest = min(len(l) for l in labels)
        mat = np.array([list(l.ljust(labelLen)) for l in labels])
        diff = np.flatnonzero(~np.all(mat == mat[0], axis=0))
        neq = (diff[0] + 1) if len(diff) else labelLen + 1
        if shortest < neq:
            cut = shortest - 4
        else:
            cut = neq - 4
            if cut > 0:
                na = [k for k, ch in enumerate(labels[0][cut:cut + 3]) if not ch.isalnum()]
                if na:
                    cut += na[-1] + 1
        if cut > 4:
            labels = [l[cut:] for l in labels]
            labelLen = max(len(l) for l in labels)
    eid = np.asarray(_get(s, 'EO.id', np.arange(1, nImg + 1)))
    dNo, dId = digits(nImg), digits(max(int(eid.max()), 1))
    imDir = _get(s, 'proj.imDir', '')
    for i in range(nImg):
        w('  %-*d (%-*d), %-*s, %s' % (dNo, i + 1, dId, eid[i], labelLen, labels[i], os.path.join(imDir, s.EO.name[i])))
    w('')
    w('')
    w('IMAGE STATISTICS')
    cnt = np.bincount(s.IP.img, minlength=nImg)
    dC = digits(cnt.max())
    w('')
    w('Image ray count:')
    w('  min : %*d' % (dC, cnt.min()))
    w('  max : %*d' % (dC, cnt.max()))
    w('  mean: %*d' % (dC, int(np.floor(cnt.mean() + 0.5))))
    edges = _sturges_edges(cnt)
    n = np.histogram(cnt, edges)[0]
    edges = edges.copy()
    edges[-1] += 1
    dE = digits(edges.max())
    w('')
    w('Image with lowest ray count: cam no (id), label, count')
    o = np.argsort(cnt, kind='stable')
    srt = cnt[o]
    for j in range(int(np.count_nonzero(srt < srt[min(3, len(srt)) - 1] * 1.1 + 0.1))):
        w('  %*d (%*d), %-*s, %*d' % (dNo, o[j] + 1, dId, eid[o[j]], labelLen, labels[o[j]], dE, srt[j]))
    w('')
    w('Image ray count histogram: nRays, nCams')
    for i in range(len(n)):
        w('  %*d-%*d: %d' % (dE, edges[i], dE, edges[i + 1] - 1, n[i]))
    if getattr(s, 'camRayAng', None) is None:
        s.camRayAng = camangles(s) * 180 / np.pi
    ca = s.camRayAng
    w('')
    w('Image ray angles (deg):')
    w('  min : %4.1f' % ca.min())
    w('  max : %4.1f' % ca.max())
    w('  mean: %4.1f' % ca.mean())
    w('')
    w('Smallest image ray angles: cam no (id), label, nRays, angle')
    o = np.argsort(ca, kind='stable')
    srt = ca[o]
    for j in range(int(np.count_nonzero(srt < srt[min(3, len(srt)) - 1] * 1.1 + 0.1))):
        w('  %*d (%*d), %-*s, %*d, %4.1f' % (dNo, o[j] + 1, dId, eid[o[j]], labelLen, labels[o[j]], dE, cnt[o[j]], srt[j]))
    aa = np.arange(0.0, 91.0, 5.0)

    def angle_hist(vals):
        h = _hist_centers(vals, aa)
        d = digits(h.max())
        for a_, c_ in zip(aa, h):
            w('  %2d, %*d' % (a_, d, c_))
    w('')
    w('Image ray angle histogram: angle, count')
    angle_hist(ca)
    rays = _ray_counts(s)
    if getattr(s, 'rayAng', None) is None:
        s.rayAng = angles(s) * 180 / np.pi
    ra = s.rayAng
    rawId = np.asarray(_get(s, 'OP.rawId', s.OP.id))
    oplab = _get(s, 'OP.label', None) or [''] * nOP
    vis_of = lambda j: ', '.join(labels[i] for i in np.sort(np.asarray(s.IP.img)[np.asarray(s.IP.op) == j]))
    for ix, title, tag in ((np.flatnonzero(isCtrl), 'CONTROL POINT STATISTICS', 'CP'),
                           (np.flatnonzero(~isCtrl), 'OBJECT POINT STATISTICS', 'OP')):
        w('')
        w('')
        w(title)
        if not np.any(ix):                                       # (`~any(ix)`: also skips a lone point with index 0)
            continue
        r = rays[ix]
        dR = digits(r.max())
        w('')
        w('%s ray count:' % tag)
        w('  min : %*d' % (dR, r.min()))
        w('  max : %*d' % (dR, r.max()))
        w('  mean: %*.1f' % (dR + 2, r.mean()))
        w('')
        w('%s ray count histogram: nRays, count' % tag)
        h = np.bincount(r)
        nz = np.flatnonzero(h)
        dn_, dc_ = digits(nz.max() + 1), digits(h.max())
        for k in nz:
            w('  %*d, %*d' % (dn_, k, dc_, h[k]))

        def worst(vals, with_angle):
            o = np.argsort(vals, kind='stable')
            srt = vals[o]
            cut = int(np.count_nonzero(srt < srt[min(3, len(srt)) - 1] * 1.1 + 0.1))
            sel = ix[o[:cut]]
            ll = max(len(oplab[j]) for j in sel)
            d1, d2 = digits(sel.max() + 1), digits(max(int(rawId[sel].max()), 1))
            d3 = digits(max(int(rays[sel].max()), 1)) if with_angle else digits(max(int(srt[cut - 1]), 1))
            w('')
            if with_angle:
                w('Smallest %s ray angles: %s no (id), %snRays, angle, (images with rays)' % (tag, tag, 'label, ' if ll else ''))
            else:
                w('%s with lowest ray count: %s no (id), %snRays, (images with rays)' % (tag, tag, 'label, ' if ll else ''))
            for j, v in zip(sel, srt[:cut]):
                head = '  %*d (%*d), ' % (d1, j + 1, d2, rawId[j]) + ('%-*s, ' % (ll, oplab[j]) if ll else '')
                body = '%*d, %4.1f, ' % (d3, rays[j], v) if with_angle else '%*d, ' % (d3, rays[j])
                w(head + body + '(%s)' % vis_of(j))
        worst(r.astype(float), False)
        w('')
        w('%s ray angles:' % tag)
        w('  min : %4.1f' % np.min(ra[ix]))
        w('  max : %4.1f' % np.max(ra[ix]))
        w('  mean: %4.1f' % np.mean(ra[ix]))
        worst(ra[ix], True)
        w('')
        w('%s ray angle histogram: angle, count' % tag)
        angle_hist(ra[ix])
    if fName is not None:
        with open(fName, 'wt') as fh:
            fh.write('\n'.join(L) + '\n')
    return s, L
