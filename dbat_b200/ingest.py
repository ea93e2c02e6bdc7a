"""Ingest: project files -> the flat DBAT struct the device path consumes.

Host mirror of the functions every reference demo calls between the file and `bundle`:

  loadpm            `code/file/loadpm.m:104-404`       PhotoModeler text export -> prob
  prob2dbatstruct   `code/misc/prob2dbatstruct.m:198-520`  prob -> struct (shared camera)
  loadcpt           `code/file/loadcpt.m:23-112`       control point table
  matchcpt          `code/misc/matchcpt.m:24-98`
  setcpt            `code/misc/setcpt.m:20-52`
  setcamvals        `code/misc/setcamvals.m:21-148`
  setcamest         `code/misc/setcamest.m:28-125`
  seteoest          `code/misc/seteoest.m:26-128`      (including the 'depend' datum)
  cleareo, clearop  `code/misc/cleareo.m`, `clearop.m`
  legacyloadeotable, matcheo, setprioreo   `code/file/legacyloadeotable.m`, `misc/matcheo.m`, `setprioreo.m`
  loadimagepts, loadctrlpts, loadimagetable, loadeotable   `code/file/*.m` (format-string table readers)
  loadpm3dtbl       `code/file/loadpm3dtbl.m`          PhotoModeler's 3-D point table (external verification)
  loadpsz, ps2pmstruct   `code/file/loadpsz.m`, `misc/ps2pmstruct.m`   PhotoScan .psz archives

The reference builds the struct with per-image loops and sparse `vis` / `ix` matrices
(`prob2dbatstruct.m:349-365`); here the image points are sorted once by (image, object point) and kept
as flat index vectors (`IP.img`, `IP.op`), which is the layout `dbat_create` uploads.  Indices are
0-based; ids are the file's.  Image-variant cameras (`individualCameras`) are not supported by the
device path and are refused.
"""
import os
import uuid as _uuid
from types import SimpleNamespace as NS

import numpy as np

from .dbatstruct import new_struct, seteoest_depend

IO_ROWS = ('cc', 'px', 'py', 'as', 'sk')


def _numbers(lines, ncol):
    if not lines:
        return np.zeros((0, ncol))
    return np.array(' '.join(lines).split(), dtype=float).reshape(len(lines), -1)


def loadpm(name, imSz=None):
    """loadpm.m: parse a PhotoModeler text export.  Five header lines (title; tolerance, iterations and
    - in newer exports - image size; default std; default camera; its std), one six-line block per image
    up to a blank line, then control points, object points and mark points, each up to a blank line.
    "Smart" points (exported with zero sigma) numbered from a range that overlaps the normal points are
    shifted above them (:385-404).  Returns prob with job / images / ctrlPts / objPts / markPts /
    rawOPids / OPlabels as the reference names them."""
    with open(name) as fh:
        lines = fh.read().split('\n')
    tol = [float(v) for v in lines[1].split()]
    if imSz is None:
        imSz = tol[2:4] if len(tol) > 2 else [np.nan, np.nan]
    job = NS(fileName=name, title=lines[0].rstrip('\r'), tol=tol[0], maxIter=tol[1],
             defStd=np.array(lines[2].split(), float), defCam=np.array(lines[3].split(), float),
             defCamStd=np.array(lines[4].split(), float), imSz=np.array(imSz, float))
    k = 5
    images = []
    while k < len(lines) and lines[k].split() and lines[k].split()[0].lstrip('-').isdigit():
        head = lines[k].split(None, 1)
        num = lambda t: np.array(t.split()[1:], float)
        cov = num(lines[k + 3]) if lines[k + 3].split() else np.full(3, np.nan)
        images.append(NS(imName=head[1].strip() if len(head) > 1 else '', outer=num(lines[k + 1]),
                         outerStd=num(lines[k + 2]), outerCov=cov, inner=num(lines[k + 4]),
                         innerStd=num(lines[k + 5]), imSz=job.imSz, id=len(images) + 1))
        k += 6
    # image labels: names without their common directory (:196-211)
    names = [im.imName.replace('\\', '/') for im in images]
    common = os.path.commonpath([os.path.dirname(n) for n in names]) if names and all('/' in n for n in names) else ''
    for im, n in zip(images, names):
        im.label = n[len(common) + 1:] if common else n

    def section(k):
        e = k
        while e < len(lines) and lines[e].strip():
            e += 1
        return lines[k:e], e + 1
    k += 1
    rows, k = section(k)
    ctrlPts = _numbers(rows, 7)
    rows, k = section(k)
    objPts = _numbers(rows, 7)
    rows, k = section(k)
    markPts = _numbers(rows, 6)
    rawOPids = objPts[:, 0].copy()
    isCtrl = np.isin(objPts[:, 0], ctrlPts[:, 0])
    OPlabels = [('%d' % i) if c else '' for i, c in zip(rawOPids, isCtrl)]
    if len(markPts) and len(objPts):
        smart = np.all(markPts[:, 4:6] == 0, axis=1)
        normId, smartId = np.unique(markPts[~smart, 1]), np.unique(markPts[smart, 1])
        split = np.flatnonzero(np.diff(objPts[:, 0]) < 0)
        if len(split) and len(normId) and len(smartId):
            shift = normId.max() + 1 - smartId.min()
            markPts[smart, 1] += shift
            isSmartObj = np.isin(objPts[:, 0], smartId)
            isSmartObj[:split[0] + 1] = False
            objPts[isSmartObj, 0] += shift
    return NS(job=job, images=images, ctrlPts=ctrlPts, cptFile='', checkPts=np.zeros((0, 7)), objPts=objPts,
              priorCamPos=np.zeros((0, 7)), rawOPids=rawOPids, OPlabels=OPlabels, markPts=markPts)


def prob2dbatstruct(prob, individualCameras=False):
    """prob2dbatstruct.m: the DBAT struct of a loaded project.  One camera shared by all images, taken
    from the export's default camera with the sign conventions of :222-232 (py, K, P negated), square
    pixels from the sensor height and the aspect folded into IO row 4 (:236-247); EO angles from the
    export's kappa, phi, omega in degrees (:262-265); legacy distortion model 1; IO fixed, EO free, OP
    free except control points with zero prior sigma; image points sorted by (image, point id); if any
    exported image sigma is zero all are set to 1 px (:367-374)."""
    if individualCameras:
        raise NotImplementedError('image-variant cameras are not supported by the device path')
    nImg = len(prob.images)
    nK, nP = 3, 2
    inner, innerStd = np.asarray(prob.job.defCam, float), np.asarray(prob.job.defCamStd, float)
    imSz = np.asarray(prob.job.imSz, float)
    IO = np.zeros(5 + nK + nP)
    IOstd = np.full(5 + nK + nP, np.nan)
    IO[0], IO[1], IO[2] = inner[0], inner[1], -inner[2]
    IOstd[0:3] = innerStd[0:3]
    IO[5:5 + nK + nP] = -inner[5:5 + nK + nP]
    IOstd[5:5 + nK + nP] = innerStd[5:5 + nK + nP]
    ssSize = inner[3:5]
    px = ssSize / imSz
    IO[3] = 1 - px[0] / px[1]
    pxSize = np.array([px[1], px[1]])
    outer = np.array([im.outer for im in prob.images]).T
    outerStd = np.array([im.outerStd for im in prob.images]).T
    EO = np.vstack([outer[0:3], np.deg2rad(outer[[5, 4, 3]])])
    EOstd = np.vstack([outerStd[0:3], np.deg2rad(outerStd[[5, 4, 3]])])
    order = np.argsort(prob.objPts[:, 0], kind='stable')
    OPid = prob.objPts[order, 0].astype(np.int64)
    nOP = len(OPid)
    OP = prob.objPts[order, 1:4].T.copy()
    priorCP = np.full((3, nOP), np.nan)
    priorCPstd = np.full((3, nOP), np.nan)
    for tab in (prob.ctrlPts, prob.checkPts):
        hit = np.isin(tab[:, 0], OPid)
        col = np.searchsorted(OPid, tab[hit, 0])
        priorCP[:, col] = tab[hit, 1:4].T
        priorCPstd[:, col] = tab[hit, 4:7].T
    isCtrl = np.isin(OPid, prob.ctrlPts[:, 0])
    isCheck = np.isin(OPid, prob.checkPts[:, 0])
    mk = prob.markPts[np.isin(prob.markPts[:, 1], OPid)]
    mstd = mk[:, 4:6].copy()
    sigmas = np.unique(mstd)
    if np.any(sigmas == 0):
        mstd[:] = 1.0
        sigmas = np.array([1.0])
    s = new_struct(np.tile(IO[:, None], (1, nImg)), EO, OP, mk[:, 2:4].T, mk[:, 0].astype(np.int64),
                   np.searchsorted(OPid, mk[:, 1]), pxSize[:, None], imSz[:, None], 1, nK, nP, mstd.T)
    names = [im.imName.replace('\\', '/') for im in prob.images]
    imDir = os.path.commonpath([os.path.dirname(n) for n in names]) if names and all('/' in n for n in names) else ''
    s.proj = NS(objUnit='m', x0desc='', title=prob.job.title, imDir=imDir + '/' if imDir else '',
                fileName=prob.job.fileName, cptFile=prob.cptFile, EOfile='', UUID=str(_uuid.uuid4()))
    s.IO.model.camUnit = 'mm'
    s.IO.sensor.ssSize = np.tile(ssSize[:, None], (1, nImg))
    s.EO.name = [n[len(imDir) + 1:] if imDir else n for n in names]
    s.EO.label = [im.label.replace('\\', '/') for im in prob.images]
    s.EO.id = np.array([im.id for im in prob.images])
    s.OP.id = OPid
    s.OP.rawId = prob.rawOPids[order].astype(np.int64)
    s.OP.label = [prob.OPlabels[i] for i in order]
    s.IP.sigmas = sigmas
    s.prior.IO.val[:] = s.IO.val
    s.prior.IO.std[:] = IOstd[:, None]
    s.prior.EO.val[:] = EO
    s.prior.EO.std[:] = EOstd
    s.prior.OP.val, s.prior.OP.std = priorCP, priorCPstd
    s.prior.OP.isCtrl, s.prior.OP.isCheck = isCtrl, isCheck
    s.prior.OP.use = np.tile(isCtrl & ~np.all(priorCPstd == 0, axis=0), (3, 1))
    s.bundle.est.IO[:] = False
    s.bundle.est.EO[:] = True
    s.bundle.est.OP = ~(priorCPstd == 0)
    return s


def loadcpt(fName, has=(True, True)):
    """loadcpt.m: control point table `id,name,x,y,z[,std...]`; 0, 1, 2 or 3 sigmas per line mean fixed,
    sigma_xyz, (sigma_xy, sigma_z), (sigma_x, sigma_y, sigma_z)."""
    hasId, hasName = has
    ids, names, pos, std = [], [], [], []
    with open(fName) as fh:
        for line in fh:
            line = line.strip()
            if not line or line[0] == '#':
                continue
            tok = [t.strip() for t in line.split(',')]
            ids.append(int(tok.pop(0)) if hasId else np.nan)
            names.append(tok.pop(0) if hasName else '')
            a = [float(t) for t in tok if t != '']
            pos.append(a[0:3])
            if len(a) == 3:
                std.append([0.0, 0.0, 0.0])
            elif len(a) == 4:
                std.append([a[3]] * 3)
            elif len(a) == 5:
                std.append([a[3], a[3], a[4]])
            elif len(a) == 6:
                std.append(a[3:6])
            else:
                raise ValueError('Bad number of items on CP line.')
    return NS(id=np.array(ids), name=names, pos=np.array(pos, float).reshape(-1, 3).T,
              std=np.array(std, float).reshape(-1, 3).T, cov=None, fileName=fName)


def matchcpt(s, pts, match='auto', check=False):
    """matchcpt.m: (i, j) - struct points i (control points, or check points with check=True) and the
    rows j of the table they match, by label, by raw id, or both ('auto': whatever the table carries)."""
    byId = match in ('id', 'both') or (match == 'auto' and np.any(~np.isnan(np.asarray(pts.id, float))))
    byLabel = match in ('label', 'both') or (match == 'auto' and any(pts.name))
    sel = np.flatnonzero(s.prior.OP.isCheck if check else s.prior.OP.isCtrl)

    def by(keys_s, keys_p):
        common, a, b = np.intersect1d(np.asarray(keys_s), np.asarray(keys_p), return_indices=True)
        return sel[a], b
    il = jl = ii = ji = None
    if byLabel:
        il, jl = by([s.OP.label[c] for c in sel], pts.name)
    if byId:
        ii, ji = by(np.asarray(s.OP.rawId)[sel], pts.id)
    if byLabel and byId:
        if len(il):
            if not (np.array_equal(il, ii) and np.array_equal(jl, ji)):
                raise ValueError('Inconsistent match')
            return il, jl
        return ii, ji
    if byLabel:
        return il, jl
    if byId:
        return ii, ji
    raise ValueError('Neither match-by-id nor match-by-label possible.')


def setcpt(s, pts, i, j, isCtrl=True):
    """setcpt.m: copy prior position / sigma / label of table rows j to points i; control points with zero
    sigma become fixed, with non-zero sigma estimated and observed; check points are estimated, not observed."""
    s.proj.cptFile = pts.fileName
    s.prior.OP.val[:, i] = pts.pos[:, j]
    s.OP.val[:, i] = pts.pos[:, j]
    s.prior.OP.std[:, i] = pts.std[:, j]
    for a, b in zip(i, j):
        if pts.name[b]:
            s.OP.label[a] = pts.name[b]
    s.prior.OP.isCtrl[i] = isCtrl
    s.prior.OP.isCheck[i] = not isCtrl
    if isCtrl:
        free = ~np.all(pts.std[:, j] == 0, axis=0)
        s.prior.OP.use[:, i] = free
        s.bundle.est.OP[:, i] = free
    else:
        s.prior.OP.use[:, i] = False
        s.bundle.est.OP[:, i] = True
    return s


def _table(path, fmt, sep=',', cmt='#'):
    """Rows of a delimited text table as dicts keyed by the format's part names."""
    parts = [p.strip() for p in fmt.split(sep)]
    rows = []
    if not os.path.exists(path) and os.path.exists(path + '.xz'):       # large tables may be stored compressed
        import lzma
        opener = lambda: lzma.open(path + '.xz', 'rt')
    else:
        opener = lambda: open(path)
    with opener() as fh:
        for n, line in enumerate(fh, 1):
            line = line.strip()
            if not line or line[0] == cmt:
                continue
            tok = [t.strip() for t in line.split(sep)]
            if len(tok) != len(parts):
                raise ValueError('%s, line %d: Wrong number of elements (got %d, expected %d)'
                                 % (path, n, len(tok), len(parts)))
            rows.append(dict(zip(parts, tok)))
    return parts, rows


def loadimagepts(fName, fmt, sep=',', cmt='#'):
    """loadimagepts.m: image measurement table; fmt lists the columns (id, im, x, y, sx, sy, sxy, ignored)."""
    parts, rows = _table(fName, fmt, sep, cmt)
    n = len(rows)
    col = lambda k: np.array([float(r[k]) for r in rows]) if k in parts else np.full(n, np.nan)
    sx, sy = (col('sxy'), col('sxy')) if 'sxy' in parts else (col('sx'), col('sy'))
    return NS(id=col('id').astype(np.int64), im=col('im').astype(np.int64), pos=np.vstack([col('x'), col('y')]),
              std=np.vstack([sx, sy]), fileName=fName)


def loadctrlpts(fName, fmt, sep=',', cmt='#'):
    """loadctrlpts.m: control point table; fmt lists the columns (id, label, x, y, z, sx, sy, sz, sxy, sxyz,
    ignored); a missing sigma is 0 (fixed)."""
    parts, rows = _table(fName, fmt, sep, cmt)
    ids = np.array([int(r['id']) for r in rows], dtype=np.int64) if 'id' in parts else np.full(len(rows), -1)
    pos = np.array([[float(r[k]) if k in r else np.nan for k in 'xyz'] for r in rows]).T.reshape(3, -1)
    std = np.zeros((3, len(rows)))
    for n, r in enumerate(rows):
        for k, v in r.items():
            if k in ('sx', 'sy', 'sz'):
                std['xyz'.index(k[1]), n] = float(v)
            elif k == 'sxy':
                std[0:2, n] = float(v)
            elif k == 'sxyz':
                std[:, n] = float(v)
    return NS(id=ids, name=[r.get('label', '') for r in rows], pos=pos, std=std, cov=None, fileName=fName)


def loadimagetable(fName, fmt, sep=',', cmt='#'):
    """loadimagetable.m: image list (id, cam, label, path, ignored); cam defaults to 1."""
    parts, rows = _table(fName, fmt, sep, cmt)
    return NS(id=np.array([int(r['id']) for r in rows], dtype=np.int64),
              cam=np.array([int(r['cam']) if 'cam' in r else 1 for r in rows]),
              name=[r.get('label', '') for r in rows], path=[r.get('path', '') for r in rows], fileName=fName)


def loadeotable(fName, fmt, sep=',', cmt='#'):
    """loadeotable.m: camera station table (id, label, x, y, z, omega, phi, kappa, sx, sy, sz, sxyz, so, sp,
    sk, sang, ignored); angles are returned in the file's unit."""
    parts, rows = _table(fName, fmt, sep, cmt)
    n = len(rows)
    num = lambda k: np.array([float(r[k]) for r in rows]) if k in parts else np.full(n, np.nan)
    std = np.vstack([num('sx'), num('sy'), num('sz')])
    if 'sxyz' in parts:
        std[:] = num('sxyz')
    angStd = np.vstack([num('so'), num('sp'), num('sk')])
    if 'sang' in parts:
        angStd[:] = num('sang')
    return NS(id=np.array([int(r['id']) for r in rows], dtype=np.int64) if 'id' in parts else np.full(n, -1),
              name=[r.get('label', '') for r in rows], pos=np.vstack([num('x'), num('y'), num('z')]),
              ang=np.vstack([num('omega'), num('phi'), num('kappa')]), std=std, angStd=angStd, fileName=fName)


def loadpm3dtbl(fName):
    """loadpm3dtbl.m (normal point table): PhotoModeler's exported 3-D point table - id, name, the photos
    a point is used in, position and precision (`fixed` precisions read as 0)."""
    import re
    ids, names, vis, pos, std = [], [], [], [], []
    with open(fName) as fh:
        for line in fh:
            m = re.match(r'^\s*(\d+)\s*,\s*"([^"]*)"\s*,\s*"([^"]*)"\s*,(.*)$', line.strip())
            if not m:
                continue
            nums = []
            for tok in m.group(4).split(','):
                tok = tok.strip()
                if tok == 'fixed':
                    nums.append(0.0)
                    continue
                try:
                    nums.append(float(tok))
                except ValueError:
                    break
            if len(nums) < 6:
                raise ValueError('Could not parse position part of string: %s' % m.group(4))
            ids.append(int(m.group(1)))
            names.append(m.group(2).strip())
            vis.append([int(v) for v in m.group(3).split(',') if v.strip()])
            pos.append(nums[0:3])
            std.append(nums[3:6])
    return NS(id=np.array(ids, dtype=np.int64), name=names, vis=vis, pos=np.array(pos).T.reshape(3, -1),
              std=np.array(std).T.reshape(3, -1), fileName=fName)


def loadpmreport(fName):
    """loadpmreport.m (the parts the external verification uses): PhotoModeler's status report - per photo
    the EO values and standard deviations [omega, phi, kappa (rad), X, Y, Z], and the processing summary
    (iterations, first / last total error)."""
    import re
    EO, STD, info = [], [], {}
    with open(fName) as fh:
        for line in fh:
            t = line.strip()
            m = re.match(r'(Number of Processing Iterations|First Error|Last Error): ([-\d.eE+]+)', t)
            if m:
                info[m.group(1)] = float(m.group(2))
            elif re.match(r'Photo \d+:', t):
                EO.append([])
                STD.append([])
            elif EO and len(STD[-1]) < 6:
                m = re.match(r'Value: ([-\d.eE+]+)', t)
                if m:
                    EO[-1].append(float(m.group(1)))
                m = re.match(r'Deviation: \w+: ([-\d.eE+]+)', t)
                if m:
                    STD[-1].append(float(m.group(1)))
    EO, STD = np.array([e[:6] for e in EO]).T, np.array(STD).T
    toRad = np.array([[np.pi / 180]] * 3 + [[1.0]] * 3)
    return NS(EO=(EO * toRad)[[3, 4, 5, 0, 1, 2]], EOstd=(STD * toRad)[[3, 4, 5, 0, 1, 2]],
              iterations=info.get('Number of Processing Iterations'), firstError=info.get('First Error'),
              lastError=info.get('Last Error'), fileName=fName)


def legacyloadeotable(fName, has=(True, True)):
    """legacyloadeotable.m: a camera-station table in the control-point format (`loadcpt`)."""
    return loadcpt(fName, has)


def matcheo(s, tbl, match='auto'):
    """matcheo.m: (i, j) - images i and the table rows j they match, by label, by id, or both."""
    byId = match in ('id', 'both') or (match == 'auto' and np.any(~np.isnan(np.asarray(tbl.id, float))))
    byLabel = match in ('label', 'both') or (match == 'auto' and any(tbl.name))
    il = jl = ii = ji = None
    if byLabel:
        _, il, jl = np.intersect1d(np.asarray(s.EO.label), np.asarray(tbl.name), return_indices=True)
    if byId:
        _, ii, ji = np.intersect1d(np.asarray(s.EO.id), np.asarray(tbl.id), return_indices=True)
    if byLabel and byId:
        if len(il):
            if not (np.array_equal(il, ii) and np.array_equal(jl, ji)):
                raise ValueError('Inconsistent match')
            return il, jl
        return ii, ji
    if byLabel:
        return il, jl
    if byId:
        return ii, ji
    raise ValueError('Neither match-by-id nor match-by-label possible.')


def setprioreo(s, tbl, i, j):
    """setprioreo.m: prior observations of camera positions; zero sigma fixes the coordinate."""
    s.proj.EOfile = tbl.fileName
    s.prior.EO.val[0:3, i] = tbl.pos[:, j]
    s.EO.val[0:3, i] = tbl.pos[:, j]
    s.prior.EO.std[0:3, i] = tbl.std[:, j]
    for a, b in zip(i, j):
        if tbl.name[b]:
            s.EO.label[a] = tbl.name[b]
    free = tbl.std[:, j] != 0
    s.prior.EO.use[0:3, i] = free
    s.bundle.est.EO[0:3, i] = free
    return s


def _io_rows(s, name, est=None):
    """Rows of IO.val addressed by a parameter name of setcamvals / setcamest."""
    nK, nP = s.IO.model.nK, s.IO.model.nP
    if name in IO_ROWS:
        return [IO_ROWS.index(name)]
    if name == 'lin' or name == 'af':
        return list(range(5))
    if name == 'pp':
        return [1, 2]
    if name == 'K':
        return list(range(5, 5 + nK))
    if name == 'P':
        return list(range(5 + nK, 5 + nK + nP))
    if name and name[0] in 'KP' and name[1:].isdigit():
        n = int(name[1:])
        if name[0] == 'K':
            if not 1 <= n <= nK:
                raise ValueError('K number out of range')
            if est is None:
                return [4 + n]
            return list(range(5, 5 + n)) if est else list(range(4 + n, 5 + nK))      # setcamest.m:84-88
        if not 1 <= n <= nP:
            raise ValueError('P number out of range')
        if est is None:
            return [4 + nK + n]
        if est:                                                                       # setcamest.m:95-101
            return list(range(5 + nK, 5 + nK + max(n, 2)))                            # P1 brings P2 along
        return list(range(4 + nK + (1 if n == 2 else n), 5 + nK + nP))                # and P2 takes P1 out
    raise ValueError("Bad parameter '%s'" % name)


def setcamvals(s, *args):
    """setcamvals.m: setcamvals(s[, 'prior'][, cams|'all'][, 'loaded'][, 'default', cc][, name, value ...])."""
    args = list(args)
    prior = bool(args) and isinstance(args[0], str) and args[0] == 'prior'
    if prior:
        args.pop(0)
    IO = s.prior.IO if prior else s.IO
    ix = slice(None)
    if args and not isinstance(args[0], str):
        ix = np.atleast_1d(args.pop(0))
    elif args and args[0] == 'all':
        args.pop(0)
    if args and args[0] == 'loaded':
        args.pop(0)
        IO.val[:, ix] = s.prior.IO.val[:, ix]
    if args and args[0] == 'default':
        args.pop(0)
        if not args or isinstance(args[0], str):
            raise ValueError('SETCAMVALS: CC must be numeric')
        ss = getattr(s.IO.sensor, 'ssSize', None)
        ss = s.IO.sensor.imSize * s.IO.sensor.pxSize if ss is None else ss
        IO.val[0, ix] = args.pop(0)
        IO.val[1, ix] = 0.5 * ss[0, ix]
        IO.val[2, ix] = -0.5 * ss[1, ix]
        IO.val[3:, ix] = 0
    if len(args) % 2:
        raise ValueError('SETCAMVALS: Trailing arguments must come in pairs')
    for name, val in zip(args[0::2], args[1::2]):
        rows = _io_rows(s, name)
        if set(rows) & {3, 4} and np.any(np.abs(s.IO.model.distModel[ix]) < 3) and np.any(np.asarray(val) != 0):
            raise ValueError('SETCAMVALS: Camera model only supports zero %s' % name)
        IO.val[np.ix_(rows, np.arange(IO.val.shape[1])[ix])] = np.reshape(val, (-1, 1)) if np.ndim(val) else val
    return s


def setcamest(s, *args):
    """setcamest.m: setcamest(s[, cams], names...) with 'not' switching the names after it off; 'all' and
    'af' leave aspect and skew off for models 1, 2 and -1, which have no affine terms."""
    args = list(args)
    cams = np.arange(s.IO.val.shape[1])
    if args and not isinstance(args[0], str):
        cams = np.atleast_1d(args.pop(0))
    affine = np.abs(s.IO.model.distModel[cams]) >= 3
    doEst = True
    for a in args:
        if a == 'not':
            if not doEst:
                raise ValueError('SETCAMEST: Cannot specify not twice')
            doEst = False
            continue
        if a in ('all', 'af'):
            rows = list(range(s.bundle.est.IO.shape[0])) if a == 'all' else list(range(5))
            val = np.full((len(rows), len(cams)), doEst)
            val[3:5] &= affine
            s.bundle.est.IO[np.ix_(rows, cams)] = val
            continue
        rows = _io_rows(s, a, doEst) if (a and a[0] in 'KP' and a[1:].isdigit()) else _io_rows(s, a)
        if doEst and set(rows) & {3, 4} and np.any(~affine) and a in ('as', 'sk'):
            raise ValueError('SETCAMEST: Model does not support skew, aspect.')
        s.bundle.est.IO[np.ix_(rows, cams)] = doEst
    return s


_EO_ROWS = {'x': [0], 'y': [1], 'z': [2], 'pos': [0, 1, 2], 'om': [3], 'ph': [4], 'ka': [5],
            'ang': [3, 4, 5], 'all': list(range(6))}


def seteoest(s, *args):
    """seteoest.m: seteoest(s[, cams], names...) ('not' as in setcamest, 'none' = nothing estimated), or
    seteoest(s, 'depend'[, camNo]) for the dependent-pair datum (camNo 1-based as in the reference)."""
    args = list(args)
    if args and isinstance(args[0], str) and args[0] == 'depend':          # setdepend, seteoest.m:90-128
        rest = args[1:]
        camNo = int(rest.pop(0)) - 1 if rest and not isinstance(rest[0], str) else 0
        if not rest or rest[0] == 'pos':
            return seteoest_depend(s, camNo)
        if rest[0] not in ('x', 'y', 'z'):
            raise ValueError('SETEOEST: Bad position argument for depend.')
        row = 'xyz'.index(rest[0])                                         # longest baseline along one axis
        offset = s.EO.val[row] - s.EO.val[row, camNo]
        s.bundle.est.EO[:] = True
        s.bundle.est.EO[:, camNo] = False
        s.bundle.est.EO[row, int(np.argmax(offset))] = False
        return s
    cams = np.arange(s.EO.val.shape[1])
    doEst = True
    for a in args:
        if not isinstance(a, str):
            cams = np.atleast_1d(a)
        elif a == 'not':
            if not doEst:
                raise ValueError('SETEOEST: Cannot specify not twice')
            doEst = False
        elif a == 'none':
            s.bundle.est.EO[np.ix_(range(6), cams)] = False
        elif a in _EO_ROWS:
            s.bundle.est.EO[np.ix_(_EO_ROWS[a], cams)] = doEst
        else:
            raise ValueError('SETEOEST: Bad argument string %s' % a)
    return s


def cleareo(s):
    """cleareo.m: NaN every EO element that is estimated and has no prior observation."""
    s.EO.val[s.bundle.est.EO & ~s.prior.EO.use] = np.nan
    return s


def clearop(s):
    """clearop.m: the same for the object points."""
    s.OP.val[s.bundle.est.OP & ~s.prior.OP.use] = np.nan
    return s


# ----------------------------------------------------------------------------- PhotoScan archives
def _ply_vertices(data):
    """Vertex table of a binary little-endian PLY file (what a .psz stores) as a dict of arrays."""
    head, _, body = data.partition(b'end_header\n')
    types = {'float': '<f4', 'double': '<f8', 'uint': '<u4', 'int': '<i4', 'uchar': 'u1', 'char': 'i1',
             'ushort': '<u2', 'short': '<i2'}
    n, fields, seen_vertex = 0, [], False
    for line in head.decode('ascii', 'replace').split('\n'):
        tok = line.split()
        if tok[:1] == ['format'] and tok[1] != 'binary_little_endian':
            raise ValueError('PLY format %s not supported' % tok[1])
        if tok[:1] == ['element']:
            seen_vertex = tok[1] == 'vertex'
            if seen_vertex:
                n = int(tok[2])
        elif tok[:1] == ['property'] and seen_vertex:
            fields.append((tok[2], types[tok[1]]))
    rec = np.frombuffer(body, dtype=np.dtype(fields), count=n)
    return {k: rec[k] for k, _ in fields}


def loadpsz(psFile):
    """loadpsz.m (the parts a bundle needs; document version 1.2, one chunk, one frame sensor): cameras
    with their chunk-local transforms, the chunk's local-to-global similarity (:203-226), markers with
    reference positions and image measurements, the sparse point cloud and its projections, the default
    accuracies (:1021-1056), the camera calibration in DBAT units (:700-868).  Positions are returned in
    the global frame: P_local = [I 0]/(T diag(1,-1,-1,1)) (:277-290), P_global = P_local/L2G (:341).
    DBAT point ids: markers 1..nCP in document order, cloud point i -> i+1+nCP (:451-470)."""
    import xml.etree.ElementTree as ET
    import zipfile
    z = zipfile.ZipFile(psFile)
    chunk = ET.fromstring(z.read('doc.xml')).find('chunks').find('chunk')
    tf = lambda txt: {'1': True, 'true': True, '0': False, 'false': False}[txt]
    vec = lambda node: np.array(node.text.split(), float)
    prop = lambda node: {p.get('name'): p.get('value') for p in node.findall('property')}
    defStd = {'tiePoints': np.nan, 'projections': np.nan, 'markers': np.nan, 'camPos': np.nan}
    st = prop(chunk.find('settings')) if chunk.find('settings') is not None else {}
    for key, name in (('tiepoints', 'tiePoints'), ('cameras', 'camPos'), ('markers', 'markers'), ('projections', 'projections')):
        if 'accuracy_' + key in st:
            defStd[name] = float(st['accuracy_' + key])
    R, T, S = np.eye(4), np.eye(4), np.eye(4)
    xf = chunk.find('transform')
    if xf is not None:
        if xf.find('rotation') is not None:
            R[:3, :3] = vec(xf.find('rotation')).reshape(3, 3)
        if xf.find('translation') is not None:
            T[:3, 3] = vec(xf.find('translation'))
        if xf.find('scale') is not None:
            S[:3, :3] *= float(xf.find('scale').text)
    L2G = T @ S @ R
    cams = [c for c in chunk.find('cameras').findall('camera')]
    keep = [c for c in cams if tf(c.get('enabled')) and c.find('transform') is not None]
    if len({c.get('sensor_id') for c in keep}) > 1:
        raise NotImplementedError('Handling of cameras for multiple sensor ids not implemented')
    camIds = [int(c.get('id')) for c in keep]
    cam_of = {v: i for i, v in enumerate(camIds)}
    CC = np.zeros((3, len(keep)))
    Rg = np.zeros((3, 3, len(keep)))
    for i, c in enumerate(keep):
        Tc = vec(c.find('transform')).reshape(4, 4)
        Pl = np.eye(3, 4) @ np.linalg.inv(Tc @ np.diag([1.0, -1.0, -1.0, 1.0]))
        Pg = Pl @ np.linalg.inv(L2G)
        CC[:, i] = -np.linalg.solve(Pg[:, :3], Pg[:, 3])           # null(P), euclidean
        Rg[:, :, i] = Pg[:, :3] / np.cbrt(np.linalg.det(Pg[:, :3]))
    # markers (control points)
    markers = chunk.find('markers').findall('marker') if chunk.find('markers') is not None else []
    rawCP = [int(m.get('id')) for m in markers]
    cp_of = {v: i + 1 for i, v in enumerate(rawCP)}
    cpPos = np.full((3, len(markers)), np.nan)
    cpStd = np.full((3, len(markers)), np.nan)
    cpEnabled = np.zeros(len(markers), bool)
    for i, m in enumerate(markers):
        ref = m.find('reference')
        cpStd[:, i] = defStd['markers']
        if ref is not None:
            for k, axis in enumerate('xyz'):
                if ref.get(axis) is not None:
                    cpPos[k, i] = float(ref.get(axis))
            if ref.get('sxyz') is not None:
                cpStd[:, i] = float(ref.get('sxyz'))
            elif ref.get('sxy') is not None:
                cpStd[0:2, i] = float(ref.get('sxy'))
            for k, axis in enumerate(('sx', 'sy', 'sz')):
                if ref.get(axis) is not None:
                    cpStd[k, i] = float(ref.get(axis))
            if ref.get('enabled') is not None:
                cpEnabled[i] = tf(ref.get('enabled'))
    nCP = len(markers)
    frame = chunk.find('frames').find('frame')
    # sparse cloud and its projections
    pc = frame.find('point_cloud')
    objPts = np.zeros((0, 4))
    objMark = np.zeros((0, 4))
    if pc is not None:
        pts = _ply_vertices(z.read(pc.find('points').get('path')))
        local = np.vstack([pts['x'], pts['y'], pts['z'], np.ones(len(pts['x']))]).astype(float)
        glob = L2G @ local
        objPts = np.column_stack([pts['id'].astype(float) + 1 + nCP, (glob[:3] / glob[3]).T])
        rows = []
        for pr in pc.findall('projections'):
            if int(pr.get('camera_id')) not in cam_of:
                continue
            v = _ply_vertices(z.read(pr.get('path')))
            rows.append(np.column_stack([np.full(len(v['id']), cam_of[int(pr.get('camera_id'))] + 1.0),
                                         v['id'].astype(float) + 1 + nCP, v['x'].astype(float), v['y'].astype(float)]))
        if rows:
            objMark = np.vstack(rows)
            objMark = objMark[np.lexsort((objMark[:, 1], objMark[:, 0]))]
    rows = []
    fm = frame.find('markers')
    for m in (fm.findall('marker') if fm is not None else []):
        for loc in m.findall('location'):
            if int(loc.get('camera_id')) in cam_of:
                rows.append([cam_of[int(loc.get('camera_id'))] + 1.0, cp_of[int(m.get('marker_id'))],
                             float(loc.get('x')), float(loc.get('y'))])
    ctrlMark = np.array(rows).reshape(-1, 4)
    ctrlMark = ctrlMark[np.lexsort((ctrlMark[:, 1], ctrlMark[:, 0]))]
    # image names
    psDir = os.path.dirname(os.path.abspath(psFile))
    imNames = [''] * len(keep)
    for c in frame.find('cameras').findall('camera'):
        if int(c.get('camera_id')) in cam_of:
            p = c.find('photo').get('path')
            absolute = (not p) or p[0] in '/\\' or (len(p) > 1 and p[1] == ':')
            imNames[cam_of[int(c.get('camera_id'))]] = p if absolute else os.path.normpath(os.path.join(psDir, p))
    # camera (:700-868)
    sid = keep[0].get('sensor_id')
    sensor = [x for x in chunk.find('sensors').findall('sensor') if x.get('id') == sid][0]
    cals = sensor.findall('calibration')
    cal = ([c for c in cals if c.get('class') == 'adjusted'] or cals)[0]
    num = lambda tag, d=None: float(cal.find(tag).text) if cal.find(tag) is not None else d
    imSz = np.array([float(sensor.find('resolution').get('width')), float(sensor.find('resolution').get('height'))])
    ppIsAbsolute = cal.find('fx') is not None or cal.find('fy') is not None
    fx, fy = num('fx', np.nan), num('fy', np.nan)
    if cal.find('f') is not None:
        ppIsAbsolute = False
        fy = num('f')
        fx = fy + num('b1', 0.0)
    cx, cy = num('cx', 0.0), num('cy', 0.0)
    if not ppIsAbsolute:
        cx, cy = cx + imSz[0] / 2, cy + imSz[1] / 2
    k = [num('k%d' % n) for n in range(1, 5) if cal.find('k%d' % n) is not None]
    p = [num('p%d' % n) for n in range(1, 5) if cal.find('p%d' % n) is not None]
    sp_ = prop(sensor)
    pixelSz = np.array([float(sp_.get('pixel_width', 1)), float(sp_.get('pixel_height', 1))])
    focal = fx * pixelSz[0]
    kk = np.array([-v * focal ** (-2 * (n + 1)) for n, v in enumerate(k)])
    pp_ = np.array(p, float)
    if len(pp_) >= 2:
        pp_[0:2] = pp_[0:2] / focal
        pp_[1] = -pp_[1]
        pp_[0:2] = pp_[[1, 0]]
    camera = NS(name=sensor.get('label'), imSz=imSz, pixelSz=pixelSz, sensorFormat=imSz * pixelSz, focal=focal,
                pp=np.array([cx, cy]) * pixelSz, k=kk, p=pp_, isFixed=tf(sp_.get('fixed', 'true')),
                isAdjusted=cal.get('class') == 'adjusted')
    return NS(fileName=psFile, camera=camera, cameraIds=np.array(camIds), cameraLabels=[c.get('label') for c in keep],
              imNames=imNames, defStd=NS(**defStd), L2G=L2G,
              glob=NS(CC=CC, R=Rg, objPts=objPts,
                      controlPts=NS(id=np.arange(1, nCP + 1), rawId=np.array(rawCP), pos=cpPos, std=cpStd,
                                    enabled=cpEnabled, labels=[m.get('label', '') for m in markers])),
              markPts=NS(ctrl=ctrlMark, obj=objMark))


def ps2pmstruct(psz):
    """ps2pmstruct.m (global frame): a loaded PhotoScan project in the shape `loadpm` returns, so that
    `prob2dbatstruct` applies - camera in PhotoModeler conventions, EO angles by `derotmat3d` of the global
    rotation (:36-41), enabled markers as control points and the others as check points (:51-66), marker
    measurements at the 'projections' accuracy and tie points at the 'tiePoints' accuracy (:86-91)."""
    cam = psz.camera
    k1k3, p1p2 = np.zeros(3), np.zeros(2)
    k1k3[:min(3, len(cam.k))] = cam.k[:3]
    p1p2[:min(2, len(cam.p))] = cam.p[:2]
    defCam = np.concatenate([[cam.focal], cam.pp, cam.sensorFormat, k1k3, p1p2])
    job = NS(fileName=psz.fileName, title='Photoscan import', defCam=defCam, defCamStd=np.zeros(10), imSz=cam.imSz)
    g = psz.glob
    images = []
    for i in range(g.CC.shape[1]):
        M = g.R[:, :, i]
        ang = np.array([np.arctan2(-M[2, 1], M[2, 2]), np.arcsin(M[2, 0]), np.arctan2(-M[1, 0], M[0, 0])])   # derotmat3d.m
        images.append(NS(imName=psz.imNames[i], outer=np.concatenate([g.CC[:, i], np.rad2deg(ang[[2, 1, 0]])]),
                         outerStd=np.zeros(6), imSz=cam.imSz, id=int(psz.cameraIds[i]), label=psz.cameraLabels[i]))
    cp = g.controlPts
    tab = lambda sel: np.column_stack([cp.id[sel], cp.pos[:, sel].T, cp.std[:, sel].T]).reshape(-1, 7)
    ctrlPts, checkPts = tab(cp.enabled), tab(~cp.enabled)
    ok = np.array([np.count_nonzero(psz.markPts.ctrl[:, 1] == i) >= 2 for i in checkPts[:, 0]], bool)
    checkPts = checkPts[ok]
    cPts = np.vstack([ctrlPts, checkPts])
    order = np.argsort(cPts[:, 0], kind='stable')
    cPts = cPts[order]
    objPts = np.vstack([cPts, np.column_stack([g.objPts, np.full((len(g.objPts), 3), np.nan)])])
    rawOPids = np.concatenate([cp.rawId[cPts[:, 0].astype(int) - 1], g.objPts[:, 0] - 1 - len(cp.id)])
    OPlabels = [cp.labels[int(i) - 1] for i in cPts[:, 0]] + [''] * len(g.objPts)
    markPts = np.vstack([np.column_stack([psz.markPts.ctrl, np.full((len(psz.markPts.ctrl), 2), psz.defStd.projections)]),
                         np.column_stack([psz.markPts.obj, np.full((len(psz.markPts.obj), 2), psz.defStd.tiePoints)])])
    markPts = markPts[np.lexsort((markPts[:, 1], markPts[:, 0]))]
    markPts = markPts[np.isin(markPts[:, 1], objPts[:, 0])]
    markPts[:, 0] -= 1
    return NS(job=job, images=images, ctrlPts=ctrlPts, cptFile='', checkPts=checkPts, objPts=objPts,
              priorCamPos=np.zeros((0, 7)), rawOPids=rawOPids, OPlabels=OPlabels, markPts=markPts)


def filterprob(prob, s0, minRays=0, minAngle=0.0):
    """loadplotpsz.m:58-90: drop the object points (never control points) seen in fewer than minRays images
    or whose largest ray intersection angle is below minAngle degrees, from prob; returns (prob, removed ids).
    Rebuild the struct with prob2dbatstruct afterwards, as the reference does."""
    from .report import angles
    isCtrl = np.asarray(s0.prior.OP.isCtrl, bool)
    bad = np.zeros(len(isCtrl), bool)
    if minRays > 0:
        bad |= (np.bincount(s0.IP.op, minlength=len(isCtrl)) < minRays) & ~isCtrl
    if minAngle > 0:
        with np.errstate(invalid='ignore'):
            bad |= (np.rad2deg(angles(s0)) < minAngle) & ~isCtrl
    ids = np.asarray(s0.OP.id)[bad]
    keepObj = ~np.isin(prob.objPts[:, 0], ids)
    out = NS(**vars(prob))
    out.objPts = prob.objPts[keepObj]
    out.rawOPids = np.asarray(prob.rawOPids)[keepObj]
    out.OPlabels = [l for l, k in zip(prob.OPlabels, keepObj) if k]
    out.markPts = prob.markPts[~np.isin(prob.markPts[:, 1], ids)]
    return out, ids
