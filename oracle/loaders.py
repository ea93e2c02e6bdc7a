"""Fixture loaders for the reference's script-format projects (test infrastructure).

Restates the subset of `code/script/{parseinput,setdbatpts,setdbatcamsandimages,
parsedbatxmlcamstruct}.m` and `code/classes/@DBATCamera/DBATCamera.m:56-134` that the
camcal XML project (`data/script/camcaldemo/`) needs: comma-separated tables with a
`format="..."` header and one XML camera.  Conventions:
internal py = -user py, internal K,P = -user K,P, internal `as` = 1 - user aspect,
square pixel size = sensor height / image height (`sensor="auto,h"`).
"""
import re
from types import SimpleNamespace as NS

import numpy as np

from .dbatstruct import new_struct


def load_table(path):
    rows = []
    with open(path) as fh:
        for line in fh:
            line = line.strip()
            if not line or line.startswith('#'):
                continue
            rows.append([c.strip() for c in line.split(',')])
    return rows


def _tag(xml, name):
    m = re.search(r'<%s>\s*([^<]*?)\s*</%s>' % (name, name), xml)
    return m.group(1) if m else None


def parse_camera_xml(text):
    """One <camera> element → dict in USER conventions (DBATCamera.m)."""
    cam = text[text.index('<camera>'):text.index('</camera>')]
    out = {}
    for k in ('sensor', 'image', 'aspect', 'focal', 'model', 'nK', 'nP', 'cc', 'pp', 'skew',
              'K', 'P'):
        out[k] = _tag(cam, k)
    return out


def camera_internal(cam, calibrated):
    """User camera dict → (IO column, pxSize[2], imSize[2], distModel, nK, nP).

    DBATCamera.m:56-134 / parsedbatxmlcamstruct.m:92-119.
    """
    imSize = np.array([float(v) for v in cam['image'].split(',')])
    sens = cam['sensor'].split(',')
    h = float(sens[1])
    px = h / imSize[1]                                   # square pixels from the height
    w = px * imSize[0] if sens[0].strip() == 'auto' else float(sens[0])
    nK, nP = int(cam['nK']), int(cam['nP'])
    io = np.zeros(5 + nK + nP)
    if calibrated:
        io[0] = float(cam['cc'])
        pp = [float(v) for v in cam['pp'].split(',')]
        io[1], io[2] = pp[0], -pp[1]
        io[3] = 1 - float(cam['aspect'])
        io[4] = float(cam['skew'])
        io[5:5 + nK] = [-float(v) for v in cam['K'].split(',')]
        io[5 + nK:] = [-float(v) for v in cam['P'].split(',')]
    else:                                                # 'default' initial values (setcamvals.m:66-76)
        io[0] = float(cam['focal'])
        io[1], io[2] = px * imSize[0] / 2, -h / 2
    return io, np.array([px, px]), imSize, int(cam['model']), nK, nP


def load_camcal_script(root, calibrated_cam_xml=None):
    """Build the DBAT struct of `data/script/camcaldemo` (or a copy under tests/golden).

    Control points are fixed (std 0 ⇒ not estimated, not observed: setcpt.m); all other OP,
    all EO, and IO except skew are estimated (camcaldemo.xml <set_bundle_estimate_params>).
    EO/OP values are left NaN — callers supply start values.
    """
    import os
    xml = open(os.path.join(root, 'camcaldemo.xml')).read()
    cam = parse_camera_xml(xml)
    if calibrated_cam_xml is not None:
        cam.update({k: v for k, v in parse_camera_xml(calibrated_cam_xml).items() if v})
    io, pxSize, imSize, model, nK, nP = camera_internal(cam, calibrated_cam_xml is not None)
    ctrl = load_table(os.path.join(root, 'reference', 'camcal-fixed.txt'))
    imgs = load_table(os.path.join(root, 'images', 'images.txt'))
    pts = load_table(os.path.join(root, 'measurements', 'markpts.txt'))
    img_ids = [int(r[0]) for r in imgs]
    im_of = {v: i for i, v in enumerate(img_ids)}
    op_ids = sorted({int(r[1]) for r in pts} | {int(r[0]) for r in ctrl})
    op_of = {v: i for i, v in enumerate(op_ids)}
    nImg, nOP = len(img_ids), len(op_ids)
    ip_img = np.array([im_of[int(r[0])] for r in pts])
    ip_op = np.array([op_of[int(r[1])] for r in pts])
    IPval = np.array([[float(r[2]), float(r[3])] for r in pts]).T
    IPstd = np.array([[float(r[4]), float(r[4])] for r in pts]).T
    OP = np.full((3, nOP), np.nan)
    s = new_struct(np.tile(io[:, None], (1, nImg)), np.full((6, nImg), np.nan), OP, IPval,
                   ip_img, ip_op, pxSize[:, None], imSize[:, None], model, nK, nP, IPstd)
    s.OP.id = np.array(op_ids)
    s.EO.id = np.array(img_ids)
    s.EO.name = [r[1].replace('\\', '/').split('/')[-1] for r in imgs]
    s.bundle.est.IO[:] = True
    s.bundle.est.IO[4, :] = False                        # <skew>false</skew>
    s.bundle.est.EO[:] = True
    s.bundle.est.OP[:] = True
    s.OP.label = [''] * nOP
    for r in ctrl:                                       # fixed control points (setcpt.m:24-51)
        j = op_of[int(r[0])]
        s.OP.val[:, j] = [float(r[2]), float(r[3]), float(r[4])]
        s.OP.label[j] = r[1]
        s.prior.OP.val[:, j] = s.OP.val[:, j]
        s.prior.OP.std[:, j] = 0.0
        s.bundle.est.OP[:, j] = False
    s.prior.OP.isCtrl = ~s.bundle.est.OP.all(axis=0)
    s.prior.OP.isCheck = np.zeros(nOP, bool)
    s.IP.sigmas = np.unique(IPstd)
    s.IO.model.camUnit = 'mm'
    m = re.search(r'<name>\s*([^<]*?)\s*</name>', xml)
    s.proj = NS(objUnit='m', x0desc='', title=m.group(1) if m else '', UUID='', EOfile='',
                fileName=os.path.join(root, 'camcaldemo.xml'),
                cptFile=os.path.join(root, 'reference', 'camcal-fixed.txt'))
    return s


def load_sxb_script(root):
    """The DBAT struct of `data/script/sxb/sxb.xml` after its input section and the operations up to
    `set_bundle_estimate_params`: one calibrated aerial camera (IO loaded, not estimated), control points
    with prior standard deviations (weighted: estimated and observed, `setcpt.m:40-51`), ids 351 and 410
    filtered out of the control set and kept as check points (prior kept for the report, not observed),
    two image-point files with their own sigmas (`sxy` 0.5 / 1.0 px), EO estimated.  OP start values
    'loaded': control and check points at their prior position, the rest NaN until forward intersection."""
    import os
    xml = open(os.path.join(root, 'sxb.xml')).read()
    cam = parse_camera_xml(xml)
    cam['nK'], cam['nP'] = len(cam['K'].split(',')), len(cam['P'].split(','))
    io, pxSize, imSize, model, nK, nP = camera_internal(cam, True)
    ref = load_table(os.path.join(root, 'reference', 'sxb-control.txt'))
    check_ids = {351, 410}                                   # <filter id="351,410">
    imgs = load_table(os.path.join(root, 'images', 'images.txt'))
    pts = []
    for fname, sxy in (('markpts.txt', 0.5), ('smartpts.txt', 1.0)):
        pts += [(int(r[1]), int(r[0]), float(r[2]), float(r[3]), sxy)
                for r in load_table(os.path.join(root, 'measurements', fname))]
    img_ids = [int(r[0]) for r in imgs]
    im_of = {v: i for i, v in enumerate(img_ids)}
    op_ids = sorted({p[1] for p in pts} | {int(r[0]) for r in ref})
    op_of = {v: i for i, v in enumerate(op_ids)}
    nImg, nOP = len(img_ids), len(op_ids)
    s = new_struct(np.tile(io[:, None], (1, nImg)), np.full((6, nImg), np.nan), np.full((3, nOP), np.nan),
                   np.array([[p[2], p[3]] for p in pts]).T, np.array([im_of[p[0]] for p in pts]),
                   np.array([op_of[p[1]] for p in pts]), pxSize[:, None], imSize[:, None], model, nK, nP,
                   np.array([[p[4], p[4]] for p in pts]).T)
    s.OP.id = np.array(op_ids)
    s.OP.label = [''] * nOP
    s.EO.id = np.array(img_ids)
    s.EO.name = [r[1].replace('\\', '/').split('/')[-1] for r in imgs]
    s.bundle.est.IO[:] = False
    s.bundle.est.EO[:] = True
    s.bundle.est.OP[:] = True
    s.prior.OP.isCtrl = np.zeros(nOP, bool)
    s.prior.OP.isCheck = np.zeros(nOP, bool)
    for r in ref:
        j = op_of[int(r[0])]
        s.OP.label[j] = r[1]
        s.prior.OP.val[:, j] = [float(v) for v in r[2:5]]
        s.prior.OP.std[:, j] = [float(v) for v in r[5:8]]
        s.OP.val[:, j] = s.prior.OP.val[:, j]
        if int(r[0]) in check_ids:
            s.prior.OP.isCheck[j] = True
        else:
            s.prior.OP.isCtrl[j] = True
            s.prior.OP.use[:, j] = True
    s.IP.sigmas = np.unique(s.IP.std)
    s.IO.model.camUnit = 'mm'
    m = re.search(r'<name>\s*([^<]*?)\s*</name>', xml)
    s.proj = NS(objUnit='m', x0desc='', title=m.group(1) if m else '', UUID='', EOfile='',
                fileName=os.path.join(root, 'sxb.xml'), cptFile=os.path.join(root, 'reference', 'sxb-control.txt'))
    return s


def load_roma_script(root):
    """The DBAT struct of `data/script/romabundledemo/romabundledemo.xml` up to (not including) its
    forward_intersection: 60 images of one calibrated camera (IO loaded; cc, px, py, K1, K2 estimated),
    EO start values from `prior/initial_eo.txt` (degrees), 26321 object points without start values,
    90561 image points at sigma 1 px, no control points.  The datum (`set_datum depend`, ref_cam 1) is
    set by the caller after the intersection, as the script does.  markpts.txt is stored xz-compressed."""
    import lzma
    import os
    xml = open(os.path.join(root, 'romabundledemo.xml')).read()
    cam = parse_camera_xml(open(os.path.join(root, 'cameras', 'EOS5DMarkII.xml')).read())
    io, pxSize, imSize, model, nK, nP = camera_internal(cam, True)
    imgs = load_table(os.path.join(root, 'images', 'images.txt'))
    eo = load_table(os.path.join(root, 'prior', 'initial_eo.txt'))
    with lzma.open(os.path.join(root, 'measurements', 'markpts.txt.xz'), 'rt') as fh:
        mk = np.loadtxt(fh, delimiter=',', comments='#')
    img_ids = [int(r[0]) for r in imgs]
    im_of = {v: i for i, v in enumerate(img_ids)}
    op_ids, ip_op = np.unique(mk[:, 1].astype(np.int64), return_inverse=True)
    nImg, nOP = len(img_ids), len(op_ids)
    EO = np.full((6, nImg), np.nan)
    for r in eo:
        i = im_of[int(r[0])]
        EO[0:3, i] = [float(v) for v in r[1:4]]
        EO[3:6, i] = np.deg2rad([float(v) for v in r[4:7]])
    s = new_struct(np.tile(io[:, None], (1, nImg)), EO, np.full((3, nOP), np.nan), mk[:, 2:4].T,
                   np.array([im_of[int(v)] for v in mk[:, 0]]), ip_op, pxSize[:, None], imSize[:, None],
                   model, nK, nP, 1.0)
    s.OP.id = op_ids
    s.OP.label = [''] * nOP
    s.EO.id = np.array(img_ids)
    s.EO.name = [r[1].replace('\\', '/').split('/')[-1] for r in imgs]
    s.bundle.est.IO[:] = False
    s.bundle.est.IO[[0, 1, 2, 5, 6], :] = True               # all but aspect, skew, P, K3
    s.bundle.est.EO[:] = True
    s.bundle.est.OP[:] = True
    s.prior.OP.isCtrl = np.zeros(nOP, bool)
    s.prior.OP.isCheck = np.zeros(nOP, bool)
    s.IP.sigmas = np.array([1.0])
    s.IO.model.camUnit = 'mm'
    m = re.search(r'<name>\s*([^<]*?)\s*</name>', xml)
    s.proj = NS(objUnit='m', x0desc='', title=m.group(1) if m else '', UUID='', cptFile='',
                fileName=os.path.join(root, 'romabundledemo.xml'),
                EOfile=os.path.join(root, 'prior', 'initial_eo.txt'))
    return s


def load_camera_stations(path):
    """result/camera_stations.txt → (ids, EO 6xN [radians], std 6xN as printed)."""
    rows = load_table(path)
    ids = np.array([int(r[1]) for r in rows])
    vals = np.array([[float(v) for v in r[2:14]] for r in rows]).T
    return ids, vals[0:6], vals[6:12]


# --------------------------------------------------------------------------- PhotoModeler text export
def load_pm_export(path, imSz=None):
    """Subset of `code/file/loadpm.m:104-330`: 5 header lines, per-image 6-line blocks
    (id name; outer x y z kappa phi omega [deg]; std; cov (blank); inner f xp yp xs ys K1 K2 K3 P1 P2;
    std), control points `id x y z sx sy sz`, object points (same columns), mark points
    `photo(0-based) id x y sx sy`; sections end with a blank line."""
    lines = open(path).read().split('\n')
    hdr2 = lines[1].split()
    # line 2 carries the image size only in newer exports; loadpm.m:8-9,117-120 then takes it from
    # the caller (or reads the image files, which are not shipped)
    if imSz is None:
        imSz = [float(hdr2[2]), float(hdr2[3])]
    job = {'imSz': np.array(imSz, float), 'title': lines[0].rstrip('\r'), 'fileName': path,
           'defCam': np.array([float(v) for v in lines[3].split()])}
    k = 5
    images = []
    while k < len(lines):
        tok = lines[k].split()
        if len(tok) < 2 or tok[1].replace('.', '').replace('-', '').isdigit():
            break
        outer = np.array([float(v) for v in lines[k + 1].split()[1:]])
        inner = np.array([float(v) for v in lines[k + 4].split()[1:]])
        images.append({'id': int(tok[0]), 'name': tok[1], 'outer': outer, 'inner': inner})
        k += 6
    def table(k, ncol):
        """Rows up to the next blank line (loadpm.m:230-333); an empty section is just its blank line."""
        rows = []
        while k < len(lines) and lines[k].strip():
            rows.append([float(v) for v in lines[k].split()])
            k += 1
        return (np.array(rows) if rows else np.zeros((0, ncol))), k + 1
    k += 1                                                 # blank line that ends the image blocks
    ctrl, k = table(k, 7)
    obj, k = table(k, 7)
    mark, k = table(k, 6)
    # normal and "smart" points numbered from overlapping ranges: shift the smart ids above the
    # normal ones (loadpm.m:385-404; smart mark points are the ones exported with zero std)
    if len(mark) and len(obj):
        smart = np.all(mark[:, 4:6] == 0, axis=1)
        normId, smartId = np.unique(mark[~smart, 1]), np.unique(mark[smart, 1])
        split = np.flatnonzero(np.diff(obj[:, 0]) < 0)
        if len(split) and len(normId) and len(smartId):
            shift = normId.max() + 1 - smartId.min()
            mark[smart, 1] += shift
            isSmartObj = np.isin(obj[:, 0], smartId)
            isSmartObj[:split[0] + 1] = False
            obj[isSmartObj, 0] += shift
    return {'job': job, 'images': images, 'ctrlPts': ctrl, 'objPts': obj, 'markPts': mark}


def prague_cam_struct(root, stub, cpfile=None, shift_cp=True, orient='no'):
    """The prague2016 'cam' projects C1 ('fixed') / C2 ('weighted') as `prague2016_pm.m:100-215` sets
    them up: `prob2dbatstruct` (`misc/prob2dbatstruct.m:198-420`: model 1, square pixels from the sensor
    height, py/K/P sign flips, angles = outer([6,5,4]) deg), fixed camera (`setcamest(s,'not','all')`),
    control points from `ref/ctrlpts-<stub>.txt` shifted by the mean offset to PhotoModeler's frame
    (`:166-190`) and set with `setcpt` (std 0 => fixed, otherwise prior observation).  EO and OP start
    values are PhotoModeler's own estimates from the export (the demo recomputes them by resection /
    intersection; the converged result does not depend on that)."""
    import os
    prob = load_pm_export(os.path.join(root, 'pmexports', '%s-%s-orient-pmexport.txt' % (stub, orient)))
    nImg = len(prob['images'])
    ids = np.unique(np.concatenate([prob['ctrlPts'][:, 0], prob['objPts'][:, 0]])).astype(int)
    op_of = {v: i for i, v in enumerate(ids)}
    nK, nP = 3, 2
    inner = prob['job']['defCam']
    imSz = prob['job']['imSz']
    IO = np.zeros((5 + nK + nP, nImg))
    IO[0] = inner[0]
    IO[1], IO[2] = inner[1], -inner[2]
    IO[5:8] = -inner[5:8, None]
    IO[8:10] = -inner[8:10, None]
    px = inner[3:5] / imSz
    IO[3] = 1 - px[0] / px[1]
    pxSize = np.array([[px[1]], [px[1]]])
    EO = np.zeros((6, nImg))
    for i, im in enumerate(prob['images']):
        EO[0:3, i] = im['outer'][0:3]
        EO[3:6, i] = np.deg2rad(im['outer'][[5, 4, 3]])
    OP = np.full((3, len(ids)), np.nan)
    for r in np.vstack([prob['ctrlPts'], prob['objPts']]):
        OP[:, op_of[int(r[0])]] = r[1:4]
    mk = prob['markPts']
    keep = np.array([int(r[1]) in op_of for r in mk])
    mk = mk[keep]
    mstd = mk[:, 4:6].copy()
    if np.any(mstd == 0):                                  # prob2dbatstruct.m:367-373: any zero => all 1 px
        mstd[:] = 1.0
    s = new_struct(IO, EO, OP, mk[:, 2:4].T, mk[:, 0].astype(int), np.array([op_of[int(v)] for v in mk[:, 1]]),
                   pxSize, imSz[:, None], 1, nK, nP, mstd.T)
    s.OP.id = ids
    s.bundle.est.IO[:] = False
    s.bundle.est.EO[:] = True
    s.bundle.est.OP[:] = True
    cp = load_table(os.path.join(root, 'ref', cpfile or 'ctrlpts-%s.txt' % stub))
    cp_id = [int(r[0]) for r in cp]
    cp_pos = np.array([[float(v) for v in r[2:5]] for r in cp]).T
    cp_std = np.array([[float(v) for v in r[5:8]] if len(r) >= 8 else [0.0, 0.0, 0.0] for r in cp]).T
    pm_pos = np.array([prob['ctrlPts'][prob['ctrlPts'][:, 0] == i, 1:4][0] for i in cp_id]).T
    if shift_cp:                                                           # sxb_prior_eo.m sets them unshifted
        cp_pos = cp_pos + np.mean(pm_pos - cp_pos, axis=1, keepdims=True)  # prague2016_pm.m:174-190
    s.prior.OP.isCtrl = np.zeros(len(ids), bool)
    s.prior.OP.isCheck = np.zeros(len(ids), bool)
    s.OP.label = [''] * len(ids)
    s.IP.sigmas = np.unique(mstd)                                          # prob2dbatstruct.m:367
    s.EO.name = [im['name'].replace('\\', '/').split('/')[-1] for im in prob['images']]
    s.IO.model.camUnit = 'mm'
    s.proj = NS(objUnit='m', x0desc='', title=prob['job']['title'], UUID='', EOfile='',
                fileName=prob['job']['fileName'], cptFile=os.path.join(root, 'ref', cpfile or 'ctrlpts-%s.txt' % stub))
    for k, i in enumerate(cp_id):                                          # setcpt.m
        j = op_of[i]
        s.OP.label[j] = cp[k][1]
        s.prior.OP.val[:, j] = cp_pos[:, k]
        s.OP.val[:, j] = cp_pos[:, k]
        s.prior.OP.std[:, j] = cp_std[:, k]
        fixed = np.all(cp_std[:, k] == 0)
        s.prior.OP.use[:, j] = not fixed
        s.bundle.est.OP[:, j] = not fixed
        s.prior.OP.isCtrl[j] = True
    return s


def stpierre_struct(root, imSz=(6912, 5212)):
    """StPierre `C5_reduced` PhotoModeler export (`data/hamburg2017/stpierre/pmexports`) set up like
    `stpierrebundledemo_ps.m:100-111`: forward Brown model (-1), self-calibration of all camera
    parameters except skew and aspect (`setcamest(s0,'all','not','sk','as')`), EO and OP start values
    as loaded, datum from the weighted control points (prior observations, `setcpt.m`).  The export's
    header has no image size and the images are not shipped; the size used here makes the pixels
    square for the exported sensor (9.9658 x 7.5152 mm) and covers the largest measured coordinate
    (6885.6, 5206.6) - there is no reference golden for this data set, it is an oracle-vs-device case."""
    import os
    prob = load_pm_export(os.path.join(root, 'C5_reduced-pmexport.txt'), imSz)
    nImg = len(prob['images'])
    ids = np.unique(np.concatenate([prob['ctrlPts'][:, 0], prob['objPts'][:, 0]])).astype(int)
    op_of = {v: i for i, v in enumerate(ids)}
    nK, nP = 3, 2
    inner = prob['job']['defCam']
    imSz = prob['job']['imSz']
    IO = np.zeros((5 + nK + nP, nImg))
    IO[0] = inner[0]
    IO[1], IO[2] = inner[1], -inner[2]                     # prob2dbatstruct.m:226-236 sign conventions
    IO[5:8] = -inner[5:8, None]
    IO[8:10] = -inner[8:10, None]
    px = inner[3:5] / imSz
    IO[3] = 1 - px[0] / px[1]
    pxSize = np.array([[px[1]], [px[1]]])
    EO = np.zeros((6, nImg))
    for i, im in enumerate(prob['images']):
        EO[0:3, i] = im['outer'][0:3]
        EO[3:6, i] = np.deg2rad(im['outer'][[5, 4, 3]])
    OP = np.full((3, len(ids)), np.nan)
    for r in np.vstack([prob['objPts'], prob['ctrlPts']]):
        OP[:, op_of[int(r[0])]] = r[1:4]
    mk = prob['markPts']
    keep = np.array([int(r[1]) in op_of for r in mk])
    mk = mk[keep]
    order = np.lexsort((mk[:, 1], mk[:, 0]))               # columns sorted by (image, id), prob2dbatstruct.m:349-365
    mk = mk[order]
    mstd = mk[:, 4:6].copy()
    if np.any(mstd == 0):                                  # prob2dbatstruct.m:367-373: any zero => all 1 px
        mstd[:] = 1.0
    s = new_struct(IO, EO, OP, mk[:, 2:4].T, mk[:, 0].astype(int), np.array([op_of[int(v)] for v in mk[:, 1]]),
                   pxSize, imSz[:, None], -1, nK, nP, mstd.T)
    s.OP.id = ids
    s.bundle.est.IO[:] = True
    s.bundle.est.IO[3:5, :] = False                        # not 'as', 'sk'
    s.bundle.est.EO[:] = True
    s.bundle.est.OP[:] = True
    s.prior.OP.isCtrl = np.zeros(len(ids), bool)
    for r in prob['ctrlPts']:                              # weighted control points: prior observations
        j = op_of[int(r[0])]
        s.prior.OP.val[:, j] = r[1:4]
        s.prior.OP.std[:, j] = r[4:7]
        s.prior.OP.use[:, j] = True
        s.prior.OP.isCtrl[j] = True
    return s


def camcal_pm_struct(pmfile, cptfile=None, focal=7.3, keep_loaded=False, model=3):
    """The PhotoModeler-export camcal demos (`camcaldemo.m:30-98`, `camcaldemo2.m`, `camcaldemo_1ray.m`,
    `camcaldemo_missing_obs.m`; with keep_loaded `camcaldemo_no_datum.m:36-52`): `loadpm` +
    `prob2dbatstruct` (`misc/prob2dbatstruct.m:198-420`), distortion model 3, `setcamvals(s0,'default',7.3)`
    (`setcamvals.m:40-48`: cc = 7.3, pp at the sensor centre, everything else 0), everything but skew
    estimated.  Default (the demos with control points): EO and non-control OP cleared (NaN) - start
    values come from `resect` / `forwintersect` - and the points with id > 1000 fixed at the coordinates
    of `ref/camcal-fixed.txt` (`setcpt`, std 0).  keep_loaded: EO/OP start values as loaded, nothing fixed."""
    prob = load_pm_export(pmfile)
    nImg = len(prob['images'])
    ids = np.unique(np.concatenate([prob['ctrlPts'][:, 0], prob['objPts'][:, 0]])).astype(int)
    op_of = {v: i for i, v in enumerate(ids)}
    nK, nP = 3, 2
    inner = prob['job']['defCam']
    imSz = prob['job']['imSz']
    ss = inner[3:5]
    px = ss / imSz
    pxSize = np.array([[px[1]], [px[1]]])                  # prob2dbatstruct.m:238-242
    IO = np.zeros((5 + nK + nP, nImg))
    IO[0] = focal
    IO[1], IO[2] = 0.5 * ss[0], -0.5 * ss[1]               # setcamvals.m:44
    EO = np.full((6, nImg), np.nan)
    OP = np.full((3, len(ids)), np.nan)
    if keep_loaded:
        for i, im in enumerate(prob['images']):
            EO[0:3, i] = im['outer'][0:3]
            EO[3:6, i] = np.deg2rad(im['outer'][[5, 4, 3]])
        for r in np.vstack([prob['objPts'], prob['ctrlPts']]) if len(prob['ctrlPts']) else prob['objPts']:
            OP[:, op_of[int(r[0])]] = r[1:4]
    mk = prob['markPts']
    keep = np.array([int(r[1]) in op_of for r in mk])
    mk = mk[keep]
    mk = mk[np.lexsort((mk[:, 1], mk[:, 0]))]
    mstd = mk[:, 4:6].copy()
    if np.any(mstd == 0):                                  # prob2dbatstruct.m:367-373
        mstd[:] = 1.0
    s = new_struct(IO, EO, OP, mk[:, 2:4].T, mk[:, 0].astype(int), np.array([op_of[int(v)] for v in mk[:, 1]]),
                   pxSize, imSz[:, None], model, nK, nP, mstd.T)
    s.OP.id = ids
    s.OP.label = [''] * len(ids)
    s.IP.sigmas = np.unique(mstd)                          # prob2dbatstruct.m:367
    s.EO.name = [im['name'].replace('\\', '/').split('/')[-1] for im in prob['images']]
    s.IO.model.camUnit = 'mm'                              # prob2dbatstruct.m:393-405
    s.proj = NS(objUnit='m', x0desc='', title=prob['job']['title'], fileName=pmfile, cptFile=cptfile or '',
                EOfile='', UUID='')
    s.bundle.est.IO[:] = True
    s.bundle.est.IO[4, :] = False
    if abs(model) < 3:                                     # setcamest.m:46-58: no affine terms in models 1, 2
        s.bundle.est.IO[3, :] = False
    s.bundle.est.EO[:] = True
    s.bundle.est.OP[:] = True
    s.prior.OP.isCtrl = (ids > 1000) & (not keep_loaded)   # camcaldemo.m:74-78; no setcpt in the no-datum demo
    s.prior.OP.isCheck = np.zeros(len(ids), bool)
    if not keep_loaded:
        cp = {int(r[0]): (r[1], [float(v) for v in r[2:5]]) for r in load_table(cptfile)}
        for j in np.flatnonzero(s.prior.OP.isCtrl):        # setcpt.m:24-51 (fixed: std 0, not observed)
            s.OP.label[j], s.OP.val[:, j] = cp[int(ids[j])]
            s.prior.OP.val[:, j] = s.OP.val[:, j]
            s.prior.OP.std[:, j] = 0.0
            s.bundle.est.OP[:, j] = False
    return s


def set_prior_eo_positions(s, prob, path):
    """`legacyloadeotable` + `matcheo` (by image label) + `setprioreo.m:20-42` for a table
    `label,X,Y,Z,std`: prior observations of the camera positions."""
    names = [im['name'].replace('\\', '/').split('/')[-1].lower() for im in prob['images']]
    for r in load_table(path):
        i = names.index(r[0].lower())
        s.prior.EO.val[0:3, i] = [float(v) for v in r[1:4]]
        s.prior.EO.std[0:3, i] = float(r[4])
        s.prior.EO.use[0:3, i] = float(r[4]) != 0
        s.bundle.est.EO[0:3, i] = float(r[4]) != 0
    if getattr(s, 'proj', None) is not None:
        s.proj.EOfile = path
    return s
