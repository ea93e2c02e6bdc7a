"""Fixture loaders for the reference's script-format projects (test infrastructure).

Restates the subset of `code/script/{parseinput,setdbatpts,setdbatcamsandimages,
parsedbatxmlcamstruct}.m` and `code/classes/@DBATCamera/DBATCamera.m:56-134` that the
camcal XML project (`data/script/camcaldemo/`) needs: comma-separated tables with a
`format="..."` header and one XML camera.  Conventions:
internal py = -user py, internal K,P = -user K,P, internal `as` = 1 - user aspect,
square pixel size = sensor height / image height (`sensor="auto,h"`).
"""
import re

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
    s.EO.name = [r[1] for r in imgs]
    s.bundle.est.IO[:] = True
    s.bundle.est.IO[4, :] = False                        # <skew>false</skew>
    s.bundle.est.EO[:] = True
    s.bundle.est.OP[:] = True
    for r in ctrl:                                       # fixed control points
        j = op_of[int(r[0])]
        s.OP.val[:, j] = [float(r[2]), float(r[3]), float(r[4])]
        s.bundle.est.OP[:, j] = False
    s.prior.OP.isCtrl = ~s.bundle.est.OP.all(axis=0)
    return s


def load_camera_stations(path):
    """result/camera_stations.txt → (ids, EO 6xN [radians], std 6xN as printed)."""
    rows = load_table(path)
    ids = np.array([int(r[1]) for r in rows])
    vals = np.array([[float(v) for v in r[2:14]] for r in rows]).T
    return ids, vals[0:6], vals[6:12]
