"""DBAT XML script runner (`code/script/rundbatscript.m` and the parsers it calls).

    python -m dbat_b200.script project.xml

Reads a `<document dbat_script_version="1.0">` script - input section (cameras, images, image points,
control / check points, prior EO), operations (check_ray_count, set_initial_values,
set_bundle_estimate_params, spatial_resection, forward_intersection, set_datum, bundle_adjustment) and
output files (report, io) - builds the flat DBAT struct and runs the operations on the device path.

  rundbatscript            `script/rundbatscript.m:22-74`
  _parse_input             `script/parseinput.m:20-105`, `parsecameras.m`, `parsedbatxmlcamstruct.m:92-119`
                           (`DBATCamera.m:56-134` conventions), `parseimages.m`, `parseimagepts.m`,
                           `parsectrlpts.m`, `parseprioreo.m`, `setdbatcamsandimages.m`, `setdbatpts.m`,
                           table readers `file/loadimagepts.m`, `loadctrlpts.m`, `loadeotable.m`, `loadimagetable.m`
  _run_operations          `script/parseops.m:30-116`, `parsesetinitialvalues.m`, `parsesetbundleest.m`,
                           `parsesetdatum.m` and `script/private/parseset*.m`
  _write_outputs           `script/parseoutputfiles.m:29-218` (report, io, eo and image_residuals files; plots are not produced)

`backend` is a namespace with resect / forwintersect / bundle / bundle_cov; the default is this package
(the device).  Tests hand in the oracle so the runner itself is checked on CPU against the three script
projects the reference ships with their result files.
"""
import os
import sys
import uuid as _uuid
import xml.etree.ElementTree as ET
from types import SimpleNamespace as NS

import numpy as np

from . import ingest
from .dbatstruct import new_struct


def _text(node, tag, default=None):
    e = node.find(tag)
    return e.text.strip() if e is not None and e.text is not None else default


def _floats(txt):
    return [float(v) for v in txt.split(',') if v.strip()]


def _path(txt, base, here):
    txt = txt.strip().replace('$HERE', here)
    return txt if os.path.isabs(txt) or not base else os.path.join(base, txt)


def _camera(node):
    """One <camera> element, user conventions (DBATCamera.m) -> internal values
    (parsedbatxmlcamstruct.m:92-119): py, K, P change sign, aspect ratio a -> 1 - a, pixels square with
    the size given by the sensor height, 'auto' sensor width from the image width."""
    g = lambda tag, d=None: _text(node, tag, d)
    imSize = np.array(_floats(g('image')))
    sens = [t.strip() for t in g('sensor').split(',')]
    h = float(sens[1])
    px = h / imSize[1]
    w = px * imSize[0] if sens[0] == 'auto' else float(sens[0])
    K = _floats(g('K', '0,0,0'))
    P = _floats(g('P', '0,0'))
    nK, nP = int(g('nK', len(K))), int(g('nP', len(P)))
    io = np.zeros(5 + nK + nP)
    if g('cc') is not None:
        io[0] = float(g('cc'))
    if g('pp') is not None:
        pp = _floats(g('pp'))
        io[1], io[2] = pp[0], -pp[1]
    io[3] = 1 - float(g('aspect', '1'))
    io[4] = float(g('skew', '0'))
    io[5:5 + len(K)] = [-v for v in K][:nK]
    io[5 + nK:5 + nK + len(P)] = [-v for v in P][:nP]
    return NS(name=g('name', ''), unit=g('unit', 'mm'), io=io, ssSize=np.array([w, h]), imSize=imSize,
              pxSize=np.array([px, px]), focal=float(g('focal', 'nan')), model=int(g('model', '3')),
              nK=nK, nP=nP, id=g('id', '1'), calibrated=g('calibrated', ''))


def _ctrl_pts(node, base, here):
    """parsectrlpts.m: the table, then the keep / remove id filters."""
    f = node.find('file')
    pts = ingest.loadctrlpts(_path(f.text, base, here), f.get('format'))
    for flt in node.findall('filter'):
        sel = np.isin(pts.id, [int(v) for v in flt.get('id').split(',') if v.strip()])
        what = flt.text.strip()
        if what not in ('keep', 'remove'):
            raise ValueError('DBAT XML input/ctrl_pts/filter error: Unknown filter %s' % what)
        keep = sel if what == 'keep' else ~sel
        pts.id, pts.pos, pts.std = pts.id[keep], pts.pos[:, keep], pts.std[:, keep]
        pts.name = [n for n, k in zip(pts.name, keep) if k]
    return pts


def _parse_input(inp, docFile):
    here = os.path.dirname(os.path.abspath(docFile))
    base = inp.get('base_dir', '').replace('$HERE', here)
    known = {'ctrl_pts', 'check_pts', 'images', 'prior_eo', 'image_pts', 'cameras', 'c'}
    unknown = [c.tag for c in inp if c.tag not in known]
    missing = [t for t in ('images', 'image_pts', 'cameras') if inp.find(t) is None]
    if unknown or missing:                                              # parseinput.m:24-28 (checkxmlfields)
        raise ValueError('DBAT XML script input error: %s' % '; '.join(
            (['unknown field(s) ' + ', '.join(unknown)] if unknown else []) +
            (['missing field(s) ' + ', '.join(missing)] if missing else [])))
    for t in ('images', 'image_pts', 'ctrl_pts', 'check_pts', 'prior_eo'):
        if inp.find(t) is not None and inp.find(t).find('file') is None:
            raise ValueError('DBAT XML input/%s error: missing file' % t)
    camsNode = inp.find('cameras')
    cams = [_camera(c) for c in camsNode.findall('camera')]
    for f in camsNode.findall('file'):
        cams += [_camera(c) for c in ET.parse(_path(f.text, base, here)).getroot().iter('camera')]
    if len(cams) != 1:
        raise NotImplementedError('exactly one camera is supported by the device path (got %d)' % len(cams))
    cam = cams[0]
    f = inp.find('images').find('file')
    ims = ingest.loadimagetable(_path(f.text, base, here), f.get('format'))
    imId, imPath = ims.id, [p.replace('\\', '/') for p in ims.path]
    imDir = os.path.commonpath([os.path.dirname(p) for p in imPath]) if all('/' in p for p in imPath) else ''
    ip = {'id': [], 'im': [], 'x': [], 'y': [], 'sx': [], 'sy': []}
    for f in inp.find('image_pts').findall('file'):
        pts = ingest.loadimagepts(_path(f.text, base, here), f.get('format'))
        if f.get('sxy') is not None:                                    # parseimagepts.m:52-63
            pts.std[:] = float(f.get('sxy'))
        if f.get('sx') is not None:
            pts.std[0] = float(f.get('sx'))
        if f.get('sy') is not None:
            pts.std[1] = float(f.get('sy'))
        for k, v in (('id', pts.id), ('im', pts.im), ('x', pts.pos[0]), ('y', pts.pos[1]), ('sx', pts.std[0]), ('sy', pts.std[1])):
            ip[k].append(v)
    ip = {k: np.concatenate(v) for k, v in ip.items()}
    none = NS(id=np.zeros(0, np.int64), name=[], pos=np.zeros((3, 0)), std=np.zeros((3, 0)), fileName='')
    ctrl = _ctrl_pts(inp.find('ctrl_pts'), base, here) if inp.find('ctrl_pts') is not None else none
    check = _ctrl_pts(inp.find('check_pts'), base, here) if inp.find('check_pts') is not None else none
    if len(np.intersect1d(ctrl.id, check.id)):
        raise ValueError('Point cannot be both control and check points')
    # setdbatpts.m
    OPid = np.unique(np.concatenate([ctrl.id, check.id, ip['id'].astype(np.int64)]))
    nOP, nImg = len(OPid), len(imId)
    im_of = {v: i for i, v in enumerate(imId)}
    s = new_struct(np.full((len(cam.io), nImg), np.nan), np.full((6, nImg), np.nan), np.full((3, nOP), np.nan),
                   np.vstack([ip['x'], ip['y']]), np.array([im_of[int(v)] for v in ip['im']]),
                   np.searchsorted(OPid, ip['id'].astype(np.int64)), cam.pxSize[:, None], cam.imSize[:, None],
                   cam.model, cam.nK, cam.nP, np.vstack([ip['sx'], ip['sy']]))
    s.OP.id = s.OP.rawId = OPid
    s.OP.label = [''] * nOP
    s.prior.OP.isCtrl = np.isin(OPid, ctrl.id)
    s.prior.OP.isCheck = np.isin(OPid, check.id)
    for tab in (ctrl, check):
        col = np.searchsorted(OPid, tab.id)
        s.prior.OP.val[:, col] = tab.pos
        s.prior.OP.std[:, col] = tab.std
        for c, n in zip(col, tab.name):
            s.OP.label[c] = n
    s.prior.OP.use = ~np.isnan(s.prior.OP.std) & (s.prior.OP.std != 0) & ~s.prior.OP.isCheck[None, :]
    s.IP.sigmas = np.unique(s.IP.std)
    # setdbatcamsandimages.m
    s.IO.model.camUnit = cam.unit
    s.IO.sensor.ssSize = np.tile(cam.ssSize[:, None], (1, nImg))
    s.prior.IO.val[:] = cam.io[:, None]
    s.prior.IO.cams = cams
    s.EO.name = [p[len(imDir) + 1:] if imDir else p for p in imPath]
    s.EO.label = list(s.EO.name)
    s.EO.id = imId
    s.bundle.est.IO[:] = False
    s.bundle.est.EO[:] = False
    s.bundle.est.OP[:] = False
    EOfile = ''
    if inp.find('prior_eo') is not None:
        f = inp.find('prior_eo').find('file')
        EOfile = _path(f.text, base, here)
        tbl = ingest.loadeotable(EOfile, f.get('format'))
        scale = {'radian': 1.0, 'degrees': np.pi / 180, 'gon': np.pi / 200}.get(f.get('units'))
        if np.any(~np.isnan(tbl.ang)) and scale is None:
            raise ValueError('DBAT XML input/prior_eo: angles need a units attribute')
        col = np.array([im_of[int(v)] for v in tbl.id])
        s.prior.EO.val[0:3, col] = tbl.pos
        s.prior.EO.val[3:6, col] = tbl.ang * (scale or 1.0)
        s.prior.EO.std[0:3, col] = tbl.std
        s.prior.EO.std[3:6, col] = tbl.angStd * (scale or 1.0)
    return s, imDir, ctrl.fileName, EOfile


def _set_initial_io(s, node):
    cam = s.prior.IO.cams[0]
    fields = [('all', node.text.strip())] if node.text and node.text.strip() else \
        [(c.tag, (c.text or '').strip()) for c in node if c.tag != 'c']
    prior = s.prior.IO.val
    for tag, txt in fields:
        if tag == 'all':
            if txt == 'loaded':
                ingest.setcamvals(s, 'loaded')
            elif txt == 'default':
                ingest.setcamvals(s, 'default', cam.focal)
            else:
                raise ValueError("DBAT XML set initial values/IO error: Unknown 'all' string '%s'" % txt)
        elif tag == 'cc':
            s.IO.val[0] = cam.focal if txt in ('focal', 'default') else prior[0] if txt == 'loaded' else float(txt)
        elif tag == 'pp':
            if txt == 'default':
                s.IO.val[1], s.IO.val[2] = 0.5 * s.IO.sensor.ssSize[0], -0.5 * s.IO.sensor.ssSize[1]
            elif txt == 'loaded':
                s.IO.val[1:3] = prior[1:3]
            else:
                s.IO.val[1:3] = np.reshape(_floats(txt), (2, 1))
        elif tag == 'aspect':
            s.IO.val[3] = 0 if txt == 'default' else prior[3] if txt == 'loaded' else 1 - float(txt)
        elif tag == 'skew':
            s.IO.val[4] = 0 if txt == 'default' else prior[4] if txt == 'loaded' else float(txt)
        elif tag in ('K', 'P'):
            rows = ingest._io_rows(s, tag)
            s.IO.val[rows] = 0 if txt == 'default' else prior[rows] if txt == 'loaded' else np.reshape(_floats(txt), (-1, 1))
        else:
            raise ValueError("DBAT XML set initial values/IO error: Unknown field '%s'" % tag)


def _fields(node, whole=('true', 'false', 'default')):
    if node.text and node.text.strip():
        if node.text.strip() not in whole:
            raise ValueError("DBAT XML script error: Unknown string '%s'" % node.text.strip())
        return [('all', node.text.strip())]
    return [(c.tag, (c.text or '').strip()) for c in node if c.tag != 'c']


def _run_operations(s, ops, backend, log):
    E = None
    for op in ops.findall('operation'):
        kids = [c for c in op if c.tag != 'c']
        if not kids:
            name = (op.text or '').strip()
            if name == 'check_ray_count':
                minRays = int(op.get('min_rays', 2))
                rays = np.bincount(s.IP.op, minlength=s.OP.val.shape[1])
                bad = np.flatnonzero(rays < minRays)
                if np.any((rays < minRays) & ~s.prior.OP.isCtrl):
                    for j in bad:
                        log('Object point number %d (id %d) has too few rays: %d.' % (j + 1, s.OP.id[j], rays[j]))
                    raise ValueError('DBAT XML error: Ray count test failed. See above for details.')
            elif name == 'spatial_resection':
                cpId = np.asarray(s.OP.id)[s.prior.OP.isCtrl]
                s, _, fail = backend.resect(s, 'all', cpId, 1, 0, cpId)
                if fail:
                    raise RuntimeError('Resection failed.')
            elif name == 'forward_intersection':
                s = backend.forwintersect(s, 'all', True)[0]
            elif name == 'bundle_adjustment':
                s0 = s
                s, ok, iters, sigma0, E = backend.bundle(s0)
                if ok:
                    log('Bundle ok after %d iterations with sigma0=%.2f (%.2f pixels)'
                        % (iters, sigma0, np.ravel(s.post.sigmas)[0]))
                else:
                    log('Bundle failed after %d iterations (code=%d). Last sigma0 estimate=%.2f (%.2f pixels)'
                        % (iters, E.code, sigma0, sigma0 * np.ravel(s0.IP.sigmas)[0]))
                s.bundle.info = E
            else:
                raise ValueError('DBAT XML script operations error: Unknown operation %s' % name)
            continue
        if len(kids) > 1:
            raise ValueError('DBAT XML script operations error: Too many fields in one operation: %s'
                             % ', '.join(k.tag for k in kids))
        node = kids[0]
        if node.tag == 'set_initial_values':
            for sub in node:
                if sub.tag == 'io':
                    _set_initial_io(s, sub)
                elif sub.tag in ('eo', 'op'):
                    for tag, txt in _fields(sub, ('loaded',)):
                        if tag != 'all' or txt != 'loaded':
                            raise ValueError("DBAT XML set initial values/%s error: Unknown field '%s'" % (sub.tag.upper(), tag))
                        tgt = s.EO if sub.tag == 'eo' else s.OP
                        tgt.val = (s.prior.EO if sub.tag == 'eo' else s.prior.OP).val.copy()
                elif sub.tag != 'c':
                    raise ValueError('DBAT XML script set_initial_values error: unknown field %s' % sub.tag)
        elif node.tag == 'set_bundle_estimate_params':
            for sub in node:
                if sub.tag == 'io':
                    for tag, txt in _fields(sub, ('true', 'false')):
                        tag = {'aspect': 'as', 'skew': 'sk'}.get(tag, tag)
                        if txt not in ('true', 'false'):
                            raise ValueError('DBAT XML script set_bundle_estimate_params/IO error: Bad string %s' % txt)
                        ingest.setcamest(s, tag) if txt == 'true' else ingest.setcamest(s, 'not', tag)
                elif sub.tag == 'eo':
                    rows = {'all': range(6), 'pos': range(3), 'angles': range(3, 6), 'x': [0], 'y': [1], 'z': [2]}
                    for tag, txt in _fields(sub):
                        ix = list(rows[tag])
                        s.bundle.est.EO[ix] = (s.prior.EO.std[ix] != 0) if txt == 'default' else (txt == 'true')
                elif sub.tag == 'op':
                    rows = {'all': range(3), 'x': [0], 'y': [1], 'z': [2]}
                    for tag, txt in _fields(sub):
                        ix = list(rows[tag])
                        s.bundle.est.OP[ix] = (~s.prior.OP.isCtrl[None, :] | (s.prior.OP.std[ix] != 0)) \
                            if txt == 'default' else (txt == 'true')
                elif sub.tag != 'c':
                    raise ValueError('DBAT XML script set_bundle_estimate_params error: unknown field %s' % sub.tag)
        elif node.tag == 'set_datum':
            if (node.text or '').strip() != 'depend':
                raise ValueError('DBAT XML operation set_datum error: Unknown datum %s' % node.text)
            base = node.get('ref_base', 'longest')
            if base == 'longest':
                ingest.seteoest(s, 'depend', int(node.get('ref_cam', 1)))
            elif base in ('x', 'y', 'z'):
                ingest.seteoest(s, 'depend', int(node.get('ref_cam', 1)), base)
            else:
                raise ValueError("DBAT XML operation set_datum error: Unknown reference base '%s'" % base)
        else:
            raise ValueError('DBAT XML script operations error: Unknown operation %s' % node.tag)
    return s, E


def _camera_xml(s, cam):
    """The calibrated camera in user conventions (parseoutputfiles.m io output / DBATCamera)."""
    io = s.IO.val[:, 0]
    nK, nP = s.IO.model.nK, s.IO.model.nP
    g17 = lambda v: '0' if v == 0 else ('%.18g' % v)
    ss = s.post.sensor.ssSize[:, 0] if getattr(s.post, 'sensor', None) is not None else cam.ssSize
    rows = [('id', cam.id), ('name', cam.name), ('unit', cam.unit), ('calibrated', 'yes'),
            ('sensor', '%s,%s' % ('%.6g' % ss[0], '%.6g' % ss[1])), ('image', '%d,%d' % tuple(cam.imSize)),
            ('aspect', g17(1 - io[3])), ('focal', '%g' % cam.focal), ('model', '%d' % cam.model),
            ('nK', '%d' % nK), ('nP', '%d' % nP), ('cc', g17(io[0])), ('pp', '%s,%s' % (g17(io[1]), g17(-io[2]))),
            ('skew', g17(io[4])), ('K', ','.join(g17(-v) for v in io[5:5 + nK])),
            ('P', ','.join(g17(-v) for v in io[5 + nK:5 + nK + nP]))]
    body = '\n'.join('         <%s>%s</%s>' % (k, v, k) for k, v in rows)
    return ('<?xml version="1.0" encoding="utf-8"?>\n<document dbat_camera_version="1.0">\n   <cameras>\n'
            '      <camera>\n%s\n      </camera>\n   </cameras>\n</document>\n' % body)


def _write_outputs(s, E, out, docFile, backend, write):
    from .report import bundle_result_file
    here = os.path.dirname(os.path.abspath(docFile))
    files = out.find('files') if out is not None else None
    lines = None
    if files is None or E is None:
        return s, lines
    base = files.get('base_dir', '').replace('$HERE', here)
    for node in files:
        if node.tag == 'report':
            f = _path(node.find('file').text, base, here)
            s, lines = bundle_result_file(s, E, f if write else None, cov=backend.bundle_cov)
        elif node.tag == 'io' and write and E.code == 0:
            with open(_path(node.find('file').text, base, here), 'wt') as fh:
                fh.write(_camera_xml(s, s.prior.IO.cams[0]))
        elif node.tag == 'eo' and write:
            with open(_path(node.find('file').text, base, here), 'wt') as fh:
                fh.write('\n'.join(_eo_listing(s, E, docFile, backend)) + '\n')
        elif node.tag == 'image_residuals' and write:
            with open(_path(node.find('file').text, base, here), 'wt') as fh:
                fh.write('\n'.join(_residual_listing(s, docFile, int(node.get('top_count', 1000)))) + '\n')
    return s, lines


def _listing_head(s, what, docFile):
    return ['# %s for dbat script file' % what, '# %s.' % docFile, '# Generated by dbat_b200.',
            '# Execution UUID: %s.' % s.proj.UUID]


def _eo_listing(s, E, docFile, backend):
    """parseoutputfiles.m WritePostEOFile: one line per image - number, id, EO (angles in radians), the six
    posterior standard deviations times 180/pi (positions included, as the reference writes them), label."""
    sd = getattr(s.post.std, 'EO', None)
    if sd is None:
        from .report import _diag_blocks
        sd = np.sqrt(np.diagonal(_diag_blocks(backend.bundle_cov(s, E, 'CEO'), 6), axis1=1, axis2=2)).T
    out = _listing_head(s, 'EO listing', docFile)
    out.append('# Format: EO number, EO id, x, y, z, omega, phi, kappa, sx, sy, sz, so, sk, label. Unit: degrees.')
    for i in range(s.EO.val.shape[1]):
        nums = list(s.EO.val[:, i]) + list(sd[:, i] * 180 / np.pi)
        out.append('%d, %d, ' % (i + 1, s.EO.id[i]) + ''.join('%.18g, ' % v for v in nums) + s.EO.label[i])
    return out


def _residual_listing(s, docFile, topCount):
    """parseoutputfiles.m WriteImageResidualsFile: the topCount largest image residuals (pixels)."""
    res = s.post.res.IP
    nrm = np.sqrt((res ** 2).sum(axis=0))
    top = np.argsort(-nrm, kind='stable')[:topCount]
    out = _listing_head(s, 'Top residual listing', docFile)
    out += ['# Listing top %d residuals.' % topCount, '# Format: OP id, image id, x, y, resx, resy, resTot']
    for k in top:
        out.append('%d, %d, %g, %g, %g, %g, %g' % (s.OP.id[s.IP.op[k]], s.EO.id[s.IP.img[k]], s.IP.val[0, k],
                                                   s.IP.val[1, k], res[0, k], res[1, k], nrm[k]))
    return out


def rundbatscript(f, verbose=False, backend=None, write=True):
    """rundbatscript.m: run the XML script f.  Returns (s, E, report_lines); with write=False no output
    file is written (the report text is still built and returned)."""
    if backend is None:
        import dbat_b200 as backend
    log = print if verbose else (lambda *a: None)
    doc = ET.parse(f).getroot()
    if doc.tag != 'document' or doc.get('dbat_script_version') is None:
        raise ValueError('DBAT XML top level error: not a DBAT script document')
    if tuple(int(v) for v in doc.get('dbat_script_version').split('.')[:2]) != (1, 0):
        raise ValueError('DBAT script error: unsupported dbat_script_version %s' % doc.get('dbat_script_version'))
    for part in ('input', 'operations', 'output'):
        if doc.find(part) is None:
            raise ValueError('DBAT XML document field error: missing %s' % part)
    meta = doc.find('meta')
    name = _text(meta, 'name', '') if meta is not None else ''
    unit = _text(meta, 'project_unit', 'm') if meta is not None else 'm'
    s, imDir, cptFile, EOfile = _parse_input(doc.find('input'), f)
    s.proj = NS(objUnit=unit, x0desc='', title=name, imDir=imDir, fileName=os.path.abspath(f), cptFile=cptFile,
                EOfile=EOfile, UUID=str(_uuid.uuid4()))
    s, E = _run_operations(s, doc.find('operations'), backend, log)
    s, lines = _write_outputs(s, E, doc.find('output'), f, backend, write)
    return s, E, lines


if __name__ == '__main__':
    if len(sys.argv) != 2:
        sys.exit(__doc__)
    rundbatscript(sys.argv[1], verbose=True)
