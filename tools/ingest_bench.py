"""Ingest (loadpm -> prob2dbatstruct) and report-side statistics at BASELINE config-4 scale, host only.

Writes a synthetic PhotoModeler export of nImg x nOP x (nOP*rays) observations, then times this package's
vectorised ingest next to a literal restatement of the reference's struct builder (the per-image loop of
`prob2dbatstruct.m:349-365`: one full-table mask, a sort and sparse vis/ix writes per image).  Runs on
any host (no GPU).  Usage: python tools/ingest_bench.py [nImg nOP]"""
import json, os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import scipy.sparse as sp
from dbat_b200 import ingest
from dbat_b200.synth import make_scene

nImg, nOP = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1000, 200000)
s, truth = make_scene(nImg, nOP, rays=10, seed=20240607, build_indices=False)
nObs = s.IP.val.shape[1]
path = os.path.join(tempfile.gettempdir(), 'ingest_bench_%d_%d.txt' % (nImg, nOP))
t = time.time()
with open(path, 'w') as fh:
    ss = s.IO.sensor.imSize[:, 0] * s.IO.sensor.pxSize[:, 0]
    io = truth['IO']
    fh.write('Synthetic block\n 0.0005 20 %d %d\n 1 0.1 10 100 100 100 20 20 20\n' % tuple(s.IO.sensor.imSize[:, 0]))
    cam = ' %.6f %.6f %.6f %.6f %.6f %.3e %.3e %.3e %.3e %.3e\n' % (io[0], io[1], -io[2], ss[0], ss[1], -io[5], -io[6], -io[7], -io[8], -io[9])
    fh.write(cam + ' 0 0 0 0 0 0 0 0 0 0\n')
    for i in range(nImg):
        e = truth['EO'][:, i]
        fh.write('%d images/img%05d.jpg\n%d %.6f %.6f %.6f %.6f %.6f %.6f\n%d 0 0 0 0 0 0\n\n%d%s%d 0 0 0 0 0 0 0 0 0 0\n'
                 % (i, i, i, e[0], e[1], e[2], np.rad2deg(e[5]), np.rad2deg(e[4]), np.rad2deg(e[3]), i, i, cam, i))
    fh.write('\n\n')                                       # end of images, empty control point table
    np.savetxt(fh, np.column_stack([np.arange(1, nOP + 1), truth['OP'].T, np.zeros((nOP, 3))]), fmt='%d %.5f %.5f %.5f %g %g %g')
    fh.write('\n')
    np.savetxt(fh, np.column_stack([s.IP.img, s.IP.op + 1, s.IP.val.T, np.full((nObs, 2), 0.5)]), fmt='%d %d %.4f %.4f %g %g')
    fh.write('\n')
t_write = time.time() - t
size_mb = os.path.getsize(path) / 1e6

t = time.time(); prob = ingest.loadpm(path); t_load = time.time() - t
t = time.time(); s2 = ingest.prob2dbatstruct(prob); t_struct = time.time() - t
assert s2.IP.val.shape[1] == nObs and np.array_equal(s2.IP.img, s.IP.img) and np.array_equal(s2.IP.op, s.IP.op)


def reference_loop(prob, nImages, OPid):
    """prob2dbatstruct.m:340-365 as written: per image a mask over the whole table, a sort, ismember, and
    column writes into sparse vis / colPos."""
    nOP = len(OPid)
    nMark = len(prob.markPts)
    markPts = np.full((2, nMark), np.nan)
    vis = sp.lil_matrix((nOP, nImages), dtype=bool)
    colPos = sp.lil_matrix((nOP, nImages))
    ii = 0
    for i in range(nImages):
        j = prob.markPts[:, 0] == i
        measured = prob.markPts[j]
        measured = measured[np.argsort(measured[:, 1], kind='stable')]
        valid = np.isin(measured[:, 1], OPid)
        n = int(valid.sum())
        markPts[:, ii:ii + n] = measured[valid, 2:4].T
        rows = np.flatnonzero(np.isin(OPid, measured[:, 1]))
        vis[rows, i] = True
        colPos[rows, i] = np.arange(ii + 1, ii + n + 1)
        ii += n
    return markPts, vis.tocsc(), colPos.tocsc()


sample = min(nImg, 50)                                      # bounded sample of the images, scaled up
t = time.time(); reference_loop(prob, sample, s2.OP.id); t_ref = (time.time() - t) * nImg / sample

from dbat_b200.report import angles, coverage
s2.OP.val[:] = truth['OP']
t = time.time(); a = angles(s2); t_angles = time.time() - t
t = time.time(); coverage(s2, np.arange(nImg)); coverage(s2, np.arange(nImg), True); t_cov = time.time() - t
print(json.dumps({'workload': '%d images x %d points x %d observations (%.0f MB export)' % (nImg, nOP, nObs, size_mb),
                  'loadpm_s': round(t_load, 2), 'prob2dbatstruct_s': round(t_struct, 2),
                  'observations_per_s': round(nObs / (t_load + t_struct)),
                  'reference_style_struct_loop_s': round(t_ref, 1), 'loop_sampled_images': sample,
                  'report_angles_s': round(t_angles, 2), 'report_coverage_s': round(t_cov, 2)}))
os.remove(path)
