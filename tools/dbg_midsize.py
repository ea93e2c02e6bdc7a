import os, sys, copy
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import __graft_entry__ as g
g.build()
import dbat_b200
from dbat_b200.synth import make_scene
from oracle import lsa
from oracle.bundle import bundle as obundle
from oracle.cameramodel import brown_euler_cam4
from oracle.dbatstruct import buildweightmatrix, serialize
s, _ = make_scene(100, 20000, rays=10, seed=20240607)
n = s.bundle.serial.n; nOPx = len(s.bundle.serial.OP.dest)
R = np.sqrt(buildweightmatrix(s))
s1, s2 = copy.deepcopy(s), copy.deepcopy(s)
s1, ok, it, s0, E = dbat_b200.bundle(s1, 'gna')
lsa.set_ordering(np.concatenate([np.arange(n - nOPx, n), np.arange(n - nOPx)[::-1]]))
s2, oko, ito, s0o, Eo = obundle(s2, 'gna')
lsa.set_ordering(None)
print('res dev', ['%.12f' % v for v in E.res]); print('res orc', ['%.12f' % v for v in Eo.res])
print('alphas', E.damping if hasattr(E,'damping') else None)
for k in range(E.trace.shape[1]):
    d = np.abs(E.trace[:, k] - Eo.trace[:, k]); 
    rr_d = np.linalg.norm(R * brown_euler_cam4(E.trace[:, k], s, False)[0]); rr_o = np.linalg.norm(R * brown_euler_cam4(Eo.trace[:, k], s, False)[0])
    print('iter', k, 'max|dx|', d.max(), 'argmax', d.argmax(), 'rel', (d / np.maximum(np.abs(Eo.trace[:, k]), 1e-300)).max(), 'cpu rr at dev x %.12f at orc x %.12f' % (rr_d, rr_o))
