"""Forward intersection at config-4 scale: device kernel time, end-to-end host call, oracle sample (GPU box)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import dbat_b200
from dbat_b200.synth import make_scene
from oracle.photogrammetry import forwintersect as ofwi

nImg, nOP = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1000, 200000)
s, truth = make_scene(nImg, nOP, rays=10, seed=20240607, build_indices=False)
s.IO.val[:] = truth['IO'][:, None]
s.EO.val[:] = truth['EO']
best = 1e9
for _ in range(3):
    t = time.time(); sg, ids, res, ms = dbat_b200.forwintersect(s, 'all', False, return_ms=True); best = min(best, time.time() - t)
err = np.abs(sg.OP.val - truth['OP']).max()
print('forwintersect %d images x %d points x %d observations: kernel %.3f ms, host call %.1f ms, max |OP - truth| %.3g (0.5 px noise)'
      % (nImg, nOP, s.IP.val.shape[1], ms, best * 1e3, err))
# oracle on a bounded sample of the same generator (1/100 of the points)
ss, tt = make_scene(max(20, nImg // 10), max(200, nOP // 100), rays=10, seed=20240607, build_indices=False)
ss.IO.val[:] = tt['IO'][:, None]; ss.EO.val[:] = tt['EO']
t = time.time(); so, _, _ = ofwi(ss, 'all', False); dt = time.time() - t
sg2, _, _ = dbat_b200.forwintersect(ss, 'all', False)
print('oracle restatement: %d points in %.2f s = %.0f points/s (device: %.2e points/s on the full block); sample agreement %.2e'
      % (ss.OP.val.shape[1], dt, ss.OP.val.shape[1] / dt, nOP / (ms * 1e-3), np.nanmax(np.abs(so.OP.val - sg2.OP.val))))
