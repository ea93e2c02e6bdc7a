"""The reference's `code/demo/prague2016_pm.m` on the device path.

    python examples/prague2016_pm.py [c1|c2|s1|s2|s3|s4] [no|with] [report.txt]

PhotoModeler export with a loaded (fixed) camera and control points that are fixed (c1, s1) or weighted;
start values by spatial resection and forward intersection; Gauss-Newton-Armijo; result file; and the
reference's external verification: the adjusted points against PhotoModeler's own 3-D table.
Needs a CUDA device.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dbat_b200 as dbat                                                     # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')
CASES = {'c1': ('prague2016cam', 'fixed', 'fixed'), 'c2': ('prague2016cam', 'weighted', 'weighted'),
         's1': ('prague2016sxb', 'f-op0', 'fixed'), 's2': ('prague2016sxb', 'w-op0', 'weighted'),
         's3': ('prague2016sxb', 'w-op1', 'weighted'), 's4': ('prague2016sxb', 'wsmart', 'weighted')}
label = sys.argv[1] if len(sys.argv) > 1 else 'c2'
orient = sys.argv[2] if len(sys.argv) > 2 else 'no'
project, stub, cps = CASES[label]
inputDir = os.path.join(GOLD, project)
reportFile = sys.argv[3] if len(sys.argv) > 3 else '%s-%s-orient-dbatreport.txt' % (stub, orient)

prob = dbat.loadpm(os.path.join(inputDir, 'pmexports', '%s-%s-orient-pmexport.txt' % (stub, orient)))   # :60-75
s0 = dbat.prob2dbatstruct(prob)                                              # :140
s0 = dbat.setcamvals(s0, 'loaded')                                           # :144
s0 = dbat.setcamest(s0, 'not', 'all')                                        # :145
ctrlPts = dbat.loadcpt(os.path.join(inputDir, 'ref', 'ctrlpts-%s.txt' % cps))   # :149
_, ia, ib = np.intersect1d(prob.ctrlPts[:, 0], ctrlPts.id, return_indices=True)
meanOffset = np.mean(prob.ctrlPts[ia, 1:4].T - ctrlPts.pos[:, ib], axis=1, keepdims=True)   # :166-175
ctrlPts.pos = ctrlPts.pos + meanOffset                                       # :186
i, j = dbat.matchcpt(s0, ctrlPts, 'id')                                      # :189
s0 = dbat.setcpt(s0, ctrlPts, i, j)                                          # :190
s0 = dbat.cleareo(s0)                                                        # :194
s0 = dbat.clearop(s0)                                                        # :195
cpId = np.asarray(s0.OP.id)[s0.prior.OP.isCtrl]
s1, _, fail = dbat.resect(s0, 'all', cpId, 1, 0, cpId)                       # :198
if fail:
    sys.exit('Resection failed.')
s2, _, _ = dbat.forwintersect(s1, 'all', True)                               # :200
result, ok, iters, sigma0, E = dbat.bundle(s2, 'gna', 'trace', 'dofverb')    # :215
print('Bundle %s after %d iterations with sigma0=%.2f (%.2f pixels)'
      % ('ok' if ok else 'failed (code %d)' % E.code, iters, sigma0, sigma0 * s2.IP.sigmas[0]))
result, _ = dbat.bundle_result_file(result, E, reportFile)                   # :224
print('Bundle report file %s generated.' % reportFile)
tbl = os.path.join(inputDir, 'pmexports', '%s-%s-orient-3dpts.txt' % (stub, orient))
if os.path.exists(tbl):                                                      # :255-275
    pts3d = dbat.loadpm3dtbl(tbl)
    _, a, b = np.intersect1d(pts3d.id, result.OP.id, return_indices=True)
    isCP = result.prior.OP.isCtrl[b]
    d = np.abs(result.OP.val[:, b] - meanOffset - pts3d.pos[:, a])
    ds = np.abs(result.post.std.OP[:, b] - pts3d.std[:, a])
    print('Max abs difference to PhotoModeler: OP %.3g, CP %.3g, OP std %.3g, CP std %.3g (project units)'
          % (d[:, ~isCP].max(), d[:, isCP].max(), ds[:, ~isCP].max(), ds[:, isCP].max()))
