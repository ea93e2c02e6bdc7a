"""The reference's `code/demo/camcaldemo.m`, call for call, on the device path.

    python examples/camcaldemo.py [export.txt [ctrlpts.txt [report.txt]]]

Loads a PhotoModeler export (default: the camcal project shipped as a test fixture), sets default
camera values, fixes the control points, computes start values by spatial resection and forward
intersection, runs the self-calibrating bundle (Gauss-Newton-Armijo) and writes the result file.
Needs a CUDA device: every numerical step runs in libdbatgpu.so.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dbat_b200 as dbat                                                     # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'camcalpm')
inputFile = sys.argv[1] if len(sys.argv) > 1 else os.path.join(GOLD, 'camcal-pmexport.txt')
cptFile = sys.argv[2] if len(sys.argv) > 2 else os.path.join(GOLD, 'camcal-fixed.txt')
reportFile = sys.argv[3] if len(sys.argv) > 3 else 'camcal-dbatreport.txt'

prob = dbat.loadpm(inputFile)                                                # camcaldemo.m:47
s0 = dbat.prob2dbatstruct(prob)                                              # :59
s0.IO.model.distModel[:] = 3                                                 # :62
s0 = dbat.setcamvals(s0, 'default', 7.3)                                     # :65
s0 = dbat.setcamest(s0, 'all', 'not', 'sk')                                  # :68
s0 = dbat.seteoest(s0, 'all')                                                # :71
s0 = dbat.cleareo(s0)                                                        # :75
if not s0.prior.OP.isCtrl.any():                                             # :78-80
    s0.prior.OP.isCtrl = np.asarray(s0.OP.id) > 1000
pts = dbat.loadcpt(cptFile)                                                  # :84
i, j = dbat.matchcpt(s0, pts)                                                # :87
s0 = dbat.setcpt(s0, pts, i, j)                                              # :90
s0 = dbat.clearop(s0)                                                        # :97
cpId = np.asarray(s0.OP.id)[s0.prior.OP.isCtrl]                              # :100-107
s1, _, fail = dbat.resect(s0, 'all', cpId, 1, 0, cpId)
if fail:
    sys.exit('Resection failed.')
s2, _, _ = dbat.forwintersect(s1, 'all', True)
s2.proj.x0desc = 'Camera calibration from EXIF value'                       # :109
result, ok, iters, sigma0, E = dbat.bundle(s2, 'gna', 'trace')                # :118
if ok:
    print('Bundle ok after %d iterations with sigma0=%.2f (%.2f pixels)' % (iters, sigma0, result.post.sigmas[0]))
else:
    print('Bundle failed after %d iterations (code=%d). Last sigma0 estimate=%.2f (%.2f pixels)'
          % (iters, E.code, sigma0, sigma0 * s0.IP.sigmas[0]))
result, _ = dbat.bundle_result_file(result, E, reportFile)                   # :131-135
print('Bundle result file %s generated.' % reportFile)
