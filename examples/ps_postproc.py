"""The reference's `code/demo/ps_postproc.m` on the device path.

    python examples/ps_postproc.py [project.psz [report.txt [minRays [minAngle]]]]

Loads a PhotoScan archive, keeps PhotoScan's own orientation and tie points as start values, uses the
enabled markers as weighted control points and runs the bundle with the forward (computer vision) lens
model; writes the result file.  Needs a CUDA device.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dbat_b200 as dbat                                                     # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')
fileName = sys.argv[1] if len(sys.argv) > 1 else os.path.join(GOLD, 'prague2016sxb', 'psprojects', 'sxb.psz')
reportFile = sys.argv[2] if len(sys.argv) > 2 else os.path.splitext(os.path.basename(fileName))[0] + '-dbatreport.txt'

psz = dbat.loadpsz(fileName)                                                 # ps_postproc.m:56
prob = dbat.ps2pmstruct(psz)                                                 # loadplotpsz.m:48
s0 = dbat.prob2dbatstruct(prob)                                              # loadplotpsz.m:52
minRays = int(sys.argv[3]) if len(sys.argv) > 3 else 0
minAngle = float(sys.argv[4]) if len(sys.argv) > 4 else 0.0
if minRays > 0 or minAngle > 0:                                              # loadplotpsz.m:58-90
    prob, removed = dbat.filterprob(prob, s0, minRays, minAngle)
    s0 = dbat.prob2dbatstruct(prob)
    print('Filtered %d object points.' % len(removed))
s0.IO.model.distModel[:] = -1                                                # ps_postproc.m:69
if psz.camera.isAdjusted:                                                    # :86-110: estimate what PhotoScan adjusted
    s0 = dbat.setcamest(s0, 'not', 'all')
    s0 = dbat.setcamest(s0, 'cc', 'px', 'py')
    for n in range(min(3, len(psz.camera.k))):
        s0 = dbat.setcamest(s0, 'K%d' % (n + 1))
    for n in range(min(2, len(psz.camera.p))):
        s0 = dbat.setcamest(s0, 'P%d' % (n + 1))
result, ok, iters, sigma0, E = dbat.bundle(s0, 'gna', 20, 'trace')           # :124
print('Bundle %s after %d iterations with sigma0=%.2f (%.2f pixels)'
      % ('ok' if ok else 'failed (code %d)' % E.code, iters, sigma0, sigma0 * s0.IP.sigmas[0]))
result, _ = dbat.bundle_result_file(result, E, reportFile)                   # :135
print('Bundle report file %s generated.' % reportFile)
