"""Debug helper (GPU): print how the device's failed-run result file differs from the reference's."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import __graft_entry__ as g
g.build()
import dbat_b200
from oracle.loaders import camcal_pm_struct
from test_report_golden import report_diff, GOLD
G = os.path.join(GOLD, 'camcalpm')
s = camcal_pm_struct(os.path.join(G, 'camcal-pmexport-missing-obs.txt'), os.path.join(G, 'camcal-fixed.txt'))
s.proj.x0desc = 'Camera calibration from EXIF value'
cpId = np.asarray(s.OP.id)[s.prior.OP.isCtrl]
s1, _, fail = dbat_b200.resect(s, 'all', cpId, 1, 0, cpId)
s2, _, _ = dbat_b200.forwintersect(s1, 'all', True)
s3, ok, it, s0, E = dbat_b200.bundle(s2, 'gna')
s3, lines = dbat_b200.bundle_result_file(s3, E)
gold = [l.rstrip('\n') for l in open(os.path.join(G, 'camcal-dbatreport-missing-obs.txt'))]
import difflib
for l in list(difflib.unified_diff(gold, lines, lineterm='', n=0))[:120]:
    print(l)
