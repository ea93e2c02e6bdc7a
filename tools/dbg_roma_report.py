"""Debug helper (GPU): diff of the device's roma result file against the reference's."""
import os, sys, difflib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import __graft_entry__ as g
g.build()
import dbat_b200
from dbat_b200.report import bundle_result_file
from test_report_golden import roma_struct, GOLD
s = roma_struct()
s, ids, _ = dbat_b200.forwintersect(s, 'all', True)
dbat_b200.seteoest_depend(s, 0)
s, ok, it, s0, E = dbat_b200.bundle(s, 'gna')
for w in ('CIO', 'CEO', 'COP'):
    Cg = dbat_b200.bundle_cov(s, E, w)
    d = Cg.diagonal()
    print(w, 'nan', int(np.isnan(d).sum()), 'neg', int((d < 0).sum()), 'max', np.nanmax(d))
s, lines = bundle_result_file(s, E)
gold = [l.rstrip('\n') for l in open(os.path.join(GOLD, 'romabundledemo', 'result', 'report.txt'))]
for l in list(difflib.unified_diff(gold, lines, lineterm='', n=0))[:80]:
    print(l)
