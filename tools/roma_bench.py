"""The reference's roma script project (data/script/romabundledemo, BASELINE.md row "roma via XML script":
MATLAB 7.28 s for the 5 GNA iterations, post-cov 0.69 + 0.01 + 0.03 + 1.12 s) on the device, with the oracle
timed beside it (GPU box).  Usage: python tools/roma_bench.py [--no-oracle]"""
import copy, json, os, sys, time
HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, 'tests'))
import numpy as np
import dbat_b200
from dbat_b200.report import bundle_result_file
from oracle.loaders import load_roma_script
from test_report_golden import report_diff

root = os.path.join(HERE, 'tests', 'golden', 'romabundledemo')
s0_ = load_roma_script(root)
out = {'workload': 'roma script project: 60 images, 26321 points, 90561 image points, n=79321',
       'matlab_published_s': {'bundle_5_iterations': 7.28, 'post_cov': 0.69 + 0.01 + 0.03 + 1.12}}
for rep in range(3):                                        # first pass pays library load and graph capture
    s = copy.deepcopy(s0_)
    t = time.time(); s, ids, _ = dbat_b200.forwintersect(s, 'all', True); t_fwi = time.time() - t
    dbat_b200.seteoest_depend(s, 0)
    t = time.time(); s, ok, it, sig0, E = dbat_b200.bundle(s, 'gna'); t_bundle = time.time() - t
    t = time.time(); s, lines = bundle_result_file(s, E); t_report = time.time() - t
out['device_s'] = {'forwintersect': t_fwi, 'bundle': t_bundle, 'cov_and_report': t_report,
                   'iterations': int(it), 'ok': bool(ok), 'sigma0': float(sig0)}
out['report_lines_differing_at_1e-5'] = len(report_diff(lines, os.path.join(root, 'result', 'report.txt'), rtol=1e-5))
if '--no-oracle' not in sys.argv:
    from oracle.photogrammetry import forwintersect as ofwi
    from oracle.dbatstruct import seteoest_depend
    from oracle.bundle import bundle as obundle
    s = copy.deepcopy(s0_)
    s, _, _ = ofwi(s, 'all', True)
    seteoest_depend(s, 0)
    t = time.time(); s, ok, it, sig0, Eo = obundle(s, 'gna'); out['oracle_bundle_s'] = time.time() - t
print(json.dumps(out))
