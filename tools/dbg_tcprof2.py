"""Debug (GPU): per-task time stamps of the tile factorisation of the config-4 reduced system."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ['DBAT_TCHOL_PROF'] = os.path.join(ROOT, 'gpurun_out', 'tcprof_c4.csv')
import numpy as np
import __graft_entry__ as g
g.build()
import dbat_b200
from dbat_b200.synth import make_scene
s, _ = make_scene(1000, 200000, rays=10, cache_dir='/tmp')
P = dbat_b200.Problem(s)
x0 = dbat_b200.serialize(s)
for _ in range(4):
    P.normal_step(x0, 0.0, trial=True, want_p=False)
import pandas as pd
d = pd.read_csv(os.environ['DBAT_TCHOL_PROF'])
print('tasks', len(d), 'total span us', d.t_end.max() / 1e3)
diag = d[(d.I == d.J) & (d['mode'] == 0)].sort_values('t_end')
print(diag[['I', 'J', 'level', 'nterms', 't_claim', 't_terms', 't_deps', 't_end']].iloc[::6].to_string())
lv = d[d['mode'] == 0].groupby('level').t_end.max()
print('level end times (us):', [round(v / 1e3, 1) for v in lv.values])
print('potrf mean', (diag.t_end - diag.t_deps).mean(), 'diag wait after terms', (diag.t_deps - diag.t_terms).mean())
tl = d[(d.I != d.J) & (d['mode'] == 0)]
print('tile solve mean', (tl.t_end - tl.t_deps).mean(), 'tile deps wait mean', (tl.t_deps - tl.t_terms).mean())
print('claim->terms mean per term (ns)', ((d.t_terms - d.t_claim) / d.nterms.clip(lower=1)).mean())
# the dependency chain of the last separator: diagonal tasks of the highest levels, and the sub-diagonal tile between them
last = diag[diag.level >= diag.level.max() - 12].sort_values('J')
prev_end = None
for _, r in last.iterrows():
    sub = d[(d.I == r.J + 1) & (d.J == r.J) & (d['mode'] == 0)]
    s_txt = ''
    if len(sub):
        q = sub.iloc[0]
        s_txt = 'sub(%d,%d): terms %.1f deps %.1f end %.1f' % (q.I, q.J, q.t_terms / 1e3, q.t_deps / 1e3, q.t_end / 1e3)
    print('J=%d terms %.1f potrf %.1f..%.1f (%.1f) step %s | %s' % (r.J, r.t_terms / 1e3, r.t_deps / 1e3, r.t_end / 1e3, (r.t_end - r.t_deps) / 1e3,
          '' if prev_end is None else '%.1f' % ((r.t_end - prev_end) / 1e3), s_txt))
    prev_end = r.t_end
