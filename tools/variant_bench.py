"""A/B of tuning builds of the library on the config-4 block (GPU box):

    python tools/variant_bench.py base=variants/libdbatgpu_base.so new=dbat_b200/libdbatgpu.so dmma=variants/libdbatgpu_dmma.so[,ENV=VAL]

Everything runs in one process (the scene is generated once).  Every variant takes one damped step from the same start vector (compared with the first variant's step: the
parity column), then a short LM walk whose accepted iterates are compared as well, then the timed loop of
bench.py (device-resident, accepted trial points).  One JSON line per variant under gpurun_out/variants.jsonl.
"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = '/tmp/variant_ref.npz'


def run_variant(name, s, x0, steps, warmup):
    import numpy as np
    import dbat_b200
    P = dbat_b200.Problem(s)
    p, st = P.normal_step(x0, 0.0, trial=True, accept=False)
    # a short walk: three accepted LM steps, the objective after each
    P.normal_step(x0, 0.0, trial=True, accept=False, want_p=False)
    fs = []
    for _ in range(3):
        _, s3 = P.normal_step(None, 0.0, trial=True, accept=True, want_p=False)
        fs.append(s3['f_new'])
    out = dict(name=name, f=st['f'], f_new=st['f_new'], jp2=st['jp2'], rjp=st['rjp'], singular=st['singular'], walk_f=fs)
    if os.path.exists(REF):
        r = np.load(REF)
        sc = np.maximum(np.abs(r['p']), 1e-12 * np.abs(r['p']).max())
        out['p_relmax_vs_first'] = float(np.max(np.abs(p - r['p']) / sc))
        out['p_rel2_vs_first'] = float(np.linalg.norm(p - r['p']) / np.linalg.norm(r['p']))
        out['f_rel'] = abs(st['f'] - float(r['f'])) / float(r['f'])
        out['fnew_rel'] = abs(st['f_new'] - float(r['f_new'])) / float(r['f_new'])
        out['jp2_rel'] = abs(st['jp2'] - float(r['jp2'])) / abs(float(r['jp2']))
        out['walk_rel'] = float(np.max(np.abs(np.array(fs) - r['walk']) / r['walk']))
    else:
        np.savez(REF, p=p, f=st['f'], f_new=st['f_new'], jp2=st['jp2'], walk=np.array(fs))
    best = None
    for rep in range(2):
        P.normal_step(x0, 0.0, trial=True, accept=False, want_p=False)
        for _ in range(warmup):
            P.normal_step(None, 0.0, trial=True, accept=True, want_p=False)
        P.normal_step(x0, 0.0, trial=True, accept=False, want_p=False)
        ms = []
        t0 = time.perf_counter()
        for _ in range(steps):
            _, s2 = P.normal_step(None, 0.0, trial=True, accept=True, want_p=False)
            ms.append(s2['device_ms'])
        wall = (time.perf_counter() - t0) / steps * 1e3
        cur = dict(ms_per_step=float(np.mean(ms)), ms_min=float(np.min(ms)), wall_ms=wall, launches=s2['launches'],
                   phases={k: round(v[0], 4) for k, v in P.phase_times().items()})
        if best is None or cur['ms_per_step'] < best['ms_per_step']:
            best = cur
    out.update(best)
    P.close()
    return out


def main():
    """All variants in one process (the scene is generated once): every variant is its own copy of the library file,
    loaded after the previous problem has been destroyed; ENV=VAL pairs are set before its first call."""
    import shutil
    import dbat_b200
    from dbat_b200 import _lib
    from dbat_b200.synth import make_scene
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    if os.path.exists(REF):
        os.remove(REF)
    nImg, nOP = int(os.environ.get('VB_NIMG', 1000)), int(os.environ.get('VB_NOP', 200000))
    s, _ = make_scene(nImg, nOP, rays=10, cache_dir='/tmp')
    x0 = dbat_b200.serialize(s)
    lines = []
    for spec in sys.argv[1:]:
        name, rest = spec.split('=', 1)
        parts = rest.split(',')
        path = '/tmp/vb_%s.so' % name
        shutil.copyfile(os.path.join(ROOT, parts[0]), path)
        saved = {}
        for kv in parts[1:]:
            k, v = kv.split('=', 1)
            saved[k] = os.environ.get(k)
            os.environ[k] = v
        _lib._lib = None
        _lib.LIB_PATH = path
        try:
            d = run_variant(name, s, x0, 20, 3)
            lines.append(json.dumps(d))
            print('%-10s %.4f ms (min %.4f, wall %.4f) launches %d  dp %.2e walk %.2e  %s' % (
                name, d['ms_per_step'], d['ms_min'], d['wall_ms'], d['launches'], d.get('p_relmax_vs_first', 0.0),
                d.get('walk_rel', 0.0), d['phases']), flush=True)
        except Exception as e:                         # a failing variant must not hide the others
            print(name, 'FAILED', repr(e), flush=True)
            lines.append(json.dumps(dict(name=name, failed=True, err=repr(e))))
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    with open(os.path.join(ROOT, 'gpurun_out', 'variants.jsonl'), 'a') as f:
        for l in lines:
            f.write(l + '\n')


if __name__ == '__main__':
    main()
