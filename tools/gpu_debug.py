"""Stage-by-stage CUDA-vs-oracle diagnostics (run on the GPU box through gpurun)."""
import copy
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import dbat_b200
from dbat_b200.synth import make_scene
from oracle.bundle import bundle as obundle, bundle_cov as ocov
from oracle.cameramodel import brown_euler_cam4
from oracle.dbatstruct import buildweightmatrix, serialize, buildserialindices


def relerr(a, b):
    return float(np.max(np.abs(a - b)) / max(1e-300, np.max(np.abs(b))))


def stage(name, fn):
    t = time.time()
    try:
        fn()
        print('[ok ] %s (%.2fs)' % (name, time.time() - t), flush=True)
    except Exception:
        print('[ERR] %s' % name, flush=True)
        traceback.print_exc()


def run_problem(tag, s):
    print('=== %s: nImg=%d nOP=%d nObs=%d' % (tag, s.EO.val.shape[1], s.OP.val.shape[1], len(s.IP.img)), flush=True)
    if s.bundle.serial is None:
        buildserialindices(s)
    x0 = serialize(s)
    W = buildweightmatrix(s)
    P = dbat_b200.Problem(copy.deepcopy(s))
    ctx = {}

    def st_resid():
        r = P(x0)
        ro, Jo = brown_euler_cam4(x0, s, True)
        ctx['ro'], ctx['Jo'] = ro, Jo
        print('   residual rel err %.3e  (|r|=%.6g)' % (relerr(r, ro), np.linalg.norm(ro)))

    def st_jac():
        J = P.jacobian(False)
        Jo = ctx['Jo']
        same = (J.shape == Jo.shape and np.array_equal(J.indptr, Jo.indptr) and np.array_equal(J.indices, Jo.indices))
        print('   J nnz %d vs %d, pattern identical: %s' % (J.nnz, Jo.nnz, same))
        if same:
            print('   J value rel err %.3e' % relerr(J.data, Jo.data))
        else:
            d = abs(J - Jo)
            print('   max |J-Jo| = %.3e' % d.max())

    def st_step():
        for lam, jac in ((0.0, False), (1e3, False), (0.0, True)):
            p, st = P.normal_step(x0, lam, jac)
            Jw = (ctx['Jo'].multiply(np.sqrt(W)[:, None])).tocsc()
            rw = ctx['ro'] * np.sqrt(W)
            N = (Jw.T @ Jw).toarray()
            po = np.linalg.solve(N + lam * np.eye(N.shape[0]), -(Jw.T @ rw))
            print('   lam=%g jacobi=%d: step rel err %.3e, f rel err %.3e, |Jp|^2 rel err %.3e, launches %d'
                  % (lam, jac, relerr(p, po), abs(st['f'] - 0.5 * rw @ rw) / (0.5 * rw @ rw),
                     abs(st['jp2'] - np.sum((Jw @ po) ** 2)) / np.sum((Jw @ po) ** 2), st['launches']))
        print('   phases', P.phase_times())

    def st_solve():
        for damp in ('gna', 'lm', 'lmp'):
            s1, s2 = copy.deepcopy(s), copy.deepcopy(s)
            t = time.time()
            s1, ok, it, s0, E = dbat_b200.bundle(s1, damp)
            tg = time.time() - t
            t = time.time()
            s2, oko, ito, s0o, Eo = obundle(s2, damp)
            to = time.time() - t
            print('   %s: cuda code=%d iters=%d s0=%.12g (%.3fs, gpu %.4fs, %d launches) | oracle code=%d iters=%d s0=%.12g (%.3fs)'
                  % (damp, E.code, it, s0, tg, E.gpu_seconds, E.problem.last.launches, Eo.code, ito, s0o, to))
            print('      x rel err %.3e, rr %s vs %s' % (relerr(E.x, Eo.x), np.array2string(E.res, precision=6), np.array2string(Eo.res, precision=6)))
            ctx[damp] = (s1, E, s2, Eo)

    def st_cov():
        s1, E, s2, Eo = ctx['gna']
        for w in ('CEO', 'CIO', 'COP'):
            Cg = dbat_b200.bundle_cov(s1, E, w).toarray()
            Co = ocov(s2, Eo, w)
            print('   %s rel err %.3e' % (w, relerr(Cg, Co)))

    stage('residual', st_resid)
    stage('jacobian', st_jac)
    stage('normal step', st_step)
    stage('solve', st_solve)
    stage('cov', st_cov)
    P.close()


if __name__ == '__main__':
    which = sys.argv[1:] or ['synth', 'camcal']
    if 'synth' in which:
        s, _ = make_scene(21, 100, rays=10, seed=7)
        run_problem('synth21', s)
    if 'camcal' in which:
        from camcal_fixture import camcal_struct
        run_problem('camcal', camcal_struct('default', seed=1))
    if 'mid' in which:
        s, _ = make_scene(60, 4000, rays=10, seed=3)
        run_problem('synth60', s)
