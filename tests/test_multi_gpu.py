"""The point-sharded multi-GPU path against the single-GPU path, on real devices (NCCL, one rank per GPU).
Needs two GPUs: `gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu`."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize('nImg,nOP', [(40, 3000), (420, 12000)])
def test_sharded_solvers_match_single_gpu(built_lib, nImg, nOP):
    """Steps, f, |Jp|^2 and complete LM / GNA / LMP runs of a 2-rank point-sharded problem equal the
    single-GPU run (1e-9; identical iteration counts).  The larger case has a dissected, tiled reduced system."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
           '--master-addr', '127.0.0.1', '--master-port', '29517', os.path.join(ROOT, 'tests', 'mg_worker.py'),
           str(nImg), str(nOP)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    sys.stdout.write(out.stdout[-4000:])
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]


@pytest.mark.gpu
def test_set_devices_one_process_two_gpus(built_lib):
    """dbat_set_devices: ONE process - as a MEX gateway would be - drives two devices through the plain
    C ABI; evaluation, a damped step, complete LM / GNA runs and the covariance blocks equal the single-device
    results (1e-9; identical iteration counts), in the single-device layout of x, p and the residual."""
    import copy
    import numpy as np
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs')
    import dbat_b200
    from dbat_b200.synth import make_scene
    s, _ = make_scene(60, 4000, rays=6, seed=5)
    s.prior.EO.use[0:3, 3] = True; s.prior.EO.val[0:3, 3] = s.EO.val[0:3, 3]; s.prior.EO.std[0:3, 3] = 0.05
    s.prior.OP.use[:, 7] = True; s.prior.OP.val[:, 7] = s.OP.val[:, 7]; s.prior.OP.std[:, 7] = 0.02
    s.prior.OP.use[:, 3990] = True; s.prior.OP.val[:, 3990] = s.OP.val[:, 3990]; s.prior.OP.std[:, 3990] = 0.02
    s.bundle.serial = None
    dbat_b200.buildserialindices(s)
    x0 = dbat_b200.serialize(s)
    P1 = dbat_b200.Problem(copy.deepcopy(s))
    for devs in ([0, 1], [1]):
        P2 = dbat_b200.Problem(copy.deepcopy(s)).set_devices(devs)

        def close(a, b, tol, what):
            err = float(np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(b).max(), 1e-300))
            assert err <= tol, (devs, what, err)

        close(P2(x0), P1(x0), 1e-12, 'residual')
        close(P2(x0, weighted=True), P1(x0, weighted=True), 1e-12, 'weighted residual')
        for lam, jac in ((0.0, False), (1e3, False), (0.0, True)):
            p2, st2 = P2.normal_step(x0, lam, jac, trial=True)
            p1, st1 = P1.normal_step(x0, lam, jac, trial=True)
            close(p2, p1, 1e-9, 'step')
            close([st2['f'], st2['f_new'], st2['jp2']], [st1['f'], st1['f_new'], st1['jp2']], 1e-11, 'f, f_new, jp2')
        for method in ('lm', 'gna', 'lmp'):
            o2, o1 = P2.solve(method, x0), P1.solve(method, x0)
            assert (o2.n, o2.code) == (o1.n, o1.code), (devs, method)
            close(o2.x, o1.x, 1e-9, method + ' x')
            close(o2.rr, o1.rr, 1e-10, method + ' rr')
            close(o2.r_w, o1.r_w, 1e-9, method + ' weighted residual')
            close(o2.T, o1.T, 1e-8, method + ' trace')
        for w in ('cio', 'ceo', 'cop'):
            close(P2.cov(w, 1.3), P1.cov(w, 1.3), 1e-8, w)
        with pytest.raises(dbat_b200._lib.DbatError):
            P2.jacobian()
        P2.close()
    P1.close()
