"""Multi-rank parity worker (run under torchrun, one rank per GPU): the point-sharded path against the
single-GPU path on the same problem.  Exits non-zero on a mismatch; tests/test_multi_gpu.py drives it."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dbat_b200
from dbat_b200.parallel import ShardedProblem
from dbat_b200.synth import make_scene

rank, world, lr = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(lr)
dist.init_process_group('nccl', device_id=torch.device('cuda', lr))
nImg, nOP = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (40, 3000)
s, _ = make_scene(nImg, nOP, rays=6, seed=5)
s.prior.EO.use[0:3, 3] = True; s.prior.EO.val[0:3, 3] = s.EO.val[0:3, 3]; s.prior.EO.std[0:3, 3] = 0.05
s.prior.OP.use[:, 7] = True; s.prior.OP.val[:, 7] = s.OP.val[:, 7]; s.prior.OP.std[:, 7] = 0.02
s.bundle.serial = None
dbat_b200.buildserialindices(s)
x0 = dbat_b200.serialize(s)
P = ShardedProblem(s, rank, world)
P1 = dbat_b200.Problem(s) if rank == 0 else None
bad = []


def check(name, got, ref, tol):
    err = float(np.abs(np.asarray(got) - np.asarray(ref)).max() / max(np.abs(ref).max(), 1e-300))
    print('%-28s rel diff %.2e' % (name, err), flush=True)
    if not err <= tol:
        bad.append(name)


for lam, jac in ((0.0, False), (1e3, False), (0.0, True)):
    p, st = P.normal_step(x0, lam, jac, trial=True)
    pg = P.gather(p)
    if rank == 0:
        p1, st1 = P1.normal_step(x0, lam, jac, trial=True)
        check('step lam=%g jac=%d' % (lam, jac), pg, p1, 1e-9)
        check('  f, f_new, jp2', [st['f'], st['f_new'], st['jp2']], [st1['f'], st1['f_new'], st1['jp2']], 1e-11)
for method in ('lm', 'gna', 'lmp'):
    o = P.solve(method, x0, want_trace=False, want_resid=False)
    xg = P.gather(o.x)
    if rank == 0:
        o1 = P1.solve(method, x0, want_trace=False, want_resid=False)
        print('%s: sharded n=%d code=%d | single n=%d code=%d' % (method, o.n, o.code, o1.n, o1.code), flush=True)
        if (o.n, o.code) != (o1.n, o1.code):
            bad.append(method + ' iteration count')
        check(method + ' x', xg, o1.x, 1e-9)
        check(method + ' rr', o.rr, o1.rr, 1e-10)
flag = torch.tensor([len(bad)], device='cuda')
dist.broadcast(flag, src=0)
dist.barrier()
dist.destroy_process_group()
if rank == 0 and bad:
    print('MISMATCH:', bad, flush=True)
sys.exit(1 if int(flag.item()) else 0)
