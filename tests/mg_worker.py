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
# posterior covariances on the shards: camera blocks identical on every rank, point blocks for the rank's own points
o = P.solve('gna', x0, want_trace=False, want_resid=False)
s0 = 1.3
ceo, cop = P.cov('ceo', s0), P.cov('cop', s0)
lo, hi = P.parts[rank]
t = torch.zeros(s.OP.val.shape[1], 3, 3, dtype=torch.float64, device='cuda')
t[lo:hi] = torch.from_numpy(cop[:hi - lo]).cuda()
dist.all_reduce(t)
if rank == 0:
    P1.solve('gna', x0, want_trace=False, want_resid=False)
    check('CEO', ceo, P1.cov('ceo', s0), 1e-8)
    check('COP', t.cpu().numpy(), P1.cov('cop', s0), 1e-8)
    # Jacobian of the shard: the rows of this rank's observations out of the single-GPU Jacobian
    Js, J1 = P.jacobian(True), P1.jacobian(True)
    sel = np.flatnonzero((s.IP.op >= lo) & (s.IP.op < hi))
    rows = np.stack([2 * sel, 2 * sel + 1], axis=1).ravel()
    check('J rows of the shard', Js[:2 * len(sel)].toarray(), J1[rows].toarray(), 1e-12)
flag = torch.tensor([len(bad)], device='cuda')
dist.broadcast(flag, src=0)
dist.barrier()
dist.destroy_process_group()
if rank == 0 and bad:
    print('MISMATCH:', bad, flush=True)
sys.exit(1 if int(flag.item()) else 0)
