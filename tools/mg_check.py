"""2+ rank parity check of the point-sharded path against the single-GPU path (torchrun)."""
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
s, _ = make_scene(40, 3000, rays=6, seed=5)
s.prior.EO.use[0:3, 3] = True; s.prior.EO.val[0:3, 3] = s.EO.val[0:3, 3]; s.prior.EO.std[0:3, 3] = 0.05
s.prior.OP.use[:, 7] = True; s.prior.OP.val[:, 7] = s.OP.val[:, 7]; s.prior.OP.std[:, 7] = 0.02
s.bundle.serial = None
dbat_b200.buildserialindices(s)
x0 = dbat_b200.serialize(s)
P = ShardedProblem(s, rank, world)
for lam, jac in ((0.0, False), (1e3, False), (0.0, True)):
    p, st = P.normal_step(x0, lam, jac, trial=True)
    pg = P.gather(p)
    if rank == 0:
        P1 = dbat_b200.Problem(s)
        p1, st1 = P1.normal_step(x0, lam, jac, trial=True)
        P1.close()
        print('lam=%g jac=%d: step rel diff %.2e  f %.15g vs %.15g  f_new %.15g vs %.15g  jp2 rel %.2e' % (
            lam, jac, np.abs(pg - p1).max() / np.abs(p1).max(), st['f'], st1['f'], st['f_new'], st1['f_new'],
            abs(st['jp2'] - st1['jp2']) / st1['jp2']), flush=True)
# full LM solve on the shards vs single GPU
o = P.solve('lm', x0, want_trace=False, want_resid=False)
xg = P.gather(o.x)
if rank == 0:
    P1 = dbat_b200.Problem(s)
    o1 = P1.solve('lm', x0, want_trace=False, want_resid=False)
    print('LM: sharded n=%d code=%d rr=%s | single n=%d code=%d rr=%s | x rel diff %.2e' % (
        o.n, o.code, np.array2string(o.rr, precision=6), o1.n, o1.code, np.array2string(o1.rr, precision=6),
        np.abs(xg - o1.x).max() / np.abs(o1.x).max()), flush=True)
dist.barrier()
dist.destroy_process_group()
