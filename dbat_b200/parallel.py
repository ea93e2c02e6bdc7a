"""Multi-GPU host plumbing: one process per GPU, points (and their observations) partitioned
by object-point index, reduced camera system combined by an NCCL allreduce inside
libdbatgpu.so.  torch.distributed is used only for the rendezvous: it
carries the 128-byte NCCL unique id from rank 0 to the other ranks.
"""
import ctypes as C

import numpy as np

from . import _lib
from .bundle import Problem


def partition_points(op_of_obs, nOP, world):
    """Contiguous object-point ranges [lo,hi) balanced by observation count."""
    cnt = np.bincount(np.asarray(op_of_obs), minlength=nOP)
    cum = np.concatenate([[0], np.cumsum(cnt)])
    total = cum[-1]
    bounds = [0]
    for r in range(1, world):
        bounds.append(int(np.searchsorted(cum, total * r / world, side='left')))
    bounds.append(nOP)
    bounds = np.maximum.accumulate(np.minimum(bounds, nOP))
    return [(int(bounds[r]), int(bounds[r + 1])) for r in range(world)]


def exchange_unique_id(rank, make_id, dist=None):
    """Rank 0 creates the 128-byte id (make_id()), every rank receives it."""
    import torch
    if dist is None:
        import torch.distributed as dist
    buf = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        buf[:] = torch.frombuffer(bytearray(make_id()), dtype=torch.uint8)
    dev = None
    if dist.get_backend() == 'nccl':
        dev = torch.device('cuda', torch.cuda.current_device())
        buf = buf.to(dev)
    dist.broadcast(buf, src=0)
    return bytes(buf.cpu().numpy().tobytes())


def gather_unique_keys(keys, dist=None):
    """Sorted union over all ranks of an int64 key array (variable length per rank)."""
    import torch
    if dist is None:
        import torch.distributed as dist
    world = dist.get_world_size()
    dev = torch.device('cuda', torch.cuda.current_device()) if dist.get_backend() == 'nccl' else torch.device('cpu')
    keys = np.unique(np.asarray(keys, dtype=np.int64))
    cnt = torch.tensor([len(keys)], dtype=torch.int64, device=dev)
    cnts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(cnts, cnt)
    m = int(max(int(c.item()) for c in cnts))
    buf = torch.full((max(1, m),), -1, dtype=torch.int64, device=dev)
    buf[:len(keys)] = torch.from_numpy(keys).to(dev)
    bufs = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(bufs, buf)
    allk = torch.cat(bufs).cpu().numpy()
    return np.unique(allk[allk >= 0])


class ShardedProblem(Problem):
    """Rank-local shard of a point-partitioned problem; x/p vectors keep the global layout.

    Entries of p (and of the iterate) that belong to other ranks' points are zero/stale on this
    rank; `gather` merges them.
    """

    def __init__(self, s, rank, world, local=False):
        """local=False: `s` is the whole project, this rank takes its share of the object points.
        local=True: `s` already holds only this rank's object points and observations (cameras and IO are
        the same on every rank; x is [IO; EO; this rank's OP]): nothing of the whole project ever has to
        exist on one host.  The co-visibility graph is then gathered over the ranks."""
        import torch.distributed as dist
        self.rank, self.world, self.local = rank, world, local
        if local:
            from .bundle import covisibility_edges
            nImg, nOP = s.EO.val.shape[1], s.OP.val.shape[1]
            ca, cb = covisibility_edges(s.IP.img, s.IP.op, nImg, nOP)
            key = gather_unique_keys(ca * nImg + cb, dist)
            super().__init__(s, cam_priors=(rank == 0), covis=(key // nImg, key % nImg))
            parts = None
        else:
            parts = partition_points(s.IP.op, s.OP.val.shape[1], world)
            super().__init__(s, points=parts[rank], cam_priors=(rank == 0))
        self.parts = parts
        L = _lib.lib()

        def make_id():
            b = C.create_string_buffer(128)
            rc = L.dbat_comm_unique_id(b)
            if rc != 0:
                raise _lib.DbatError(rc, 'ncclGetUniqueId failed')
            return b.raw

        uid = exchange_unique_id(rank, make_id, dist)
        self._check(L.dbat_comm_init(self._h, world, rank, C.create_string_buffer(uid, 128)))
        des = s.bundle.deserial.OP
        if local:
            self.own_cols = des.src
        else:
            lo, hi = parts[rank]
            self.own_cols = des.src[(des.dest >= 3 * lo) & (des.dest < 3 * hi)]

    def gather(self, v):
        """Merge a per-rank vector in x layout: camera part from any rank, point part from owners."""
        import torch
        import torch.distributed as dist
        nC = self.n - len(self.s.bundle.serial.OP.dest)
        t = torch.zeros(self.n, dtype=torch.float64)
        t[torch.from_numpy(self.own_cols)] = torch.from_numpy(np.asarray(v)[self.own_cols])
        if self.rank == 0:
            t[:nC] = torch.from_numpy(np.asarray(v)[:nC])
        if dist.get_backend() == 'nccl':
            t = t.cuda()
        dist.all_reduce(t)
        return t.cpu().numpy()
