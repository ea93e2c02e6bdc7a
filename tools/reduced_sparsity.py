"""Block sparsity of the reduced camera system S (SURVEY §8(f) N4 groundwork), host only.

S = U - W V^-1 W' has a non-zero 6x6 block (i, j) exactly when images i and j see a common object point.
This tool builds that co-visibility pattern for a project, orders it (natural order, reverse Cuthill-McKee,
and a minimum-degree ordering), runs a symbolic block Cholesky and reports fill and factorisation work
next to the dense factorisation the device path does today.  The shared-IO rows are a dense border of 9
rows under any ordering and are counted separately.

    python tools/reduced_sparsity.py [nImg nOP | roma]
"""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import scipy.sparse as sp
from scipy.sparse.csgraph import reverse_cuthill_mckee


def covisibility(img, op, nImg):
    """nImg x nImg boolean pattern (with diagonal) of images that share an object point."""
    A = sp.csr_matrix((np.ones(len(img), np.int32), (op, img)))
    G = (A.T @ A).tocsr()
    G.data[:] = 1
    return G.astype(bool)


def symbolic_cholesky(G, perm):
    """Column counts of the Cholesky factor of the permuted pattern (elimination by merging adjacency sets)."""
    n = G.shape[0]
    P = G[perm][:, perm].tocsc()
    adj = [set(P.indices[P.indptr[j]:P.indptr[j + 1]][P.indices[P.indptr[j]:P.indptr[j + 1]] > j]) for j in range(n)]
    counts = np.zeros(n, np.int64)
    for j in range(n):
        rows = adj[j]
        counts[j] = len(rows)
        if rows:
            p = min(rows)                        # parent in the elimination tree inherits the rest
            adj[p] |= rows - {p}
    return counts


def minimum_degree(G):
    """Plain minimum-degree ordering on the quotient-free graph (fine for a few thousand nodes)."""
    n = G.shape[0]
    C = G.tocsr()
    nb = [set(C.indices[C.indptr[i]:C.indptr[i + 1]]) - {i} for i in range(n)]
    alive = np.ones(n, bool)
    deg = np.array([len(s) for s in nb])
    order = []
    for _ in range(n):
        i = int(np.flatnonzero(alive)[np.argmin(deg[alive])])
        order.append(i)
        alive[i] = False
        ns = nb[i]
        for a in ns:
            nb[a].discard(i)
            nb[a] |= ns - {a}
            deg[a] = len(nb[a])
        nb[i] = set()
    return np.array(order)


def work(counts, b=6):
    """Flops of a block right-looking Cholesky with b x b blocks: per column c sub-diagonal blocks ->
    potrf b^3/3 + trsm c b^3 + update c(c+1)/2 * 2 b^3."""
    c = counts.astype(float)
    return float(np.sum(b ** 3 / 3 + c * b ** 3 + c * (c + 1) * b ** 3))


def main():
    if len(sys.argv) > 1 and sys.argv[1] == 'roma':
        sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
        from oracle.loaders import load_roma_script
        s = load_roma_script(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                          'tests', 'golden', 'romabundledemo'))
        name = 'roma script project'
    else:
        from dbat_b200.synth import make_scene
        nImg, nOP = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1000, 200000)
        s, _ = make_scene(nImg, nOP, rays=10, seed=20240607, build_indices=False)
        name = 'synthetic block %d x %d (BASELINE config generator)' % (nImg, nOP)
    nImg = s.EO.val.shape[1]
    G = covisibility(np.asarray(s.IP.img), np.asarray(s.IP.op), nImg)
    nnzb = G.nnz
    dense_flops = (6.0 * nImg) ** 3 / 3
    out = {'project': name, 'images': nImg, 'observations': int(s.IP.val.shape[1]),
           'S_block_density': round(nnzb / nImg ** 2, 4), 'dense_cholesky_gflop': round(dense_flops / 1e9, 3)}
    coo = G.tocoo()
    for label, perm in (('natural', np.arange(nImg)), ('rcm', np.asarray(reverse_cuthill_mckee(G.tocsr(), symmetric_mode=True))),
                        ('minimum_degree', minimum_degree(G))):
        cnt = symbolic_cholesky(G, perm)
        inv = np.empty(nImg, np.int64)
        inv[perm] = np.arange(nImg)
        band = int(np.abs(inv[coo.row] - inv[coo.col]).max())
        out[label] = {'bandwidth_blocks': band, 'band_cholesky_gflop': round(6.0 * nImg * (6.0 * band) ** 2 / 1e9, 3),
                      'L_block_density': round(float((cnt.sum() + nImg) / (nImg * (nImg + 1) / 2)), 4),
                      'gflop': round(work(cnt) / 1e9, 3), 'vs_dense': round(work(cnt) / dense_flops, 4),
                      'largest_front_blocks': int(cnt.max())}
    print(json.dumps(out))


if __name__ == '__main__':
    main()
