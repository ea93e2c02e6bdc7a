"""Host prototype of the banded tile Cholesky proposed in DESIGN.md §9 (no GPU code; a verified design aid).

The reduced camera system, with the images in reverse Cuthill-McKee order and the shared IO block (and the
rhs row) moved to the end, is a band matrix with a dense border.  The device factorisation is a right-looking
tile algorithm (NB = 128): potrf of the diagonal tile, panel solve below it, trailing update.  `band_tile_cholesky`
runs the same loop but only over the tiles that can be non-zero - the tiles within `bt` tile rows of the
diagonal and the border tile rows - and counts the tile operations, i.e. the launches' work.

    python tools/band_cholesky_proto.py [nImg nOP]      # tile-operation counts for a synthetic block
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def band_tile_cholesky(A, NB, bt, nborder):
    """In-place lower Cholesky of SPD A (n x n, n a multiple of NB) whose leading part has tile bandwidth bt
    and whose last nborder tile rows are dense.  Returns (L, counts)."""
    n = A.shape[0]
    nb = n // NB
    ne = nb - nborder                                   # tile rows of the banded part
    L = np.tril(A).copy()
    T = lambda i, j: (slice(i * NB, (i + 1) * NB), slice(j * NB, (j + 1) * NB))
    cnt = {'potrf': 0, 'trsm_tiles': 0, 'gemm_tiles': 0, 'steps': nb}
    for k in range(nb):
        L[T(k, k)] = np.linalg.cholesky(L[T(k, k)] + np.triu(L[T(k, k)], 1).T * 0)
        cnt['potrf'] += 1
        rows = list(range(k + 1, min(k + bt, ne - 1) + 1)) if k < ne else []
        rows += list(range(max(k + 1, ne), nb))          # border tile rows are always touched
        inv = np.linalg.inv(L[T(k, k)])
        for i in rows:                                    # panel solve: L_ik = A_ik inv(L_kk)'
            L[T(i, k)] = L[T(i, k)] @ inv.T
            cnt['trsm_tiles'] += 1
        for a, i in enumerate(rows):                      # trailing update, lower triangle of the touched tiles
            for j in rows[:a + 1]:
                L[T(i, j)] -= L[T(i, k)] @ L[T(j, k)].T
                cnt['gemm_tiles'] += 1
    for k in range(nb):
        L[T(k, k)] = np.tril(L[T(k, k)])
    return L, cnt


def dense_counts(nb):
    return {'potrf': nb, 'trsm_tiles': nb * (nb - 1) // 2, 'gemm_tiles': sum((nb - k - 1) * (nb - k) // 2 for k in range(nb)), 'steps': nb}


def main():
    from dbat_b200.synth import make_scene
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from reduced_sparsity import covisibility
    from scipy.sparse.csgraph import reverse_cuthill_mckee
    nImg, nOP = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1000, 200000)
    s, _ = make_scene(nImg, nOP, rays=10, seed=20240607, build_indices=False)
    G = covisibility(np.asarray(s.IP.img), np.asarray(s.IP.op), nImg)
    perm = np.asarray(reverse_cuthill_mckee(G.tocsr(), symmetric_mode=True))
    inv = np.empty(nImg, int)
    inv[perm] = np.arange(nImg)
    coo = G.tocoo()
    NB = 128
    # tile bandwidth: the farthest tile row any column's band reaches (6 columns per image, 7 EO elements fixed)
    col_lo = 6 * np.minimum(inv[coo.row], inv[coo.col])
    col_hi = 6 * np.maximum(inv[coo.row], inv[coo.col]) + 5
    bt = int(((col_hi // NB) - (col_lo // NB)).max())
    nC = 6 * nImg - 7 + 9
    nb = (nC + 1 + NB - 1) // NB                           # + rhs row, padded to tiles
    ne = (6 * nImg - 7) // NB
    nborder = nb - ne
    # count only (same loop as band_tile_cholesky, no arithmetic)
    cnt = {'potrf': 0, 'trsm_tiles': 0, 'gemm_tiles': 0, 'steps': nb}
    for k in range(nb):
        rows = (min(k + bt, ne - 1) - k if k < ne else 0)
        rows = max(rows, 0) + (nb - max(k + 1, ne))
        cnt['potrf'] += 1
        cnt['trsm_tiles'] += rows
        cnt['gemm_tiles'] += rows * (rows + 1) // 2
    d = dense_counts(nb)
    flop = lambda c: (c['potrf'] / 3 + c['trsm_tiles'] + 2 * c['gemm_tiles']) * NB ** 3
    print(json.dumps({'workload': '%d images x %d points' % (nImg, nOP), 'order': nC, 'tiles': nb, 'band_tiles': bt,
                      'border_tile_rows': nborder, 'banded': cnt, 'dense': d,
                      'gemm_tile_ratio': round(cnt['gemm_tiles'] / d['gemm_tiles'], 4),
                      'gflop_banded': round(flop(cnt) / 1e9, 2), 'gflop_dense': round(flop(d) / 1e9, 2)}))


if __name__ == '__main__':
    main()
