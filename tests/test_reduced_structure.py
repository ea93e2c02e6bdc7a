"""Structure of the reduced camera system (groundwork for SURVEY §8(f) N4; tools/reduced_sparsity.py).

Executable form of the design note in DESIGN.md §9: the block pattern of S is the co-visibility graph of the
images; with the images in reverse Cuthill-McKee order and the shared IO block last, the Cholesky factor stays
inside the band of that pattern; in x order (IO first, what the device path uses today) it fills completely."""
import os
import sys

import numpy as np
import scipy.sparse as sp
from scipy.sparse.csgraph import reverse_cuthill_mckee
from scipy.sparse.linalg import splu

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tools'))


def test_band_of_the_reduced_system_survives_only_with_the_io_block_last():
    from dbat_b200.synth import make_scene
    from dbat_b200.dbatstruct import seteoest_depend
    from oracle.dbatstruct import buildserialindices, serialize, buildweightmatrix
    from oracle.cameramodel import brown_euler_cam4
    from reduced_sparsity import covisibility, symbolic_cholesky
    s, _ = make_scene(100, 4000, rays=8, seed=3, build_indices=False)
    seteoest_depend(s, 0)
    buildserialindices(s)
    x, W = serialize(s), buildweightmatrix(s)
    J = brown_euler_cam4(x, s, True)[1]
    Jw = (sp.diags(np.sqrt(W)) @ J).tocsc()
    N = (Jw.T @ Jw).tocsc()
    nIO, nEO = len(s.bundle.serial.IO.dest), len(s.bundle.serial.EO.dest)
    nC = nIO + nEO
    B = N[:nC, nC:]
    S = N[:nC, :nC].toarray() - (B @ splu(N[nC:, nC:].tocsc()).solve(B.T.toarray()))
    d = 1 / np.sqrt(np.diag(S))
    S = S * d[:, None] * d[None, :]
    nImg = s.EO.val.shape[1]
    eo_col = np.full(6 * nImg, -1)
    eo_col[s.bundle.serial.EO.src] = s.bundle.serial.EO.dest
    eo_col = eo_col.reshape(6, nImg, order='F')
    # (1) the EO x EO block pattern of S is the co-visibility graph
    G = covisibility(np.asarray(s.IP.img), np.asarray(s.IP.op), nImg)
    img_of = np.full(nC, -1)
    for i in range(nImg):
        img_of[eo_col[:, i][eo_col[:, i] >= 0]] = i
    r, c = np.nonzero(np.abs(S) > 1e-13)
    both = (img_of[r] >= 0) & (img_of[c] >= 0)
    pat = sp.csr_matrix((np.ones(both.sum()), (img_of[r[both]], img_of[c[both]])), shape=(nImg, nImg)).astype(bool)
    free = np.flatnonzero((eo_col >= 0).any(axis=0))              # the datum camera has no columns in S
    assert (pat[free][:, free] != G[free][:, free]).nnz == 0 and pat.nnz == G[free][:, free].nnz
    # (2) RCM image order, IO last: the factor's EO part stays inside the band of the pattern
    perm = np.asarray(reverse_cuthill_mckee(G.tocsr(), symmetric_mode=True))
    inv = np.empty(nImg, int)
    inv[perm] = np.arange(nImg)
    coo = G.tocoo()
    band = int(np.abs(inv[coo.row] - inv[coo.col]).max())
    assert band < nImg // 2
    order = np.array([k for i in perm for k in eo_col[:, i] if k >= 0] + list(s.bundle.serial.IO.dest))
    L = np.linalg.cholesky(S[np.ix_(order, order)])
    i, j = np.nonzero(np.abs(L[:nEO, :nEO]) > 1e-13)
    assert (i - j).max() <= 6 * (band + 1) - 1
    # and the symbolic factorisation of the tool predicts a superset of the numerical fill
    cnt = symbolic_cholesky(G, perm)
    blocks = {(inv[img_of[order[a]]], inv[img_of[order[b]]]) for a, b in zip(i, j)}
    assert len({(a, b) for a, b in blocks if a > b}) <= cnt.sum()
    # (3) x order (IO first, today's layout of S): the arrowhead fills the whole factor
    L0 = np.linalg.cholesky(S)
    assert np.count_nonzero(np.abs(L0) > 1e-13) > 0.98 * nC * (nC + 1) / 2


def test_banded_tile_cholesky_prototype_equals_the_dense_factor():
    """tools/band_cholesky_proto.py: the tile loop restricted to band + border tiles gives the dense factor
    on a band-plus-border SPD matrix, with the predicted number of tile operations."""
    from band_cholesky_proto import band_tile_cholesky, dense_counts
    rng = np.random.default_rng(1)
    NB, nb, bt, nborder = 8, 12, 3, 2
    n = NB * nb
    ne = nb - nborder
    M = rng.standard_normal((n, n))
    keep = np.zeros((nb, nb), bool)
    for i in range(nb):
        for j in range(nb):
            keep[i, j] = abs(i - j) <= bt or i >= ne or j >= ne
    M = M * np.kron(keep, np.ones((NB, NB)))
    A = M @ M.T                                            # band 2*bt in tiles ... so use it as the pattern source
    A = A * np.kron(keep, np.ones((NB, NB))) + n * np.eye(n)   # SPD, band bt + dense border
    L, cnt = band_tile_cholesky(A, NB, bt, nborder)
    np.testing.assert_allclose(L, np.linalg.cholesky(A), rtol=1e-10, atol=1e-12)
    assert cnt['gemm_tiles'] < dense_counts(nb)['gemm_tiles'] and cnt['potrf'] == nb
    # one tile less of band is not enough: the factor is then wrong
    Lbad, _ = band_tile_cholesky(A, NB, bt - 1, nborder)
    assert np.abs(Lbad - np.linalg.cholesky(A)).max() > 1e-6


def test_library_camera_order_is_a_valid_rcm_order():
    """`dbat_camera_order` (host code in libdbatgpu.so, no device needed): a permutation whose bandwidth on
    the co-visibility graph is reported correctly, is far below the natural one for a shuffled block and
    comparable to SciPy's reverse Cuthill-McKee; isolated images and empty input are handled."""
    from dbat_b200 import _lib
    from dbat_b200.synth import make_scene
    from reduced_sparsity import covisibility
    s, _ = make_scene(196, 6000, rays=8, seed=5, build_indices=False)
    nImg, nOP = s.EO.val.shape[1], s.OP.val.shape[1]
    rng = np.random.default_rng(0)
    shuffle = rng.permutation(nImg)                               # destroy the generator's strip order
    img, op = shuffle[np.asarray(s.IP.img)], np.asarray(s.IP.op)
    perm, bw = _lib.camera_order(img, op, nImg, nOP)
    assert sorted(perm) == list(range(nImg))
    G = covisibility(img, op, nImg).tocoo()
    pos = np.empty(nImg, int)
    pos[perm] = np.arange(nImg)
    assert bw == np.abs(pos[G.row] - pos[G.col]).max()
    natural = np.abs(G.row - G.col).max()
    ref = np.asarray(reverse_cuthill_mckee(G.tocsr(), symmetric_mode=True))
    rpos = np.empty(nImg, int)
    rpos[ref] = np.arange(nImg)
    scipy_bw = np.abs(rpos[G.row] - rpos[G.col]).max()
    assert bw < 0.5 * natural and bw <= 1.3 * scipy_bw
    # two images that see nothing in common with the rest, and no observations at all
    perm2, bw2 = _lib.camera_order(np.array([0, 1, 0, 1]), np.array([0, 0, 1, 1]), 4, 2)
    assert sorted(perm2) == [0, 1, 2, 3] and bw2 == 1
    perm3, bw3 = _lib.camera_order(np.zeros(0, int), np.zeros(0, int), 3, 0)
    assert sorted(perm3) == [0, 1, 2] and bw3 == 0
    import pytest
    with pytest.raises(_lib.DbatError):
        _lib.camera_order(np.array([5]), np.array([0]), 3, 1)
