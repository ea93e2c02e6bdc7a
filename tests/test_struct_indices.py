"""Index contract (buildserialindices / serialize / deserialize / buildweightmatrix):
oracle vs the reference's worked examples in plan/re-index.org, and product vs oracle."""
import copy

import numpy as np
import pytest

import dbat_b200.dbatstruct as prod
import oracle.dbatstruct as orc


def _dist_matrix(des, shape):
    d = np.zeros(shape[0] * shape[1], dtype=int)
    d[des.dest] = des.src + 1
    return d.reshape(shape, order='F')


EX1_BLOCK = np.array([[1, 1, 1, 1, 1], [1, 2, 3, 4, 5], [1, 2, 3, 4, 5], [1, 1, 1, 1, 1], [1, 1, 1, 1, 1]])
EX1_SER = [1, 2, 3, 4, 5, 7, 8, 12, 13, 17, 18, 22, 23]
EX1_DES = [[1, 1, 1, 1, 1], [2, 6, 8, 10, 12], [3, 7, 9, 11, 13], [4, 4, 4, 4, 4], [5, 5, 5, 5, 5]]
EX3_BLOCK = np.array([[1, 1, 1, 1, 1, 1], [1, 1, 2, 2, 3, 3], [1, 1, 2, 2, 3, 3], [1, 1, 1, 1, 1, 1],
                      [1, 1, 1, 1, 1, 1], [0, 0, 0, 0, 0, 0]])
EX3_SER = [1, 2, 3, 4, 5, 14, 15, 26, 27]
EX3_DES = [[1] * 6, [2, 2, 6, 6, 8, 8], [3, 3, 7, 7, 9, 9], [4] * 6, [5] * 6, [0] * 6]


@pytest.mark.parametrize('mod', [orc, prod])
@pytest.mark.parametrize('block,ser,des', [(EX1_BLOCK, EX1_SER, EX1_DES), (EX3_BLOCK, EX3_SER, EX3_DES)])
def test_reindex_org_examples(mod, block, ser, des):
    """plan/re-index.org: serialize / deserialize tables (1-based in the document)."""
    est = block > 0
    use = np.zeros(block.shape, bool)
    leading, serial, deserial, _ = mod._serializeblock(block.copy(), est, use)[:4]
    assert list(serial.src + 1) == ser
    assert np.array_equal(_dist_matrix(deserial, block.shape), np.array(des))


def _random_struct(rng, nImg=7, nOP=12, shared_io=True, prior=True):
    NC = 10
    ip_img, ip_op = np.nonzero(rng.random((nImg, nOP)) < 0.7)
    s = prod.new_struct(rng.random((NC, nImg)), rng.random((6, nImg)), rng.random((3, nOP)),
                        rng.random((2, len(ip_img))) * 100, ip_img, ip_op, np.array([[0.01], [0.012]]),
                        np.array([[640.], [480.]]), IPstd=0.3)
    if shared_io:
        s.IO.struct.block[:] = 1
    else:
        s.IO.struct.block = np.tile(np.arange(1, nImg + 1), (NC, 1))
        s.IO.struct.block[0, :] = 1                      # focal shared, the rest image-variant
        s.IO.struct.block[1:3, 0:4] = [[1, 1, 2, 2]] * 2  # pp shared pairwise for the first images
    s.bundle.est.IO[:] = True
    s.bundle.est.IO[4, :] = False
    s.bundle.est.EO[:] = True
    s.bundle.est.EO[:, 0] = False
    s.bundle.est.OP[:] = rng.random((3, nOP)) < 0.9
    if prior:
        s.prior.EO.use[0:3, 2] = True
        s.prior.EO.val[0:3, 2] = 1.0
        s.prior.EO.std[0:3, 2] = 0.1
        s.prior.OP.use[:, 3] = s.bundle.est.OP[:, 3]
        s.prior.OP.val[:, 3] = 0.5
        s.prior.OP.std[:, 3] = 0.05
        s.prior.IO.use[0, :] = True
        s.prior.IO.val[0, :] = 7.0
        s.prior.IO.std[0, :] = 0.2
    return s


@pytest.mark.parametrize('shared_io', [True, False])
def test_product_indices_equal_oracle(shared_io):
    rng = np.random.default_rng(3)
    for _ in range(5):
        s = _random_struct(rng, shared_io=shared_io)
        so = copy.deepcopy(s)
        prod.buildserialindices(s)
        orc.buildserialindices(so)
        for nm in ('IO', 'EO', 'OP'):
            a, b = getattr(s.bundle.serial, nm), getattr(so.bundle.serial, nm)
            assert np.array_equal(a.src, b.src) and np.array_equal(a.dest, b.dest) and np.array_equal(a.obs, b.obs)
            a, b = getattr(s.bundle.deserial, nm), getattr(so.bundle.deserial, nm)
            assert np.array_equal(a.src, b.src) and np.array_equal(a.dest, b.dest)
        assert s.bundle.serial.n == so.bundle.serial.n
        for k in ('IP', 'IO', 'EO', 'OP'):
            assert np.array_equal(getattr(s.post.res.ix, k), getattr(so.post.res.ix, k))
        assert np.array_equal(s.prior.IO.use, so.prior.IO.use)
        x, xo = prod.serialize(s), orc.serialize(so)
        assert np.array_equal(x, xo)
        np.testing.assert_array_equal(prod.buildweightmatrix(s), orc.buildweightmatrix(so))
        x2 = x + 1.0
        prod.deserialize(s, x2)
        IO, EO, OP = orc.deserialize(so, x2)
        assert np.array_equal(s.IO.val, IO) and np.array_equal(s.EO.val, EO) and np.array_equal(s.OP.val, OP)
        # round trip: shared elements received the same value everywhere
        assert np.array_equal(prod.serialize(s), x2)


def test_depend_datum():
    rng = np.random.default_rng(0)
    s = _random_struct(rng, prior=False)
    so = copy.deepcopy(s)
    prod.seteoest_depend(s, 0)
    orc.seteoest_depend(so, 0)
    assert np.array_equal(s.bundle.est.EO, so.bundle.est.EO)
    assert s.bundle.est.EO.size - s.bundle.est.EO.sum() == 7      # seteoest.m:125-128


def test_deserialize_rewinds_to_an_iteration_of_the_trace():
    """deserialize.m:31-100: deserialize(s,E,i) puts iteration i of E.trace back into the struct (the
    reference's only resume mechanism); deserialize(s,E,v,'EO') stacks a parameter group over iterations."""
    from types import SimpleNamespace as NS
    rng = np.random.default_rng(5)
    s = _random_struct(rng, shared_io=True)
    prod.buildserialindices(s)
    x0 = prod.serialize(s)
    trace = np.stack([x0 + k for k in range(4)], axis=1)          # 4 iterations, every unknown +k
    E = NS(trace=trace)
    fixedEO = s.EO.val.copy()
    for i, col in ((0, 0), (2, 2), (np.inf, 3)):
        t = prod.deserialize(copy.deepcopy(s), E, i)
        assert np.array_equal(prod.serialize(t), trace[:, col])
        est = s.bundle.est.EO
        assert np.array_equal(t.EO.val[~est], fixedEO[~est])       # what is not estimated keeps its value
    EO = prod.deserialize(s, E, 'all', 'EO')
    assert EO.shape == s.EO.val.shape + (4,)
    for k in range(4):
        np.testing.assert_array_equal(EO[:, :, k], prod.deserialize(copy.deepcopy(s), E, k).EO.val)
    IO = prod.deserialize(s, E, [1, 3], 'IO')
    assert IO.shape[2] == 2 and np.array_equal(IO[:, :, 1], prod.deserialize(copy.deepcopy(s), E, 3).IO.val)
    assert np.array_equal(prod.serialize(s), x0)                   # the 4-argument form leaves s alone
    with pytest.raises(ValueError):
        prod.deserialize(s, None, 1)
