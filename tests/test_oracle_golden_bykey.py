"""CPU: the oracle's by-key reductions and ragged max against the known answers of the reference's own tests
(tests/golden/kat_bykey.py, transcribed from /root/reference/test/reduce.cpp)."""
import numpy as np
import pytest

from oracle import pyoracle as po
from tests.golden import kat_bykey as kb

TYPES = [po.F32, po.F64, po.S32, po.U32, po.S64, po.U64, po.S16, po.U16, po.U8, po.S8]


@pytest.fixture(scope="module")
def orc():
    return po.AfOracle("port")


def as_type(vals, code):
    return np.asarray(vals).astype(po.NP_OF[code])


@pytest.mark.parametrize("code", TYPES)
def test_typed_by_key_kats(orc, code):
    for name, fn, keys, vals, rkeys, gold, where in kb.TYPED:
        k = np.asarray(keys, np.int32)
        tk, wk, tv, wv = getattr(orc, fn)(k, as_type(vals, code), 0, dtype=code)
        assert wk.reshape(-1).tolist() == rkeys, (name, where)
        assert wv.reshape(-1).astype(np.float64).tolist() == [float(g) for g in gold], (name, where, code)


@pytest.mark.parametrize("code", [po.F32, po.F64])
def test_by_key_nans(orc, code):
    name, fn, keys, vals, nanval, rkeys, gold = kb.NAN_CASE
    tk, wk, tv, wv = getattr(orc, fn)(np.asarray(keys, np.int32), as_type(vals, code), 0, dtype=code, nanval=nanval)
    assert wk.reshape(-1).tolist() == rkeys and wv.reshape(-1).tolist() == [float(g) for g in gold]


def test_by_key_tiled_dims(orc):
    """nDim0..nDim3ReduceByKey (test/reduce.cpp:862-976): the same line tiled along the other dims reduces to the same answer"""
    k = np.asarray(kb.KEYS_A, np.int32)
    v = np.asarray(kb.VALS_A, np.float32)
    gold = [0, 8, 6, 10, 4]
    for dim, shape in ((0, (8, 2, 2, 2)), (1, (2, 8)), (2, (2, 1, 8)), (3, (2, 1, 1, 8))):
        idx = [None] * len(shape)
        idx[dim] = slice(None)
        x = np.asfortranarray(np.broadcast_to(v[tuple(idx)], shape).astype(np.float32))
        tk, wk, tv, wv = orc.sum_by_key(k, x, dim, nanval=0.0)
        w = np.moveaxis(wv.reshape(po.dims4(wv.shape), order="F"), dim, 0).reshape(5, -1)
        assert all(w[:, j].tolist() == gold for j in range(w.shape[1])), (dim, w)


def test_ragged_max(orc):
    for name, arr, dims, rl, rdims, dim, gv, gi in kb.RAGGED:
        x = np.asarray(arr, np.int32).reshape(dims, order="F")
        r = np.asarray(rl, np.uint32).reshape(rdims, order="F")
        t, wv, wl = orc.max_ragged(x, r, dim)
        assert wv.reshape(-1).tolist() == gv and wl.reshape(-1).tolist() == gi, name


def test_issues(orc):
    key, runs = kb.issue_2955()
    val = np.random.default_rng(0).random(256).astype(np.float32)
    tk, wk, tv, wv = orc.sum_by_key(key, val, 0)
    assert wk.size == runs and wv.size == runs
    tk, wk, tv, wv = orc.sum_by_key(key, np.asfortranarray(np.random.default_rng(1).random((8, 256)).astype(np.float32)), 1)
    assert wk.size == runs and wv.reshape(po.dims4(wv.shape), order="F").shape[1] == runs
    zeros, ones, n = kb.issue_3062()
    assert int(orc.sum_by_key(zeros, ones, 0)[3].reshape(-1)[0]) == n
    assert int(orc.count_by_key(zeros, ones, 0)[3].reshape(-1)[0]) == n
    ones2 = np.asfortranarray(np.ones((2, 129), np.uint32))
    assert orc.sum_by_key(zeros, ones2, 1)[3].reshape(-1).tolist() == [n, n]
    assert orc.count_by_key(zeros, ones2, 1)[3].reshape(-1).tolist() == [n, n]
