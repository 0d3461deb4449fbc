"""CPU: the oracle against further known answers / properties of /root/reference/test/reduce.cpp (tests/golden/kat_reduce.py)."""
import numpy as np
import pytest

from oracle import pyoracle as po
from tests.golden import kat_reduce as kr


@pytest.fixture(scope="module")
def orc():
    return po.AfOracle("port")


def test_minmax_cplx_nan(orc):
    np.testing.assert_array_equal(orc.min(kr.CPLX_NAN, 0)[1].reshape(-1), kr.CPLX_NAN_MIN)
    np.testing.assert_array_equal(orc.max(kr.CPLX_NAN, 0)[1].reshape(-1), kr.CPLX_NAN_MAX)


def test_minmax_nan_property(orc):
    a = kr.nan_vector()
    lo, hi = float(orc.min_all(a)[1][0]), float(orc.max_all(a)[1][0])
    ok = a[~np.isnan(a)]
    assert not np.isnan(lo) and not np.isnan(hi) and lo == ok.min() and hi == ok.max()


def test_sum_product_nan(orc):
    a = kr.nan_vector(seed=11)
    assert np.isnan(float(orc.sum_all(a)[1][0]))
    want = float(np.where(np.isnan(a), 0, a).astype(np.float64).sum())
    got = float(orc.sum_nan_all(a, 0.0)[1][0])
    assert abs(got / a.size - want / a.size) < 1e-5
    p = np.random.default_rng(12).random(5).astype(np.float32)
    p[2] = np.nan
    assert np.isnan(float(orc.product_all(p)[1][0]))
    got = float(orc.product_nan_all(p, 1.0)[1][0])
    assert abs(got - float(np.prod(np.where(np.isnan(p), 1, p).astype(np.float64)))) < 1e-5


def test_count_anyall_nan(orc):
    a, b = kr.count_nan_pair()
    assert int(orc.count_all(a)[1][0]) == int(orc.count_all(b)[1][0]) == int(np.count_nonzero(b))
    a, b = kr.anyall_nan_pair()
    assert bool(orc.any_true_all(b)[1][0]) and bool(orc.all_true_all(b)[1][0])
    assert bool(orc.any_true_all(a)[1][0]) and not bool(orc.all_true_all(a)[1][0])


def test_product_of_bools_issue_2543(orc):
    ones = np.ones((5, 5), np.int8, order="F")
    assert int(orc.product_all(ones, dtype=po.B8)[1][0]) > 0
    x = np.asfortranarray((np.random.default_rng(13).random((5, 5)) < 0.8).astype(np.int8))
    assert int(orc.product_all(x, dtype=po.B8)[1][0]) == int(np.prod(x.astype(np.uint32)))
    t_all, w_all = orc.all_true(x, 0, dtype=po.B8)
    t_prod, w_prod = orc.product(x, 0, dtype=po.B8)
    assert w_all.reshape(-1).astype(np.int64).tolist() == w_prod.reshape(-1).astype(np.int64).tolist()


def test_nanval_issue_3255(orc):
    for i in range(8):
        v = kr.ISSUE_3255_VALS.copy()
        v[i] = np.nan
        for fn in ("product_by_key", "sum_by_key"):
            tk, wk, tv, wv = getattr(orc, fn)(kr.ISSUE_3255_KEYS, v, 0, nanval=1.0)
            assert not np.isnan(wv).any(), (fn, i)
            assert wk.reshape(-1).tolist() == [0, 1, 2, 0]


def test_reduce_half(orc):
    for value, dims, dim in kr.reduce_half_cases():
        x = np.full(dims, value, np.float16, order="F")
        d = dim if dim >= 0 else 0
        n = dims[d]
        t, w = orc.sum(x, d, dtype=po.F16)
        assert t == po.F32 and np.all(w == np.float32(value * n)), (dims, dim)
        with np.errstate(over="ignore"):
            p = np.float32(value) ** n
        if np.isfinite(p):
            t, w = orc.product(x, d, dtype=po.F16)
            assert t == po.F32 and np.all(w == p), (dims, dim)
    r = kr.HALF_RAMP
    assert orc.min(r, 0, dtype=po.F16)[1].reshape(-1)[0] == np.float16(1) and orc.max(r, 0, dtype=po.F16)[1].reshape(-1)[0] == np.float16(7)
    assert int(orc.count(r, 0, dtype=po.F16)[1].reshape(-1)[0]) == 7
    assert bool(orc.any_true(r, 0, dtype=po.F16)[1].reshape(-1)[0]) and bool(orc.all_true(r, 0, dtype=po.F16)[1].reshape(-1)[0])
