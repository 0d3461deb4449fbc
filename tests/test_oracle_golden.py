"""Pins the restated oracle to the known-answer vectors the reference's own tests hold for this path
(tests/golden/kat.py, each transcribed with its /root/reference/test file:line)."""
import ctypes

import numpy as np
import pytest

from oracle import pyoracle as po
from tests.golden import kat


@pytest.fixture(scope="module", params=["port", "reference"])
def orc(request, port_oracle):
    if request.param == "port":
        return port_oracle
    if not po.Oracle.available("reference"):
        pytest.skip("oracle/_ref not built")
    return po.AfOracle("reference")


@pytest.mark.parametrize("case", kat.IREDUCE_DIM, ids=lambda c: c[0])
def test_ireduce_dim(orc, case):
    name, cite, op, x, dim, gold_val, gold_idx = case
    _, val, idx = getattr(orc, op)(x, dim)
    np.testing.assert_array_equal(val.reshape(-1, order="F"), gold_val)
    assert idx.reshape(-1, order="F").tolist() == gold_idx
    assert idx.dtype == np.uint32


@pytest.mark.parametrize("case", kat.IREDUCE_ALL, ids=lambda c: c[0])
def test_ireduce_all(orc, case):
    name, cite, op, x, gold_val, gold_idx = case
    val, idx = getattr(orc, op)(x)
    assert val == gold_val and idx == gold_idx


@pytest.mark.parametrize("case", kat.DIMWISE, ids=lambda c: c[0])
def test_dimwise(orc, case):
    name, cite, op, x, dim, gold, gdims = case
    t, out = getattr(orc, op)(x, dim)
    assert t == po.F32
    assert tuple(po.dims4(out.shape)) == gdims
    assert out.reshape(-1, order="F").tolist() == gold


def test_ireduce_dim_ge_ndims(orc):
    """test/ireduce.cpp:176-194: reducing a singleton trailing dim returns the input and zero indices."""
    x = np.asfortranarray(np.random.default_rng(0).random((10, 10, 1)).astype(np.float32))
    for op in ("imin", "imax"):
        t, v, l = getattr(orc, op)(x, 2)
        np.testing.assert_array_equal(v.reshape(10, 10), x.reshape(10, 10))
        assert not l.any()


@pytest.mark.parametrize("case", list(kat.exclusive_ones_cases()), ids=lambda c: c[0])
def test_exclusive_scan_of_ones(orc, case):
    name, x, dim, gold = case
    t, out = orc.scan(x, dim, po.AF_BINARY_ADD, False)
    assert t == po.S32
    np.testing.assert_array_equal(out.reshape(po.dims4(out.shape), order="F"), gold.reshape(po.dims4(gold.shape), order="F"))


def test_crand_matches_glibc():
    libc = ctypes.CDLL("libc.so.6")
    for seed in (0, 1, 12345):
        libc.srand(seed)
        r = kat.CRand(seed)
        assert [libc.rand() for _ in range(1000)] == [r.rand() for _ in range(1000)]


SBK = [(op, incl, dims, dim, np_dt, lo, hi)
       for op in (0, 2, 3) for incl in (True, False)
       for dims, dim in (((16 * 1024, 64), 0), ((1024, 128), 1))
       for np_dt, lo, hi in ((np.int32, -15, 15), (np.float32, -5, 5))]


@pytest.mark.parametrize("op,incl,dims,dim,np_dt,lo,hi", SBK)
def test_scan_by_key_reference_generator(orc, op, incl, dims, dim, np_dt, lo, hi):
    """test/scan_by_key.cpp:111-195: the reference's srand(0) key/data generator (node lengths {37,256},
    keys [0,15)) checked with the reference test's own verify<> (second dim shrunk to keep CPU time low)."""
    key, data = kat.scan_by_key_case(dims, dim, [37, 256], 0, 15, lo, hi, np_dt)
    t, out = orc.scan_by_key(key, data, dim, op, incl)
    assert kat.scan_by_key_verify(dims, data, key, out, dim, op, incl, 1e-3) == 0


def test_scan_by_key_simple(orc):
    """test/scan_by_key.cpp:197-223 Simple_0 / Simple_1"""
    for dims, dim in (((16, 8, 2, 1), 0), ((8, 256 + 128, 1, 1), 1)):
        key, data = kat.scan_by_key_case(dims, dim, [4, 8], 0, 15, 2, 4, np.int32)
        t, out = orc.scan_by_key(key, data, dim, po.AF_BINARY_ADD, False)
        assert kat.scan_by_key_verify(dims, data, key, out, dim, 0, False, 1e-5) == 0


def test_where_empty_and_maxdim(orc):
    """test/where.cpp:134-138 (no hits -> 0 elements) and :121-132 (MaxDim, shrunk)."""
    a = np.asfortranarray(np.random.default_rng(1).random((10, 10, 10)).astype(np.float32))
    assert orc.where((a > 2)).size == 0
    n = 65535 * 32 + 2
    x = (np.arange(n) % 2 == 0).reshape((1, n), order="F")
    np.testing.assert_array_equal(orc.where(x), (2 * np.arange(n // 2)).astype(np.uint32))
    x4 = x.reshape((1, 1, 1, n), order="F")
    np.testing.assert_array_equal(orc.where(x4), (2 * np.arange(n // 2)).astype(np.uint32))


def test_global_reductions_match_serial_host_loop(orc):
    """test/reduce.cpp:992-1049: sum/count/min/max over 10000 elements equal a serial host loop exactly."""
    rng = np.random.default_rng(7)
    a = np.round(2 * rng.random(10000)).astype(np.float32)
    gold = np.float32(0)
    for v in a:
        gold = np.float32(gold + v)
    assert orc.sum_all(a)[1][0] == gold
    b = (a != 0)
    assert orc.count_all(b)[1][0] == int(b.sum())
    d = rng.random(10000)
    assert orc.min_all(d)[1][0] == d.min()
    f = rng.random(10000).astype(np.float32)
    assert orc.max_all(f)[1][0] == f.max()


def test_nan_rules(orc):
    """test/reduce.cpp:1147-1169, 1236-1299: min/max never return NaN; count treats NaN as non-zero;
    sum propagates NaN unless the _nan variant substitutes; any/all treat NaN as true."""
    rng = np.random.default_rng(8)
    A = rng.random(10000).astype(np.float32)
    A[A < 0.25] = np.nan
    mn, mx = orc.min_all(A)[1][0], orc.max_all(A)[1][0]
    assert mn == np.nanmin(A) and mx == np.nanmax(A)
    B = np.round(5 * rng.random(10000)).astype(np.float32)
    Bn = B.copy()
    Bn[B == 2] = np.nan
    assert orc.count_all(Bn)[1][0] == orc.count_all(B)[1][0]
    assert np.isnan(orc.sum_all(A)[1][0])
    s = orc.sum_nan_all(A, 0.0)[1][0]
    assert abs(s / 10000 - np.nansum(A.astype(np.float64)) / 10000) < 1e-5
    P = rng.random(5).astype(np.float32)
    P[2] = np.nan
    assert np.isnan(orc.product_all(P)[1][0])
    assert abs(orc.product_nan_all(P, 1.0)[1][0] - np.nanprod(P)) < 1e-5
    Z = (rng.random(10000) > 0.5).astype(np.float32)
    Zn = Z.copy()
    Zn[Z == 0] = np.nan
    assert orc.any_true_all(Zn)[1][0] == 1 and orc.all_true_all(Zn)[1][0] == 1
    assert orc.any_true_all(Z)[1][0] == 1 and orc.all_true_all(Z)[1][0] == 0


def test_b8_rules(orc):
    """test/reduce.cpp:1357-1374: b8 product == allTrue; b8 sum == count."""
    rng = np.random.default_rng(9)
    A = rng.integers(0, 2, (5, 5)).astype(bool, order="F")
    t1, p = orc.product(A, 0)
    t2, a = orc.all_true(A, 0)
    assert t1 == t2 == po.B8
    np.testing.assert_array_equal(p, a)
    t3, s = orc.sum(A, 0)
    assert t3 == po.U32
    np.testing.assert_array_equal(s.reshape(-1), A.sum(axis=0).astype(np.uint32))
    ones = np.ones((5, 5), bool, order="F")
    assert orc.product_all(ones)[1][0] > 0
