"""`-m gpu`: the known-answer vectors the reference's own tests hold for this path (tests/golden/kat.py, each transcribed
with its /root/reference/test file:line) run through the CUDA path via the C-ABI -- the same cases
tests/test_oracle_golden.py pins the oracle with -- plus the reference tests that only make sense on a device:
Accum.MaxDim / Where.MaxDim (grid-dimension folding, test/scan.cpp:168-210, test/where.cpp:121-132) and
FixOverflowWrite (test/scan_by_key.cpp:225-243)."""
import numpy as np
import pytest

from oracle import pyoracle as po
from tests.golden import kat
from tests.gpu_util import gpu_available

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def af():
    if not gpu_available():
        pytest.skip("no CUDA device")
    import arrayfire_b200
    return arrayfire_b200


def dev(af, x):
    """device array of a KAT input; the one strided view among them (a(seq(1,3), seq(2,4)), test/ireduce.cpp:424-431)
    goes up as its 5x5 parent and is re-created with af_index, so the sub-array path is what runs"""
    if x is kat.SUBA_VIEW:
        return af.Array.from_numpy(kat._SUBA).index(af.seq(1, 3, 1), af.seq(2, 4, 1))
    return af.Array.from_numpy(np.asfortranarray(x))


@pytest.mark.parametrize("case", kat.IREDUCE_DIM, ids=lambda c: c[0])
def test_ireduce_dim(af, case):
    name, cite, op, x, dim, gold_val, gold_idx = case
    val, idx = getattr(af, op)(dev(af, x), dim)
    np.testing.assert_array_equal(val.to_numpy().reshape(-1, order="F"), gold_val)
    assert idx.to_numpy().reshape(-1, order="F").tolist() == gold_idx
    assert idx.dtype == po.U32


@pytest.mark.parametrize("case", kat.IREDUCE_ALL, ids=lambda c: c[0])
def test_ireduce_all(af, case):
    name, cite, op, x, gold_val, gold_idx = case
    val, idx = getattr(af, op)(dev(af, x))
    assert val == float(gold_val) and idx == gold_idx


@pytest.mark.parametrize("case", kat.DIMWISE, ids=lambda c: c[0])
def test_dimwise(af, case):
    name, cite, op, x, dim, gold, gdims = case
    out = getattr(af, op)(dev(af, x), dim)
    assert out.dtype == po.F32 and out.dims == gdims
    assert out.to_numpy().reshape(-1, order="F").tolist() == gold


def test_ireduce_dim_ge_ndims(af):
    """test/ireduce.cpp:176-194"""
    x = np.asfortranarray(np.random.default_rng(0).random((10, 10, 1)).astype(np.float32))
    for op in ("imin", "imax"):
        v, l = getattr(af, op)(af.Array.from_numpy(x), 2)
        np.testing.assert_array_equal(v.to_numpy().reshape(10, 10), x.reshape(10, 10))
        assert not l.to_numpy().any()


@pytest.mark.parametrize("case", list(kat.exclusive_ones_cases()), ids=lambda c: c[0])
def test_exclusive_scan_of_ones(af, case):
    name, x, dim, gold = case
    out = af.scan(af.Array.from_numpy(x), dim, af.AF_BINARY_ADD, False)
    assert out.dtype == po.S32
    np.testing.assert_array_equal(out.to_numpy().reshape(po.dims4(gold.shape), order="F"), gold.reshape(po.dims4(gold.shape), order="F"))


SBK = [(op, incl, dims, dim, np_dt, lo, hi)
       for op in (0, 2, 3) for incl in (True, False)
       for dims, dim in (((16 * 1024, 64), 0), ((1024, 128), 1))
       for np_dt, lo, hi in ((np.int32, -15, 15), (np.float32, -5, 5))]


@pytest.mark.parametrize("op,incl,dims,dim,np_dt,lo,hi", SBK)
def test_scan_by_key_reference_generator(af, port_oracle, op, incl, dims, dim, np_dt, lo, hi):
    """test/scan_by_key.cpp:111-195 with the reference test's own verify<>, and bit-exact against the oracle"""
    key, data = kat.scan_by_key_case(dims, dim, [37, 256], 0, 15, lo, hi, np_dt)
    out = af.scan_by_key(af.Array.from_numpy(key), af.Array.from_numpy(data), dim, op, incl).to_numpy()
    assert kat.scan_by_key_verify(dims, data, key, out, dim, op, incl, 1e-3) == 0
    want = port_oracle.scan_by_key(key, data, dim, op, incl)[1]
    assert out.reshape(-1, order="F").tobytes() == want.reshape(-1, order="F").tobytes()


def test_scan_by_key_simple(af):
    """test/scan_by_key.cpp:197-223 Simple_0 / Simple_1"""
    for dims, dim in (((16, 8, 2, 1), 0), ((8, 256 + 128, 1, 1), 1)):
        key, data = kat.scan_by_key_case(dims, dim, [4, 8], 0, 15, 2, 4, np.int32)
        out = af.scan_by_key(af.Array.from_numpy(key), af.Array.from_numpy(data), dim, af.AF_BINARY_ADD, False).to_numpy()
        assert kat.scan_by_key_verify(dims, data, key, out, dim, 0, False, 1e-5) == 0


def test_scan_by_key_fix_overflow_write(af):
    """test/scan_by_key.cpp:225-243 FixOverflowWrite: N = 41000, keys all equal; a following allocation must not be
    clobbered by an out-of-bounds write of the scan"""
    n = 41000
    keys = af.constant(0, (n,), af.s32)
    vals = af.constant(1.0, (n,), af.f32)
    guard = af.constant(0.0, (64, 1), af.f32)
    out = af.scan_by_key(keys, vals, 0, af.AF_BINARY_ADD, True)
    np.testing.assert_array_equal(out.to_numpy().reshape(-1), np.arange(1, n + 1, dtype=np.float32))
    assert af.sum_all(guard) == 0.0
    assert af.sum_all(vals) == float(n)


def test_accum_max_dim(af):
    """test/scan.cpp:168-210 Accum.MaxDim: dims beyond 65535 * 32 along dims 1..3 (grid folding in the reference)"""
    n = 65535 * 32 + 1
    for shape, dim in (((1, n), 0), ((1, 1, n), 0), ((1, 1, 1, n), 0), ((2, n), 1), ((1, 2, n // 2 + 3), 2)):
        x = af.constant(1.0, shape, af.f32)
        out = af.accum(x, dim).to_numpy().reshape(shape, order="F")
        want = np.cumsum(np.ones(shape, np.float32), axis=dim)
        np.testing.assert_array_equal(out, want)


def test_where_max_dim_and_empty(af):
    """test/where.cpp:121-132 (MaxDim) and :134-138 (ISSUE_1259: no hits -> 0 elements)"""
    n = 65535 * 32 + 2
    host = (np.arange(n) % 2 == 0)
    for shape in ((1, n), (1, 1, n), (1, 1, 1, n)):
        x = np.asfortranarray(host.reshape(shape, order="F"))
        got = af.where(af.Array.from_numpy(x)).to_numpy().reshape(-1)
        np.testing.assert_array_equal(got, (2 * np.arange(n // 2)).astype(np.uint32))
    a = np.asfortranarray(np.random.default_rng(1).random((10, 10, 10)).astype(np.float32))
    r = af.where(af.Array.from_numpy((a > 2)))
    assert r.elements == 0 and r.dtype == po.U32


def test_global_reductions_match_serial_host_loop(af):
    """test/reduce.cpp:992-1049"""
    rng = np.random.default_rng(7)
    a = np.round(2 * rng.random(10000)).astype(np.float32)
    gold = np.float32(0)
    for v in a:
        gold = np.float32(gold + v)
    assert af.sum_all(af.Array.from_numpy(a)) == float(gold)
    b = (a != 0)
    assert af.count_all(af.Array.from_numpy(b)) == int(b.sum())
    d = rng.random(10000)
    assert af.min_all(af.Array.from_numpy(d)) == d.min()
    f = rng.random(10000).astype(np.float32)
    assert af.max_all(af.Array.from_numpy(f)) == float(f.max())


def test_nan_rules(af):
    """test/reduce.cpp:1147-1169, 1236-1299"""
    rng = np.random.default_rng(8)
    A = rng.random(10000).astype(np.float32)
    A[A < 0.25] = np.nan
    gA = af.Array.from_numpy(A)
    assert af.min_all(gA) == float(np.nanmin(A)) and af.max_all(gA) == float(np.nanmax(A))
    B = np.round(5 * rng.random(10000)).astype(np.float32)
    Bn = B.copy()
    Bn[B == 2] = np.nan
    assert af.count_all(af.Array.from_numpy(Bn)) == af.count_all(af.Array.from_numpy(B))
    assert np.isnan(af.sum_all(gA))
    s = af.sum_all(gA, nanval=0.0)
    assert abs(s / 10000 - np.nansum(A.astype(np.float64)) / 10000) < 1e-5
    P = rng.random(5).astype(np.float32)
    P[2] = np.nan
    gP = af.Array.from_numpy(P)
    assert np.isnan(af.product_all(gP))
    assert abs(af.product_all(gP, nanval=1.0) - np.nanprod(P)) < 1e-5
    Z = (rng.random(10000) > 0.5).astype(np.float32)
    Zn = Z.copy()
    Zn[Z == 0] = np.nan
    assert af.any_true_all(af.Array.from_numpy(Zn)) == 1 and af.all_true_all(af.Array.from_numpy(Zn)) == 1
    assert af.any_true_all(af.Array.from_numpy(Z)) == 1 and af.all_true_all(af.Array.from_numpy(Z)) == 0


def test_b8_rules(af):
    """test/reduce.cpp:1357-1374: b8 product == allTrue; b8 sum == count"""
    rng = np.random.default_rng(9)
    A = rng.integers(0, 2, (5, 5)).astype(bool, order="F")
    g = af.Array.from_numpy(A)
    p, a = af.product(g, 0), af.all_true(g, 0)
    assert p.dtype == a.dtype == po.B8
    np.testing.assert_array_equal(p.to_numpy(), a.to_numpy())
    s = af.sum(g, 0)
    assert s.dtype == po.U32
    np.testing.assert_array_equal(s.to_numpy().reshape(-1), A.sum(axis=0).astype(np.uint32))


# ---- by-key reductions and ragged max: the known answers of test/reduce.cpp (tests/golden/kat_bykey.py) -------------------
from tests.golden import kat_bykey as kb  # noqa: E402

BYKEY_TYPES = [po.F32, po.F64, po.S32, po.U32, po.S64, po.U64, po.S16, po.U16, po.U8, po.S8]


@pytest.mark.parametrize("code", BYKEY_TYPES)
def test_by_key_kats(af, code):
    for name, fn, keys, vals, rkeys, gold, where in kb.TYPED:
        gk = af.Array.from_numpy(np.asarray(keys, np.int32))
        gv = af.Array.from_numpy(np.asarray(vals).astype(po.NP_OF[code]), dtype=code)
        ko, vo = getattr(af, fn)(gk, gv, 0)
        assert ko.to_numpy().reshape(-1).tolist() == rkeys, (name, where)
        assert vo.to_numpy().reshape(-1).astype(np.float64).tolist() == [float(g) for g in gold], (name, where, code)


@pytest.mark.parametrize("code", [po.F32, po.F64])
def test_by_key_nans_kat(af, code):
    name, fn, keys, vals, nanval, rkeys, gold = kb.NAN_CASE
    ko, vo = getattr(af, fn)(af.Array.from_numpy(np.asarray(keys, np.int32)),
                             af.Array.from_numpy(np.asarray(vals).astype(po.NP_OF[code])), 0, nanval=nanval)
    assert ko.to_numpy().reshape(-1).tolist() == rkeys and vo.to_numpy().reshape(-1).tolist() == [float(g) for g in gold]


def test_by_key_tiled_dims_kat(af):
    k = np.asarray(kb.KEYS_A, np.int32)
    v = np.asarray(kb.VALS_A, np.float32)
    gold = [0, 8, 6, 10, 4]
    for dim, shape in ((0, (8, 2, 2, 2)), (1, (2, 8)), (2, (2, 1, 8)), (3, (2, 1, 1, 8))):
        idx = [None] * len(shape)
        idx[dim] = slice(None)
        x = np.asfortranarray(np.broadcast_to(v[tuple(idx)], shape).astype(np.float32))
        ko, vo = af.sum_by_key(af.Array.from_numpy(k), af.Array.from_numpy(x), dim, nanval=0.0)
        w = np.moveaxis(vo.to_numpy().reshape(po.dims4(vo.to_numpy().shape), order="F"), dim, 0).reshape(5, -1)
        assert all(w[:, j].tolist() == gold for j in range(w.shape[1])), (dim, w)


def test_ragged_max_kat(af):
    for name, arr, dims, rl, rdims, dim, gv, gi in kb.RAGGED:
        x = np.asfortranarray(np.asarray(arr, np.int32).reshape(dims, order="F"))
        r = np.asfortranarray(np.asarray(rl, np.uint32).reshape(rdims, order="F"))
        val, idx = af.max_ragged(af.Array.from_numpy(x), af.Array.from_numpy(r), dim)
        assert val.to_numpy().reshape(-1).tolist() == gv and idx.to_numpy().reshape(-1).tolist() == gi, name


def test_by_key_issues_kat(af):
    key, runs = kb.issue_2955()
    val = np.random.default_rng(0).random(256).astype(np.float32)
    ko, vo = af.sum_by_key(af.Array.from_numpy(key), af.Array.from_numpy(val), 0)
    assert ko.dims[0] == runs and vo.dims[0] == runs
    val2 = np.asfortranarray(np.random.default_rng(1).random((8, 256)).astype(np.float32))
    ko, vo = af.sum_by_key(af.Array.from_numpy(key), af.Array.from_numpy(val2), 1)
    assert ko.dims[0] == runs and vo.dims[1] == runs
    zeros, ones, n = kb.issue_3062()
    gz = af.Array.from_numpy(zeros)
    assert int(af.sum_by_key(gz, af.Array.from_numpy(ones), 0)[1].to_numpy().reshape(-1)[0]) == n
    assert int(af.count_by_key(gz, af.Array.from_numpy(ones), 0)[1].to_numpy().reshape(-1)[0]) == n
    ones2 = af.Array.from_numpy(np.asfortranarray(np.ones((2, 129), np.uint32)))
    assert af.sum_by_key(gz, ones2, 1)[1].to_numpy().reshape(-1).tolist() == [n, n]
    assert af.count_by_key(gz, ones2, 1)[1].to_numpy().reshape(-1).tolist() == [n, n]


# ---- further reduction known answers / properties of test/reduce.cpp (tests/golden/kat_reduce.py) ------------------------------
from tests.golden import kat_reduce as kr  # noqa: E402


def test_reduce_minmax_cplx_nan_kat(af):
    g = af.Array.from_numpy(np.asfortranarray(kr.CPLX_NAN))
    np.testing.assert_array_equal(af.min(g, 0).to_numpy().reshape(-1), kr.CPLX_NAN_MIN)
    np.testing.assert_array_equal(af.max(g, 0).to_numpy().reshape(-1), kr.CPLX_NAN_MAX)


def test_reduce_nan_properties_kat(af):
    a = kr.nan_vector()
    g = af.Array.from_numpy(a)
    ok = a[~np.isnan(a)]
    assert af.min_all(g) == float(ok.min()) and af.max_all(g) == float(ok.max())          # MinMax.MinMaxNaN
    assert np.isnan(af.sum_all(g))                                                         # Sum.NaN
    want = float(np.where(np.isnan(a), 0, a).astype(np.float64).sum())
    assert abs(af.sum_all(g, nanval=0.0) / a.size - want / a.size) < 1e-5
    p = np.random.default_rng(12).random(5).astype(np.float32)
    p[2] = np.nan
    gp = af.Array.from_numpy(p)
    assert np.isnan(af.product_all(gp))                                                    # Product.NaN
    assert abs(af.product_all(gp, nanval=1.0) - float(np.prod(np.where(np.isnan(p), 1, p).astype(np.float64)))) < 1e-5
    a, b = kr.count_nan_pair()
    assert af.count_all(af.Array.from_numpy(a)) == af.count_all(af.Array.from_numpy(b)) == float(np.count_nonzero(b))   # Count.NaN
    a, b = kr.anyall_nan_pair()
    ga, gb = af.Array.from_numpy(a), af.Array.from_numpy(b)
    assert af.any_true_all(gb) and af.all_true_all(gb) and af.any_true_all(ga) and not af.all_true_all(ga)            # AnyAll.NaN


def test_reduce_product_of_bools_issue_2543_kat(af):
    ones = af.Array.from_numpy(np.ones((5, 5), np.int8, order="F"), dtype=po.B8)
    assert af.product_all(ones) > 0
    x = np.asfortranarray((np.random.default_rng(13).random((5, 5)) < 0.8).astype(np.int8))
    g = af.Array.from_numpy(x, dtype=po.B8)
    assert af.product_all(g) == float(np.prod(x.astype(np.uint32)))
    assert af.all_true(g, 0).to_numpy().reshape(-1).astype(np.int64).tolist() == \
        af.product(g, 0).to_numpy().reshape(-1).astype(np.int64).tolist()


def test_reduce_nanval_issue_3255_kat(af):
    gk = af.Array.from_numpy(kr.ISSUE_3255_KEYS)
    for i in range(8):
        v = kr.ISSUE_3255_VALS.copy()
        v[i] = np.nan
        gv = af.Array.from_numpy(v)
        for fn in (af.product_by_key, af.sum_by_key):
            ko, vo = fn(gk, gv, 0, nanval=1.0)
            assert not np.isnan(vo.to_numpy()).any(), (fn.__name__, i)
            assert ko.to_numpy().reshape(-1).tolist() == [0, 1, 2, 0]


def test_reduce_max_all_indexed_kat(af):
    """MaxAll.IndexedSmall / IndexedBig / Reduce.AllSmallIndexed (test/reduce.cpp:1301-1355): whole-array max of a sub-array view"""
    rng = np.random.default_rng(14)
    for num, st, en in ((1000, 10, 900), (100000, 1000, 99000)):
        a = rng.random(num).astype(np.float32)
        g = af.Array.from_numpy(a).index(af.seq(st, en, 1))
        assert af.max_all(g) == float(a[st:en + 1].max())
    ln = 512
    r = np.asfortranarray(np.stack([np.arange(ln, dtype=np.float32)] * 2, axis=1))
    gb = af.Array.from_numpy(r).index(af.seq(0, ln // 2 - 1, 1), af.seq(0, 1, 1))
    for _ in range(50):
        assert af.max_all(gb) == ln // 2 - 1


def test_reduce_half_kat(af):
    """ReduceHalf.* (test/reduce.cpp:1375-1548): sums / products of constant f16 arrays come back as f32 (value * n, value ** n);
    min / max stay f16; count / any / all of the 1..7 ramp"""
    for value, dims, dim in kr.reduce_half_cases():
        x = np.full(dims, value, np.float16, order="F")
        g = af.Array.from_numpy(x)
        n = dims[dim if dim >= 0 else 0]
        s = af.sum(g, dim)
        assert s.dtype == po.F32 and np.all(s.to_numpy() == np.float32(value * n)), (dims, dim)
        with np.errstate(over="ignore"):
            p = np.float32(value) ** n
        if np.isfinite(p):
            q = af.product(g, dim)
            assert q.dtype == po.F32 and np.all(q.to_numpy() == p), (dims, dim)
    g = af.Array.from_numpy(kr.HALF_RAMP)
    mn, mx = af.min(g, 0), af.max(g, 0)
    assert mn.dtype == po.F16 and mx.dtype == po.F16
    assert mn.to_numpy().reshape(-1)[0] == np.float16(1) and mx.to_numpy().reshape(-1)[0] == np.float16(7)
    assert int(af.count(g, 0).to_numpy().reshape(-1)[0]) == 7
    assert bool(af.any_true(g, 0).to_numpy().reshape(-1)[0]) and bool(af.all_true(g, 0).to_numpy().reshape(-1)[0])
