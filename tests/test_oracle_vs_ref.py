"""Pins the restated oracle (oracle/oracle.cpp) to the reference's OWN CPU kernel templates
(oracle/_ref/libafref.so = /root/reference/src/backend/cpu/kernel/*.hpp compiled in place).

Randomised shapes / sub-array views / dtypes / ops; results must be BIT-identical, including
floats (both sides accumulate serially in the same order).
"""
import numpy as np
import pytest

from oracle import pyoracle as po

ALL = [po.F32, po.F64, po.C32, po.C64, po.S32, po.U32, po.S64, po.U64, po.S16, po.U16, po.S8, po.U8, po.B8, po.F16]
SHAPES = [(1,), (7,), (1000,), (5, 3), (1, 9), (33, 1, 4), (4, 5, 3, 2), (2, 1, 1, 3), (257, 3)]


def rand(rng, code, shape, nan_frac=0.0, small=False):
    n = int(np.prod(shape))
    dt = po.NP_OF[code]
    if code in (po.F32, po.F64, po.F16):
        a = (rng.integers(-8, 9, n) if small else rng.standard_normal(n) * 4).astype(dt)
    elif code in (po.C32, po.C64):
        a = (rng.integers(-4, 5, n) + 1j * rng.integers(-4, 5, n)).astype(dt)
    elif code == po.B8:
        a = rng.integers(0, 2, n).astype(dt) * rng.choice([1, 1, 2, -3], n).astype(dt)
    else:
        info = np.iinfo(dt)
        lo, hi = (max(info.min, -6), min(info.max, 6)) if small else (info.min, info.max)
        a = rng.integers(lo, hi, n, dtype=np.int64 if dt != np.uint64 else np.uint64, endpoint=True).astype(dt)
    if nan_frac > 0 and code in (po.F32, po.F64, po.F16, po.C32, po.C64):
        a[rng.random(n) < nan_frac] = np.nan
    return np.asfortranarray(a.reshape(shape, order="F"))


def views(rng, a):
    """the array itself plus a strided sub-array view with an offset (af_index with seqs)."""
    yield a
    sl = tuple(slice(int(rng.integers(0, max(1, s // 2))), s) for s in a.shape)
    v = a[sl]
    if v.size and v.size != a.size:
        yield v


def same(x, y):
    x, y = np.asarray(x), np.asarray(y)
    assert x.shape == y.shape and x.dtype == y.dtype
    assert x.tobytes() == y.tobytes()


@pytest.mark.parametrize("code", ALL)
def test_reduce_family_bitwise(port_oracle, ref_oracle, code):
    rng = np.random.default_rng(100 + code)
    for shape in SHAPES:
        base = rand(rng, code, shape, nan_frac=0.1)
        for x in views(rng, base):
            for dim in range(4):
                for name in ("sum", "product", "min", "max", "count", "all_true", "any_true"):
                    tp, p = getattr(port_oracle, name)(x, dim, dtype=code)
                    tr, r = getattr(ref_oracle, name)(x, dim, dtype=code)
                    assert tp == tr
                    same(p, r)
                tp, p = port_oracle.sum_nan(x, dim, 2.5, dtype=code)
                tr, r = ref_oracle.sum_nan(x, dim, 2.5, dtype=code)
                same(p, r)
                tp, p = port_oracle.product_nan(x, dim, -1.0, dtype=code)
                tr, r = ref_oracle.product_nan(x, dim, -1.0, dtype=code)
                same(p, r)
            for name in ("sum_all", "product_all", "min_all", "max_all", "count_all", "all_true_all", "any_true_all"):
                tp, p = getattr(port_oracle, name)(x, dtype=code)
                tr, r = getattr(ref_oracle, name)(x, dtype=code)
                assert tp == tr
                same(p, r)
            same(port_oracle.sum_nan_all(x, 1.5, dtype=code, scalar_api=False)[1],
                 ref_oracle.sum_nan_all(x, 1.5, dtype=code, scalar_api=False)[1])


@pytest.mark.parametrize("code", ALL)
def test_ireduce_bitwise(port_oracle, ref_oracle, code):
    rng = np.random.default_rng(200 + code)
    for shape in SHAPES:
        for nan_frac in (0.0, 0.3, 1.0):
            base = rand(rng, code, shape, nan_frac=nan_frac, small=True)
            for x in views(rng, base):
                for dim in range(4):
                    for name in ("imin", "imax"):
                        tp, pv, pl = getattr(port_oracle, name)(x, dim, dtype=code)
                        tr, rv, rl = getattr(ref_oracle, name)(x, dim, dtype=code)
                        same(pv, rv)
                        same(pl, rl)
                if code != po.F16:
                    for name in ("imin_all", "imax_all"):
                        pv, pl = getattr(port_oracle, name)(x, dtype=code)
                        rv, rl = getattr(ref_oracle, name)(x, dtype=code)
                        assert np.asarray(pv).tobytes() == np.asarray(rv).tobytes() and pl == rl


@pytest.mark.parametrize("code", [c for c in ALL if c != po.F16])
def test_scan_bitwise(port_oracle, ref_oracle, code):
    rng = np.random.default_rng(300 + code)
    for shape in SHAPES:
        base = rand(rng, code, shape, nan_frac=0.1)
        for x in views(rng, base):
            for dim in range(4):
                same(port_oracle.accum(x, dim, dtype=code)[1], ref_oracle.accum(x, dim, dtype=code)[1])
                for op in range(4):
                    for incl in (True, False):
                        tp, p = port_oracle.scan(x, dim, op, incl, dtype=code)
                        tr, r = ref_oracle.scan(x, dim, op, incl, dtype=code)
                        assert tp == tr
                        same(p, r)


@pytest.mark.parametrize("code", [c for c in ALL if c != po.F16])
def test_scan_by_key_bitwise(port_oracle, ref_oracle, code):
    rng = np.random.default_rng(400 + code)
    for shape in SHAPES:
        base = rand(rng, code, shape, nan_frac=0.1, small=True)
        for kcode in (po.S32, po.U32, po.S64, po.U64):
            keys = np.asfortranarray(rng.integers(0, 3, base.shape).astype(po.NP_OF[kcode]))
            sl = tuple(slice(int(rng.integers(0, max(1, s // 2))), s) for s in base.shape)
            for x, k in ((base, keys), (base[sl], keys[sl])):
                if x.size == 0:
                    continue
                for dim in range(4):
                    for op in range(4):
                        for incl in (True, False):
                            tp, p = port_oracle.scan_by_key(k, x, dim, op, incl, dtype=code, key_dtype=kcode)
                            tr, r = ref_oracle.scan_by_key(k, x, dim, op, incl, dtype=code, key_dtype=kcode)
                            assert tp == tr
                            same(p, r)


@pytest.mark.parametrize("code", [c for c in ALL if c != po.F16])
def test_where_bitwise(port_oracle, ref_oracle, code):
    rng = np.random.default_rng(500 + code)
    for shape in SHAPES:
        base = rand(rng, code, shape, nan_frac=0.1, small=True)
        for x in views(rng, base):
            same(port_oracle.where(x, dtype=code), ref_oracle.where(x, dtype=code))


@pytest.mark.parametrize("code", ALL)
def test_reduce_by_key_and_ragged_bitwise(port_oracle, ref_oracle, code):
    """the restated by-key reductions / ragged max against the reference's own
    kernel::n_reduced_keys + kernel::reduce_dim_by_key and kernel::ireduce_dim with ragged lengths"""
    rng = np.random.default_rng(900 + code)
    for shape in [(2,), (9,), (300,), (6, 4), (3, 7, 2), (2, 3, 4, 5), (130, 3)]:
        base = rand(rng, code, shape, nan_frac=0.1, small=True)
        for dim in range(len(shape)):
            n = shape[dim]
            if n < 2:
                continue
            keys = np.repeat(rng.integers(0, 3, n), rng.integers(1, 5, n))[:n].astype(np.int32 if dim % 2 else np.uint32)
            for name in ("sum_by_key", "product_by_key", "min_by_key", "max_by_key", "count_by_key", "all_true_by_key",
                         "any_true_by_key"):
                p = getattr(port_oracle, name)(keys, base, dim, dtype=code)
                r = getattr(ref_oracle, name)(keys, base, dim, dtype=code)
                assert p[0] == r[0] and p[2] == r[2]
                same(p[1], r[1])
                same(p[3], r[3])
            p = port_oracle.sum_by_key(keys, base, dim, dtype=code, nanval=3.0)
            r = ref_oracle.sum_by_key(keys, base, dim, dtype=code, nanval=3.0)
            same(p[3], r[3])
            od = list(po.dims4(shape))
            od[dim] = 1
            rl = rng.integers(0, n + 2, od).astype(np.uint32)
            p = port_oracle.max_ragged(base, rl, dim, dtype=code)
            r = ref_oracle.max_ragged(base, rl, dim, dtype=code)
            same(p[1], r[1])
            same(p[2], r[2])
