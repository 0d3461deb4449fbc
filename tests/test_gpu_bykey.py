"""`-m gpu`: SURVEY.md section 8(f)-1 -- af_*_by_key reductions and af_max_ragged -- and the array-handling calls the
C++ wrapper needs around the path (af_cast, af_moddims, af_flat), through the C-ABI against the oracle.
Bit-exact for integers, min / max, count / any / all and for float sums of integer-valued data."""
import numpy as np
import pytest

from oracle import pyoracle as po
from tests.gpu_util import gpu_available, rand_data, same_bits

pytestmark = pytest.mark.gpu

ALL = [po.F32, po.F64, po.C32, po.C64, po.S32, po.U32, po.S64, po.U64, po.S16, po.U16, po.S8, po.U8, po.B8, po.F16]
OPS = ("sum_by_key", "product_by_key", "min_by_key", "max_by_key", "count_by_key", "all_true_by_key", "any_true_by_key")


@pytest.fixture(scope="module")
def af():
    if not gpu_available():
        pytest.skip("no CUDA device")
    import arrayfire_b200
    return arrayfire_b200


def d4(a):
    return a.reshape(po.dims4(a.shape), order="F")


def make_keys(rng, n, kind, dt):
    if kind == "same":
        return np.zeros(n, dt)
    if kind == "distinct":
        return np.arange(n).astype(dt)
    if kind == "long":
        return np.repeat(rng.integers(0, 3, n // 3000 + 2), rng.integers(2500, 9000, n // 3000 + 2))[:n].astype(dt)
    return np.repeat(rng.integers(0, 5, n), rng.integers(1, 9, n))[:n].astype(dt)


@pytest.mark.parametrize("code", ALL)
def test_reduce_by_key(af, port_oracle, code):
    orc = port_oracle
    rng = np.random.default_rng(11000 + code)
    shapes = [((2,), 0), ((7,), 0), ((2049,), 0), ((5000,), 0), ((70001,), 0), ((4100, 3), 0), ((3, 4100), 1), ((5, 3, 300), 2),
              ((2, 2, 2, 130), 3), ((2500, 2, 2), 0)]
    for shape, dim in shapes:
        n = shape[dim]
        x = rand_data(rng, code, shape, nan_frac=0.05, small=True, exact=True)
        gx = af.Array.from_numpy(x, dtype=code)
        for kind in ("short", "same", "distinct", "long"):
            if kind != "short" and len(shape) > 2:
                continue
            kdt = np.int32 if (n + code) % 2 else np.uint32
            keys = make_keys(rng, n, kind, kdt)
            gk = af.Array.from_numpy(keys)
            for name in OPS:
                if name == "product_by_key" and code in (po.F32, po.F64, po.C32, po.C64, po.F16):
                    continue   # float products are not exactly representable: order-dependent rounding (see below)
                tk, wk, tv, wv = getattr(orc, name)(keys, x, dim, dtype=code)
                ko, vo = getattr(af, name)(gk, gx, dim)
                assert ko.dtype == tk and vo.dtype == tv, (name, shape)
                same_bits(ko.to_numpy().reshape(-1), wk, "%s keys %s %s" % (name, shape, kind))
                same_bits(vo.to_numpy(), d4(wv), "%s vals %s dim%d %s" % (name, shape, dim, kind))
            tk, wk, tv, wv = orc.sum_by_key(keys, x, dim, dtype=code, nanval=2.0)
            ko, vo = af.sum_by_key(gk, gx, dim, nanval=2.0)
            same_bits(vo.to_numpy(), d4(wv), "sum_by_key_nan %s %s" % (shape, kind))
            if code in (po.F32, po.F64, po.C32, po.C64, po.F16) and kind == "short":
                # float products: runs of at most a few elements of +-1 / +-2 stay exact in any order
                xs = np.where(x == x, np.sign(x.real) * np.minimum(np.abs(x.real), 2), x).astype(x.dtype)
                tk, wk, tv, wv = orc.product_by_key(keys, xs, dim, dtype=code, nanval=1.0)
                ko, vo = af.product_by_key(gk, af.Array.from_numpy(np.asfortranarray(xs), dtype=code), dim, nanval=1.0)
                same_bits(vo.to_numpy(), d4(wv), "product_by_key_nan %s" % (shape,))


def test_reduce_by_key_large_and_views(af, port_oracle):
    orc = port_oracle
    rng = np.random.default_rng(11500)
    n = (1 << 22) + 77
    x = rng.integers(-15, 15, n).astype(np.int32)
    keys = np.repeat(np.arange(n // 256 + 2, dtype=np.int32), rng.integers(1, 1025, n // 256 + 2))[:n]
    keys[1 << 20:(1 << 20) + 300000] = 7          # one run spanning ~150 tiles
    gk, gx = af.Array.from_numpy(keys), af.Array.from_numpy(x)
    for name in ("sum_by_key", "max_by_key", "count_by_key"):
        tk, wk, tv, wv = getattr(orc, name)(keys, x, 0)
        ko, vo = getattr(af, name)(gk, gx, 0)
        same_bits(ko.to_numpy().reshape(-1), wk, name)
        same_bits(vo.to_numpy(), d4(wv), name)
    # strided sub-array views of keys and values
    base = np.asfortranarray(rng.integers(-4, 5, (9000, 5)).astype(np.float32))
    kbase = make_keys(rng, 9000, "short", np.int32)
    gb = af.Array.from_numpy(base).index(af.seq(3, 8002, 1), af.seq(1, 3, 1))
    gkv = af.Array.from_numpy(kbase).index(af.seq(3, 8002, 1))
    tk, wk, tv, wv = orc.sum_by_key(kbase[3:8003], base[3:8003, 1:4], 0)
    ko, vo = af.sum_by_key(gkv, gb, 0)
    same_bits(ko.to_numpy().reshape(-1), wk, "view keys")
    same_bits(vo.to_numpy(), d4(wv), "view vals")
    # f32 accuracy on generic data: within 1e-5 of an f64 truth
    xf = rng.random(n).astype(np.float32)
    ko, vo = af.sum_by_key(gk, af.Array.from_numpy(xf), 0)
    heads = np.flatnonzero(np.concatenate([[True], keys[1:] != keys[:-1]]))
    truth = np.add.reduceat(xf.astype(np.float64), heads)
    np.testing.assert_allclose(vo.to_numpy().reshape(-1).astype(np.float64), truth, rtol=1e-5)


def test_reduce_by_key_tile_multiples(af, port_oracle):
    """key counts that are exact multiples of the 2048-key tile (the end-of-line marker of the run-head bits then sits in
    the extra byte behind the last tile), runs ending exactly on tile boundaries, several value columns"""
    orc = port_oracle
    rng = np.random.default_rng(11700)
    for n in (2048, 4096, 6144, 16384):
        for kind in ("short", "same", "distinct", "tile"):
            keys = make_keys(rng, n, kind, np.int32) if kind != "tile" else (np.arange(n) // 2048).astype(np.int32)
            x = np.asfortranarray(rng.integers(-9, 10, (n, 3)).astype(np.int32))
            gk, gx = af.Array.from_numpy(keys), af.Array.from_numpy(x)
            for name in ("sum_by_key", "min_by_key", "count_by_key"):
                tk, wk, tv, wv = getattr(orc, name)(keys, x, 0)
                ko, vo = getattr(af, name)(gk, gx, 0)
                same_bits(ko.to_numpy().reshape(-1), wk, "%s keys n=%d %s" % (name, n, kind))
                same_bits(vo.to_numpy(), d4(wv), "%s vals n=%d %s" % (name, n, kind))


def test_by_key_argument_errors(af):
    x = af.Array.from_numpy(np.ones((8, 3), np.float32, order="F"))
    k = af.Array.from_numpy(np.zeros(8, np.int32))
    with pytest.raises(af.AfError) as e:
        af.sum_by_key(af.Array.from_numpy(np.zeros(7, np.int32)), x, 0)
    assert e.value.code == 202
    with pytest.raises(af.AfError) as e:
        af.sum_by_key(af.Array.from_numpy(np.zeros(8, np.float32)), x, 0)
    assert e.value.code == 204
    with pytest.raises(af.AfError) as e:
        af.sum_by_key(af.Array.from_numpy(np.zeros((4, 2), np.int32, order="F")), x, 0)
    assert e.value.code == 202     # keys must be a vector
    ko, vo = af.sum_by_key(k, x, 0)
    assert ko.dims == (1, 1, 1, 1) and vo.dims == (1, 3, 1, 1)


@pytest.mark.parametrize("code", ALL)
def test_max_ragged(af, port_oracle, code):
    orc = port_oracle
    rng = np.random.default_rng(12000 + code)
    for shape in [(7,), (1000,), (5, 3), (33, 1, 4), (4, 5, 3, 2), (257, 70), (40, 300)]:
        x = rand_data(rng, code, shape, nan_frac=0.2, small=True)
        g = af.Array.from_numpy(x, dtype=code)
        for dim in range(len(shape)):
            od = list(po.dims4(shape))
            od[dim] = 1
            rl = np.asfortranarray(rng.integers(0, shape[dim] + 3, od).astype(np.uint32))
            t, wv, wl = orc.max_ragged(x, rl, dim, dtype=code)
            gv, gl = af.max_ragged(g, af.Array.from_numpy(rl), dim)
            assert gv.dtype == t and gl.dtype == po.U32
            same_bits(gl.to_numpy(), d4(wl), "ragged idx %s dim%d" % (shape, dim))
            same_bits(gv.to_numpy(), d4(wv), "ragged val %s dim%d" % (shape, dim))


def test_cast_moddims_flat(af):
    rng = np.random.default_rng(12500)
    x = np.asfortranarray((rng.standard_normal((37, 11)) * 50).astype(np.float32))
    g = af.Array.from_numpy(x)
    for code, npdt in ((po.S32, np.int32), (po.F64, np.float64), (po.U8, None), (po.S16, np.int16), (po.C32, np.complex64),
                       (po.F16, np.float16), (po.S64, np.int64), (po.B8, None)):
        r = af.cast(g, code)
        assert r.dtype == code and r.dims == (37, 11, 1, 1)
        got = r.to_numpy()
        if code == po.B8:
            np.testing.assert_array_equal(got.reshape(37, 11), (x != 0).astype(np.int8))
        elif code == po.U8:
            xs = np.clip(np.trunc(x), 0, 255)      # cvt saturates on the device; keep to the in-range part
            ok = (x >= 0) & (x < 256)
            np.testing.assert_array_equal(got.reshape(37, 11)[ok], xs[ok].astype(np.uint8))
        else:
            np.testing.assert_array_equal(got.reshape(37, 11), x.astype(npdt))
    assert af.cast(g, po.F32).dtype == po.F32
    c = af.cast(g, po.C64)
    with pytest.raises(af.AfError) as e:
        af.cast(c, po.F32)
    assert e.value.code == 204
    # views: a sub-array is cast element-wise in view order
    v = g.index(af.seq(2, 30, 1), af.seq(1, 9, 1))
    np.testing.assert_array_equal(af.cast(v, po.S32).to_numpy().reshape(29, 9), x[2:31, 1:10].astype(np.int32))
    # moddims / flat share the buffer of a linear array and copy a strided view
    m = af.moddims(g, (11, 37))
    assert m.dims == (11, 37, 1, 1)
    np.testing.assert_array_equal(m.to_numpy().reshape(-1, order="F"), x.reshape(-1, order="F"))
    f = af.flat(v)
    assert f.dims == (29 * 9, 1, 1, 1)
    np.testing.assert_array_equal(f.to_numpy().reshape(-1), x[2:31, 1:10].reshape(-1, order="F"))
    with pytest.raises(af.AfError) as e:
        af.moddims(g, (5, 5))
    assert e.value.code == 203
