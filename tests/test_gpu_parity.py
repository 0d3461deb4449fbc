"""`-m gpu` parity tests: the CUDA path, called through the C-ABI of libafb200.so, against the oracle.

Bit-exact for integer/byte/index work, min/max values, ireduce indices, count/where output and integer
scans.  Floating-point sums/scans are checked (a) BIT-EXACT on inputs whose every partial sum is exactly
representable (any association order then gives the same bits), and (b) within 1e-5 (f32) / 1e-12 (f64)
relative of an f64/long-double truth on generic data -- the tolerances BASELINE.json's north_star states.
"""
import numpy as np
import pytest

from oracle import pyoracle as po
from tests.gpu_util import gpu_available, rand_data, random_subview, same_bits, to_gpu

pytestmark = pytest.mark.gpu

ALL = [po.F32, po.F64, po.C32, po.C64, po.S32, po.U32, po.S64, po.U64, po.S16, po.U16, po.S8, po.U8, po.B8, po.F16]
NOHALF = [c for c in ALL if c != po.F16]
IS_FLOAT = {po.F32, po.F64, po.C32, po.C64, po.F16}

# shapes chosen to hit every kernel regime (see arrayfire_b200/csrc/kernels/*_host.cuh)
SHAPES_SMALL = [(1,), (7,), (1000,), (5, 3), (1, 9), (33, 1, 4), (4, 5, 3, 2), (2, 1, 1, 3), (257, 70), (40, 300)]
SHAPES_BIG = [(70001,), (5000, 3), (3000, 100), (200003, 3), (130, 1000), (4, 50001), (96, 64, 2, 3), (1 << 20,), (2500, 4800)]


@pytest.fixture(scope="module")
def af():
    if not gpu_available():
        pytest.skip("no CUDA device")
    import arrayfire_b200
    return arrayfire_b200


@pytest.fixture(scope="module")
def orc(port_oracle):
    return port_oracle


def views_of(rng, base):
    yield base, None
    if base.size > 1:
        sl = random_subview(rng, base.shape)
        if base[sl].size:
            yield base[sl], sl


def d4(a):
    return a.reshape(po.dims4(a.shape), order="F")


@pytest.mark.parametrize("code", ALL)
def test_reduce_dimwise_small(af, orc, code):
    rng = np.random.default_rng(1000 + code)
    for shape in SHAPES_SMALL:
        base = rand_data(rng, code, shape, nan_frac=0.1, exact=True)
        for x, sl in views_of(rng, base):
            g = to_gpu(af, base, sl, dtype=code)
            for dim in range(4):
                for name in ("sum", "product", "min", "max", "count", "all_true", "any_true"):
                    if name == "product" and code in IS_FLOAT:
                        continue  # long float products are not exactly representable; see test_float_sum_accuracy
                    t, want = getattr(orc, name)(x, dim, dtype=code)
                    got = getattr(af, name)(g, dim)
                    assert got.dtype == t, (name, shape, dim)
                    same_bits(got.to_numpy(), d4(want), "%s %s dim%d view=%s" % (name, shape, dim, sl))
                t, want = orc.sum_nan(x, dim, 2.0, dtype=code)
                same_bits(af.sum(g, dim, nanval=2.0).to_numpy(), d4(want), "sum_nan %s dim%d" % (shape, dim))


@pytest.mark.parametrize("code", [po.F32, po.S32, po.U8, po.F64, po.S16, po.C32, po.B8, po.U64])
def test_reduce_dimwise_big(af, orc, code):
    rng = np.random.default_rng(2000 + code)
    for shape in SHAPES_BIG:
        base = rand_data(rng, code, shape, nan_frac=0.01, exact=True)
        for x, sl in views_of(rng, base):
            g = to_gpu(af, base, sl, dtype=code)
            for dim in range(len(shape)):
                for name in ("sum", "min", "max", "count", "any_true", "all_true"):
                    t, want = getattr(orc, name)(x, dim, dtype=code)
                    got = getattr(af, name)(g, dim)
                    assert got.dtype == t
                    same_bits(got.to_numpy(), d4(want), "%s %s dim%d view=%s" % (name, shape, dim, sl))


@pytest.mark.parametrize("code", ALL)
def test_reduce_all(af, orc, code):
    rng = np.random.default_rng(3000 + code)
    for shape in SHAPES_SMALL + [(70001,), (300, 1000), (1 << 21,)]:
        base = rand_data(rng, code, shape, nan_frac=0.0, exact=True)
        for x, sl in views_of(rng, base):
            g = to_gpu(af, base, sl, dtype=code)
            for name in ("sum_all", "min_all", "max_all", "count_all", "all_true_all", "any_true_all"):
                t, want = getattr(orc, name)(x, dtype=code)
                got = getattr(af, name)(g)
                w = want[0]
                if code in (po.C32, po.C64) and name in ("sum_all", "min_all", "max_all"):
                    assert got == complex(w), (name, shape, got, w)
                else:
                    assert got == float(np.real(w)), (name, shape, got, w)
            # *_all_array keeps the promoted dtype (no trip through double)
            t, want = orc.sum_all(x, dtype=code, scalar_api=False)
            arr = af.sum_all_array(g)
            assert arr.dtype == t and arr.dims == (1, 1, 1, 1)
            same_bits(arr.to_numpy().reshape(1), want, "sum_all_array %s" % (shape,))


@pytest.mark.parametrize("code", [po.F32, po.F64, po.C32, po.F16])
def test_float_sum_accuracy(af, orc, code):
    """generic data: within 1e-5 (f32, f16->f32) / 1e-12 (f64) relative of a high-precision truth."""
    rng = np.random.default_rng(4000 + code)
    n = 1 << 22
    if code == po.C32:
        x = (rng.random(n) + 1j * rng.random(n)).astype(np.complex64)
        truth = complex(np.sum(x.real.astype(np.float64)), np.sum(x.imag.astype(np.float64)))
    else:
        x = rng.random(n).astype(po.NP_OF[code])
        truth = float(np.sum(x.astype(np.longdouble)))
    got = af.sum_all(af.Array.from_numpy(x, dtype=code))
    tol = 1e-12 if code == po.F64 else 1e-5
    assert abs(got - truth) <= tol * abs(truth), (got, truth)
    m = np.asfortranarray(x.reshape((1 << 11, 1 << 11), order="F"))
    for dim in (0, 1):
        got = af.sum(af.Array.from_numpy(m, dtype=code), dim).to_numpy().reshape(-1)
        want = m.astype(np.complex128 if code == po.C32 else np.float64).sum(axis=dim)
        np.testing.assert_allclose(got.astype(want.dtype), want, rtol=tol)
        sc = af.accum(af.Array.from_numpy(m, dtype=code), dim).to_numpy() if code != po.F16 else None
        if sc is not None:
            want = np.cumsum(m.astype(np.complex128 if code == po.C32 else np.float64), axis=dim)
            np.testing.assert_allclose(sc.reshape(m.shape, order="F").astype(want.dtype), want, rtol=tol * 4)
    if code == po.C32:
        p = (0.5 + rng.random(64) + 0.1j).astype(np.complex64)
    else:
        p = (0.5 + rng.random(64)).astype(po.NP_OF[code])
    got = af.product_all(af.Array.from_numpy(p, dtype=code))
    want = np.prod(p.astype(np.complex128 if code == po.C32 else np.float64))
    assert abs(got - want) <= 1e-4 * abs(want)


@pytest.mark.parametrize("code", ALL)
def test_ireduce(af, orc, code):
    rng = np.random.default_rng(5000 + code)
    for shape in SHAPES_SMALL + [(70001,), (5000, 3), (130, 1000), (4, 50001)]:
        for nan_frac in (0.0, 0.3):
            if nan_frac and code not in IS_FLOAT:
                continue
            base = rand_data(rng, code, shape, nan_frac=nan_frac, small=True)
            for x, sl in views_of(rng, base):
                g = to_gpu(af, base, sl, dtype=code)
                for dim in range(4):
                    for name in ("imin", "imax"):
                        t, wv, wl = getattr(orc, name)(x, dim, dtype=code)
                        gv, gl = getattr(af, name)(g, dim)
                        assert gv.dtype == t and gl.dtype == po.U32
                        same_bits(gl.to_numpy(), d4(wl), "%s idx %s dim%d view=%s" % (name, shape, dim, sl))
                        same_bits(gv.to_numpy(), d4(wv), "%s val %s dim%d view=%s" % (name, shape, dim, sl))
                if code != po.F16:
                    for name in ("imin_all", "imax_all"):
                        wv, wl = getattr(orc, name)(x, dtype=code)
                        gv, gl = getattr(af, name)(g)
                        assert gl == wl, (name, shape, sl, gl, wl)
                        assert gv == (complex(wv) if code in (po.C32, po.C64) else float(wv)), (name, shape)


def test_ireduce_all_nan_rows(af, orc):
    """all-NaN lines return (identity, 0); a NaN seed with later values behaves like the serial loop"""
    x = np.full((5, 4), np.nan, np.float32, order="F")
    x[2, 1] = 3.0
    x[4, 2] = np.inf
    x[0, 3] = -np.inf
    g = af.Array.from_numpy(x)
    for name in ("imin", "imax"):
        for dim in (0, 1):
            t, wv, wl = getattr(orc, name)(x, dim)
            gv, gl = getattr(af, name)(g, dim)
            same_bits(gl.to_numpy(), d4(wl), name)
            same_bits(gv.to_numpy(), d4(wv), name)


@pytest.mark.parametrize("code", NOHALF)
def test_scan(af, orc, code):
    rng = np.random.default_rng(6000 + code)
    shapes = SHAPES_SMALL + [(70001,), (5000, 3), (130, 1000), (3, 20000), (2, 3, 5000)]
    for shape in shapes:
        base = rand_data(rng, code, shape, nan_frac=0.05, exact=True)
        for x, sl in views_of(rng, base):
            g = to_gpu(af, base, sl, dtype=code)
            for dim in range(4):
                t, want = orc.accum(x, dim, dtype=code)
                got = af.accum(g, dim)
                assert got.dtype == t
                same_bits(got.to_numpy(), d4(want), "accum %s dim%d view=%s" % (shape, dim, sl))
                for op in range(4):
                    if op == po.AF_BINARY_MUL and (code in IS_FLOAT):
                        continue
                    for incl in (True, False):
                        t, want = orc.scan(x, dim, op, incl, dtype=code)
                        got = af.scan(g, dim, op, incl)
                        assert got.dtype == t
                        same_bits(got.to_numpy(), d4(want), "scan op%d incl%d %s dim%d view=%s" % (op, incl, shape, dim, sl))


@pytest.mark.parametrize("code", NOHALF)
def test_scan_by_key(af, orc, code):
    rng = np.random.default_rng(7000 + code)
    shapes = [(1,), (7,), (1000,), (5, 3), (33, 1, 4), (4, 5, 3, 2), (70001,), (5000, 3), (130, 1000), (3, 20000)]
    for shape in shapes:
        base = rand_data(rng, code, shape, nan_frac=0.05, exact=True)
        n = int(np.prod(shape))
        for kcode in (po.S32, po.U64):
            # runs of random length so segments cross vector / warp / tile boundaries
            runs = rng.integers(1, 40 if n < 5000 else 3000, n)
            keys = np.repeat(rng.integers(0, 5, n), runs)[:n].astype(po.NP_OF[kcode])
            keys = np.asfortranarray(keys.reshape(shape, order="F"))
            sl = random_subview(rng, shape) if n > 1 else None
            for xs in (None, sl):
                if xs is None and sl is None and False:
                    continue
                x = base if xs is None else base[xs]
                k = keys if xs is None else keys[xs]
                if x.size == 0:
                    continue
                g = to_gpu(af, base, xs, dtype=code)
                gk = to_gpu(af, keys, xs, dtype=kcode)
                for dim in range(min(len(shape), 2) + 1):
                    for op in (0, 2, 3):
                        for incl in (True, False):
                            t, want = orc.scan_by_key(k, x, dim, op, incl, dtype=code, key_dtype=kcode)
                            got = af.scan_by_key(gk, g, dim, op, incl)
                            assert got.dtype == t
                            same_bits(got.to_numpy(), d4(want), "sbk op%d incl%d %s dim%d view=%s" % (op, incl, shape, dim, xs))
                if sl is None:
                    break


@pytest.mark.parametrize("code", NOHALF)
def test_where(af, orc, code):
    rng = np.random.default_rng(8000 + code)
    for shape in SHAPES_SMALL + [(70001,), (300, 1000), (1 << 21,)]:
        for density in (0.5, 0.01, 0.0, 1.0):
            base = rand_data(rng, code, shape, nan_frac=0.0, small=True)
            mask = rng.random(base.shape) < density
            one = np.ones((), base.dtype)
            base = np.asfortranarray(np.where(mask, np.where(base == 0, one, base), 0 * one).astype(base.dtype))
            for x, sl in views_of(rng, base):
                g = to_gpu(af, base, sl, dtype=code)
                want = orc.where(x, dtype=code)
                got = af.where(g)
                assert got.dtype == po.U32 and got.dims == (want.size, 1, 1, 1), (shape, density, got.dims, want.size)
                same_bits(got.to_numpy().reshape(-1), want, "where %s p=%g view=%s" % (shape, density, sl))
                assert af.count_all(g) == want.size


def test_where_nan_counts_as_nonzero(af, orc):
    x = np.array([0, np.nan, 1, 0, -0.0, np.inf], np.float32)
    got = af.where(af.Array.from_numpy(x)).to_numpy().reshape(-1)
    assert got.tolist() == orc.where(x).tolist() == [1, 2, 5]


@pytest.mark.parametrize("code", ALL)
def test_nan_variants_and_all_array_twins(af, orc, code):
    """af_product_nan, af_sum_nan_all / af_product_nan_all (scalar and _array forms) and every *_all_array twin"""
    rng = np.random.default_rng(8500 + code)
    for shape in [(7,), (1000,), (5, 3), (33, 1, 4), (70001,), (300, 1000)]:
        small = rand_data(rng, code, (11,) + shape[1:], nan_frac=0.2, small=True) if len(shape) > 1 else rand_data(rng, code, (11,), nan_frac=0.2, small=True)
        base = rand_data(rng, code, shape, nan_frac=0.05, exact=True)
        for x, sl in views_of(rng, base):
            g = to_gpu(af, base, sl, dtype=code)
            pairs = [("sum_all_array", "sum_all", {}), ("min_all_array", "min_all", {}), ("max_all_array", "max_all", {}),
                     ("count_all_array", "count_all", {}), ("all_true_all_array", "all_true_all", {}),
                     ("any_true_all_array", "any_true_all", {})]
            for gname, oname, kw in pairs:
                if oname == "sum_all":
                    t, want = orc.sum_all(x, dtype=code, scalar_api=False)
                else:
                    t, want = getattr(orc, oname)(x, dtype=code)
                arr = getattr(af, gname)(g)
                assert arr.dtype == t and arr.dims == (1, 1, 1, 1), (gname, shape)
                same_bits(arr.to_numpy().reshape(1), want, "%s %s" % (gname, shape))
            t, want = orc.sum_nan_all(x, 2.0, dtype=code, scalar_api=False)
            same_bits(af.sum_all_array(g, nanval=2.0).to_numpy().reshape(1), want, "sum_nan_all_array %s" % (shape,))
            t, want = orc.sum_nan_all(x, 2.0, dtype=code)
            got = af.sum_all(g, nanval=2.0)
            assert got == (complex(want[0]) if code in (po.C32, po.C64) else float(want[0])) or (got != got and want[0] != want[0])
        # products: integer types wrap identically in any order; float types only on a 9-element line of tiny
        # integers, where every partial product is exactly representable
        if code in IS_FLOAT:
            small = small.reshape(-1, order="F")[:9].copy()
            small = np.where(small == small, np.clip(small.real, -2, 2), small).astype(small.dtype)
        gs = af.Array.from_numpy(small, dtype=code)
        for dim in range(small.ndim):
            t, want = orc.product_nan(small, dim, 1.0, dtype=code)
            got = af.product(gs, dim, nanval=1.0)
            assert got.dtype == t
            same_bits(got.to_numpy(), d4(want), "product_nan %s dim%d" % (small.shape, dim))
            t, want = orc.product(small, dim, dtype=code)
            same_bits(af.product(gs, dim).to_numpy(), d4(want), "product %s dim%d" % (small.shape, dim))
        line = small.reshape(-1, order="F")[:9].copy()
        gl = af.Array.from_numpy(line, dtype=code)
        t, want = orc.product_nan_all(line, 1.0, dtype=code, scalar_api=False)
        same_bits(af.product_all_array(gl, nanval=1.0).to_numpy().reshape(1), want, "product_nan_all_array")
        t, want = orc.product_all(line, dtype=code, scalar_api=False)
        same_bits(af.product_all_array(gl).to_numpy().reshape(1), want, "product_all_array")
        t, want = orc.product_nan_all(line, 1.0, dtype=code)
        got = af.product_all(gl, nanval=1.0)
        w0 = complex(want[0]) if code in (po.C32, po.C64) else float(want[0])
        assert got == w0 or (got != got and w0 != w0), ("product_nan_all", got, want)


def test_lifecycle_and_memory_calls(af):
    """af_constant / af_write_array / af_get_scalar / af_retain_array / af_copy_array / af_alloc_device + af_device_array /
    af_lock_array + af_get_device_ptr / af_device_mem_info / af_device_gc (src/api/c/array.cpp, memory.cpp, data.cpp)"""
    import ctypes
    lib = af.lib
    vp = ctypes.c_void_p

    def mem_info():
        v = [ctypes.c_size_t() for _ in range(4)]
        af._check(lib.af_device_mem_info(*[ctypes.byref(x) for x in v]))
        return tuple(x.value for x in v)   # alloc_bytes, alloc_buffers, lock_bytes, lock_buffers

    af.device_gc()
    base = mem_info()
    c = af.constant(2.5, (1000, 3), af.f32)
    assert c.dims == (1000, 3, 1, 1) and af.sum_all(c) == 7500.0
    assert c.scalar() == np.float32(2.5)
    now = mem_info()
    assert now[2] >= base[2] + 12000 and now[3] == base[3] + 1            # one more buffer in use
    # af_write_array from the host, then read back
    host = np.arange(3000, dtype=np.float32)
    af._check(lib.af_write_array(c.handle, vp(host.ctypes.data), ctypes.c_size_t(host.nbytes), ctypes.c_int(1)))   # afHost
    np.testing.assert_array_equal(c.to_numpy().reshape(-1, order="F"), host)
    # retain shares the buffer (use count 2), copy does not
    r = c.retain()
    cnt = ctypes.c_int()
    af._check(lib.af_get_data_ref_count(ctypes.byref(cnt), c.handle))
    assert cnt.value == 2
    cp = c.copy()
    af._check(lib.af_get_data_ref_count(ctypes.byref(cnt), cp.handle))
    assert cnt.value == 1
    del r
    af._check(lib.af_get_data_ref_count(ctypes.byref(cnt), c.handle))
    assert cnt.value == 1
    # af_write_array device -> device
    af._check(lib.af_write_array(cp.handle, vp(c.device_ptr()), ctypes.c_size_t(host.nbytes), ctypes.c_int(0)))   # afDevice
    c.unlock()
    np.testing.assert_array_equal(cp.to_numpy().reshape(-1, order="F"), host)
    # af_alloc_device + af_device_array: the library adopts the pointer (and takes it back into the pool on release)
    p = vp()
    af._check(lib.af_alloc_device(ctypes.byref(p), ctypes.c_longlong(4096)))
    assert p.value
    h = vp()
    dims = (ctypes.c_longlong * 1)(1024)
    af._check(lib.af_device_array(ctypes.byref(h), p, 1, dims, af.f32))
    adopted = af.Array(h)
    af._check(lib.af_write_array(adopted.handle, vp(host.ctypes.data), ctypes.c_size_t(4096), ctypes.c_int(1)))
    assert af.sum_all(adopted) == float(host[:1024].sum())
    # af_get_device_ptr locks: the buffer survives the release of its last handle until the user frees it
    q = adopted.device_ptr()
    assert q == p.value
    locked = ctypes.c_bool()
    af._check(lib.af_is_locked_array(ctypes.byref(locked), adopted.handle))
    assert locked.value
    del adopted
    again = vp()
    af._check(lib.af_device_array(ctypes.byref(again), vp(q), 1, dims, af.f32))
    back = af.Array(again)
    assert af.sum_all(back) == float(host[:1024].sum())       # the data is still there
    del back
    # gc returns cached (unlocked) buffers to the driver
    del c, cp
    af.device_gc()
    end = mem_info()
    assert end[2] <= base[2] + 4096 and end[0] <= now[0]
    # af_get_scalar of a view reads the first element of the VIEW
    x = af.Array.from_numpy(np.arange(20, dtype=np.int32))
    assert int(x.index(af.seq(7, 12, 1)).scalar()) == 7
    # step size getter / setter
    step = ctypes.c_size_t()
    af._check(lib.af_get_mem_step_size(ctypes.byref(step)))
    assert step.value == 1024


def test_argument_errors(af):
    a = af.Array.from_numpy(np.ones((4, 3), np.float32, order="F"))
    for fn in (af.sum, af.min, af.count, af.accum):
        with pytest.raises(af.AfError) as e:
            fn(a, 4)
        assert e.value.code == 202  # AF_ERR_ARG
    with pytest.raises(af.AfError) as e:
        af.scan(a, 0, 7, True)
    assert e.value.code == 202
    k = af.Array.from_numpy(np.ones((4, 2), np.int32, order="F"))
    with pytest.raises(af.AfError) as e:
        af.scan_by_key(k, a, 0)
    assert e.value.code == 202  # key dims must match
    kf = af.Array.from_numpy(np.ones((4, 3), np.float32, order="F"))
    with pytest.raises(af.AfError) as e:
        af.scan_by_key(kf, a, 0)
    assert e.value.code == 204  # AF_ERR_TYPE: float keys
    h = af.Array.from_numpy(np.ones(4, np.float16))
    with pytest.raises(af.AfError) as e:
        af.where(h)
    assert e.value.code == 204
    with pytest.raises(af.AfError) as e:
        af.imax_all(h)
    assert e.value.code == 204
    # dim >= ndims returns the input unpromoted
    u = af.Array.from_numpy(np.arange(6, dtype=np.uint8).reshape((2, 3), order="F"))
    r = af.sum(u, 2)
    assert r.dtype == po.U8 and r.dims == (2, 3, 1, 1)
    # empty input
    e0 = af.Array.from_numpy(np.zeros((0,), np.float32))
    assert af.where(e0).elements == 0 and af.sum_all(e0) == 0.0 and af.product_all(e0) == 1.0
    assert af.sum(e0, 0).elements == 0


# ---------------------------------------------------------------------------------------------------------
# the software-pipelined TMA kernels (pipe_scan_kernel / pipe_where_kernel): 4-byte values, long aligned columns
# ---------------------------------------------------------------------------------------------------------
PIPE_TILE, PIPE_TILE_KEYED = 12 * 32 * 20, 16 * 32 * 12


def _keys(rng, n, kind, dtype):
    if kind == "same":
        k = np.zeros(n, dtype)
    elif kind == "distinct":
        k = np.arange(n).astype(dtype)
    elif kind == "long":      # runs longer than several tiles
        k = np.repeat(rng.integers(0, 3, n // 20000 + 2), rng.integers(15000, 40000, n // 20000 + 2))[:n].astype(dtype)
    elif kind == "short":
        k = np.repeat(rng.integers(0, 5, n), rng.integers(1, 9, n))[:n].astype(dtype)
    else:                     # BASELINE C4: run lengths uniform[1,1024]
        k = np.repeat(np.arange(n // 256 + 2), rng.integers(1, 1025, n // 256 + 2))[:n].astype(dtype)
    return k


@pytest.mark.parametrize("code", [po.F32, po.S32, po.U32, po.F64, po.S64, po.U64, po.C32, po.U8, po.S8, po.B8, po.S16, po.U16])
def test_pipe_scan(af, orc, code):
    rng = np.random.default_rng(9000 + code)
    shapes = [(4 * PIPE_TILE,), (4 * PIPE_TILE + 1,), (4 * PIPE_TILE + 4,), (5 * PIPE_TILE - 4,), ((1 << 21) + 3,),
              (40000, 3), (29000, 2, 2), (1 << 22,)]
    for shape in shapes:
        base = rand_data(rng, code, shape, nan_frac=0.02, exact=True)
        g = to_gpu(af, base, None, dtype=code)
        for op in range(4):
            if op == po.AF_BINARY_MUL and code in IS_FLOAT:
                continue
            for incl in (True, False):
                t, want = orc.scan(base, 0, op, incl, dtype=code)
                got = af.scan(g, 0, op, incl)
                same_bits(got.to_numpy(), d4(want), "pipe scan op%d incl%d %s" % (op, incl, shape))
        t, want = orc.accum(base, 0, dtype=code)      # b8: counts the non-zeros instead of adding the bytes
        same_bits(af.accum(g, 0).to_numpy(), d4(want), "pipe accum %s" % (shape,))
    # an aligned column view of a larger parent (strided columns, 16-byte aligned offsets)
    base = rand_data(rng, code, (60000, 4), exact=True)
    sl = (slice(8, 8 + 4 * PIPE_TILE + 8), slice(1, 4))
    g = to_gpu(af, base, sl, dtype=code)
    for incl in (True, False):
        t, want = orc.scan(base[sl], 0, 0, incl, dtype=code)
        same_bits(af.scan(g, 0, 0, incl).to_numpy(), d4(want), "pipe scan view incl%d" % incl)


@pytest.mark.parametrize("code", [po.F32, po.S32, po.U32, po.F64, po.S64, po.C32])
def test_pipe_scan_by_key(af, orc, code):
    rng = np.random.default_rng(9100 + code)
    shapes = [(4 * PIPE_TILE_KEYED,), (4 * PIPE_TILE_KEYED + 3,), (6 * PIPE_TILE_KEYED - 4,), ((1 << 21) + 1,), (24000, 3),
              (1 << 22,)]
    if code not in (po.F32, po.S32):
        shapes = shapes[:3] + shapes[4:5] + [((1 << 20) + 5,)]
    for shape in shapes:
        n = int(np.prod(shape))
        base = rand_data(rng, code, shape, nan_frac=0.02, exact=True)
        g = to_gpu(af, base, None, dtype=code)
        for kind in ("same", "distinct", "long", "short", "c4"):
            for kcode in (po.S32, po.U32, po.S64):
                keys = np.asfortranarray(_keys(rng, n, kind, po.NP_OF[kcode]).reshape(shape, order="F"))
                gk = to_gpu(af, keys, None, dtype=kcode)
                for op in ((0, 2, 3) if kcode == po.S32 else (0,)):
                    if op != 0 and code == po.C32 and kind not in ("c4", "long"):
                        continue
                    for incl in (True, False):
                        t, want = orc.scan_by_key(keys, base, 0, op, incl, dtype=code, key_dtype=kcode)
                        got = af.scan_by_key(gk, g, 0, op, incl)
                        same_bits(got.to_numpy(), d4(want), "pipe sbk %s op%d incl%d %s" % (kind, op, incl, shape))


@pytest.mark.parametrize("code", [po.U8, po.B8, po.F32, po.S16, po.F64, po.C64])
def test_pipe_where(af, orc, code):
    rng = np.random.default_rng(9200 + code)
    esz = np.dtype(po.NP_OF[code]).itemsize
    tile = 4 * 16 * 4 * 32 * (16 // esz)
    sub = tile // 4
    for n in (2 * tile, 2 * tile + 1, 3 * tile - 5, 2 * tile + sub + 3, 2 * tile + 2 * sub, (1 << 22) + 17):
        for density in (0.5, 0.01, 0.0, 1.0, 0.9):
            base = rand_data(rng, code, (n,), small=True)
            mask = rng.random(n) < density
            one = np.ones((), base.dtype)
            base = np.where(mask, np.where(base == 0, one, base), 0 * one).astype(base.dtype)
            g = to_gpu(af, base, None, dtype=code)
            want = orc.where(base, dtype=code)
            got = af.where(g)
            assert got.dims == (want.size, 1, 1, 1), (n, density, got.dims, want.size)
            same_bits(got.to_numpy().reshape(-1), want, "pipe where n=%d p=%g" % (n, density))


@pytest.mark.parametrize("code", [po.U8, po.S16, po.F32, po.F64, po.B8])
def test_pipe_where_unaligned_views(af, orc, code):
    """views whose first element is not 16-byte aligned go through the pipelined kernel too: it reads from the aligned
    address below and drops the elements in front (hits right at the start, in the dropped part and at the very end)"""
    rng = np.random.default_rng(9250 + code)
    esz = np.dtype(po.NP_OF[code]).itemsize
    tile = 4 * 16 * 4 * 32 * (16 // esz)
    per_vec = 16 // esz
    n = 2 * tile + 77
    for off in sorted({1, per_vec - 1, max(1, per_vec // 2)}):
        for density in (0.5, 0.02, 1.0):
            base = rand_data(rng, code, (n + off,), small=True)
            mask = rng.random(n + off) < density
            mask[:off + 2] = True              # hits in front of the view (must not be reported) and on its first elements
            mask[-1] = True
            one = np.ones((), base.dtype)
            base = np.where(mask, np.where(base == 0, one, base), 0 * one).astype(base.dtype)
            g = to_gpu(af, base, None, dtype=code).index(af.seq(off, n + off - 1, 1))
            want = orc.where(base[off:], dtype=code)
            got = af.where(g)
            assert got.dims == (want.size, 1, 1, 1), (off, density, got.dims, want.size)
            same_bits(got.to_numpy().reshape(-1), want, "pipe where view off=%d p=%g" % (off, density))


def test_pipe_back_to_back(af, orc):
    """many launches in a row share the descriptor array (generation tags) and the ticket counter"""
    rng = np.random.default_rng(9300)
    n = (1 << 20) + 8
    x = rng.integers(-4, 5, n).astype(np.int32)
    k = _keys(rng, n, "c4", np.int32)
    m = (rng.random(n) < 0.3).astype(np.uint8) * 7
    gx, gk, gm = af.Array.from_numpy(x), af.Array.from_numpy(k), af.Array.from_numpy(m)
    w_scan = orc.scan(x, 0, 0, True)[1]
    w_sbk = orc.scan_by_key(k, x, 0, 0, False)[1]
    w_where = orc.where(m)
    for _ in range(25):
        a, b, c = af.scan(gx, 0, 0, True), af.scan_by_key(gk, gx, 0, 0, False), af.where(gm)
        same_bits(a.to_numpy(), d4(w_scan), "b2b scan")
        same_bits(b.to_numpy(), d4(w_sbk), "b2b sbk")
        same_bits(c.to_numpy().reshape(-1), w_where, "b2b where")


# ---------------------------------------------------------------------------------------------------------
# BASELINE.json's configurations at FULL size (C1..C5), checked against numpy / the oracle
# ---------------------------------------------------------------------------------------------------------
def test_full_size_configs(af, orc):
    rng = np.random.default_rng(0xAF00)
    # C1: inclusive scan of 2^26 integer-valued f32 (every prefix exactly representable -> bit-exact in any order)
    n1 = 1 << 26
    x1 = rng.integers(-4, 5, n1).astype(np.float32)
    want = np.cumsum(x1.astype(np.float64))
    assert np.abs(want).max() < 2 ** 24
    g1 = af.Array.from_numpy(x1)
    got = af.accum(g1, 0).to_numpy().reshape(-1)
    assert np.array_equal(got, want.astype(np.float32))
    assert af.sum_all(g1) == float(want[-1])
    del got, want, g1
    # C4: exclusive add scan_by_key of 2^28 s32, run lengths uniform[1,1024] -> the oracle's serial loop
    n4 = 1 << 28
    v4 = rng.integers(-15, 15, n4).astype(np.int32)
    k4 = np.repeat(np.arange(n4 // 256 + 2, dtype=np.int32), rng.integers(1, 1025, n4 // 256 + 2))[:n4]
    got = af.scan_by_key(af.Array.from_numpy(k4), af.Array.from_numpy(v4), 0, af.AF_BINARY_ADD, False).to_numpy().reshape(-1)
    want = orc.scan_by_key(k4, v4, 0, 0, False)[1].reshape(-1)
    assert got.tobytes() == want.tobytes()
    del got, want, v4, k4
    # C3: sum_all / imax_all (first occurrence among exact ties) of 2^30 f32; C5: count / where of a 2^30 u8 mask
    n3 = 1 << 30
    u3 = (rng.integers(0, 1 << 24, n3, dtype=np.int32).astype(np.float32) / np.float32(1 << 24))
    g3 = af.Array.from_numpy(u3)
    v, i = af.imax_all(g3)
    assert (v, i) == (float(u3.max()), int(np.argmax(u3)))
    v, i = af.imin_all(g3)
    last_min = n3 - 1 - int(np.argmin(u3[::-1]))
    assert (v, i) == (float(u3.min()), last_min)
    del g3, u3
    for p in (0.5, 0.01):
        m = np.where(rng.random(n3, dtype=np.float32) < p, rng.integers(1, 256, n3, dtype=np.uint8), 0).astype(np.uint8)
        gm = af.Array.from_numpy(m)
        want = np.flatnonzero(m).astype(np.uint32)
        assert af.count_all(gm) == want.size
        got = af.where(gm)
        assert got.dims == (want.size, 1, 1, 1)
        assert got.to_numpy().reshape(-1).tobytes() == want.tobytes()
        del got, want, gm, m
