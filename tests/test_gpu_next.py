"""`-m gpu`: SURVEY.md section 8(f)-3 -- af_sat, af_diff1 / af_diff2, af_set_unique (sorted input) and the tiled single-pass
scan along dim > 0 (dim_scan_kernel) that af_sat stands on, against the oracle.  Integer and exactly-representable float
inputs: bit-exact."""
import numpy as np
import pytest

from oracle import pyoracle as po
from tests.gpu_util import gpu_available, rand_data, random_subview, same_bits, to_gpu

pytestmark = pytest.mark.gpu

SAT_TYPES = [po.F32, po.F64, po.S32, po.U32, po.B8, po.S8, po.U8, po.S64, po.U64, po.S16, po.U16]
DIFF_TYPES = [po.F32, po.F64, po.C32, po.C64, po.S32, po.U32, po.S64, po.U64, po.S16, po.U16, po.S8, po.U8, po.B8]


@pytest.fixture(scope="module")
def af():
    if not gpu_available():
        pytest.skip("no CUDA device")
    import arrayfire_b200
    return arrayfire_b200


def d4(a):
    return np.asarray(a).reshape(po.dims4(np.asarray(a).shape), order="F")


@pytest.mark.parametrize("code", [po.F32, po.S32, po.U32, po.U8, po.S16, po.B8, po.F64, po.C32])
def test_scan_along_higher_dims_at_scale(af, port_oracle, code):
    """shapes that give the tiled look-back kernel many splits, ragged last tiles, narrow and unaligned dim 0, batches"""
    orc = port_oracle
    rng = np.random.default_rng(8100 + code)
    shapes = [(128, 1000), (4, 50001), (32, 4097, 3), (130, 515), (1000, 70, 2, 3), (7, 300, 5), (256, 2, 700), (512, 2048)]
    for shape in shapes:
        base = rand_data(rng, code, shape, small=True)
        views = [(base, None)]
        sl = random_subview(rng, base.shape)
        if base[sl].size:
            views.append((base[sl], sl))
        for x, s in views:
            g = to_gpu(af, base, s, dtype=code)
            for dim in range(1, len(shape)):
                for op, incl in ((po.AF_BINARY_ADD, True), (po.AF_BINARY_ADD, False), (po.AF_BINARY_MAX, True), (po.AF_BINARY_MIN, False)):
                    if op != po.AF_BINARY_ADD and code in (po.B8,):
                        continue
                    t, want = orc.scan(x, dim, op, incl, dtype=code)
                    got = af.scan(g, dim, op, incl)
                    assert got.dtype == t
                    same_bits(got.to_numpy(), d4(want), "scan %s dim %d op %d incl %s" % (shape, dim, op, incl))
            t, want = orc.accum(x, 1, dtype=code)
            same_bits(af.accum(g, 1).to_numpy(), d4(want), "accum %s" % (shape,))


@pytest.mark.parametrize("code", SAT_TYPES)
def test_sat(af, port_oracle, code):
    orc = port_oracle
    rng = np.random.default_rng(8200 + code)
    for shape in [(5, 3), (64, 64), (257, 70), (1000, 300), (33, 65, 3), (3000, 2000)]:
        x = rand_data(rng, code, shape, small=True)
        t, want = orc.sat(x, dtype=code)
        got = af.sat(af.Array.from_numpy(x, dtype=code))
        assert got.dtype == t, (shape, got.dtype, t)
        same_bits(got.to_numpy(), d4(want), "sat %s" % (shape,))
    with pytest.raises(af.AfError):      # sat.cpp:38: at least two dimensions
        af.sat(af.Array.from_numpy(np.arange(10, dtype=np.float32)))
    with pytest.raises(af.AfError):
        af.sat(af.Array.from_numpy(np.zeros((4, 4), dtype=np.complex64, order="F")))


@pytest.mark.parametrize("code", DIFF_TYPES)
def test_diff(af, port_oracle, code):
    orc = port_oracle
    rng = np.random.default_rng(8300 + code)
    for shape in [(1,), (2,), (1000,), (5, 3), (33, 1, 4), (4, 5, 3, 2), (257, 70), (100000, 3)]:
        base = rand_data(rng, code, shape, small=True)
        views = [(base, None)]
        if base.size > 1:
            sl = random_subview(rng, base.shape)
            if base[sl].size:
                views.append((base[sl], sl))
        for x, s in views:
            g = to_gpu(af, base, s, dtype=code)
            for dim in range(4):
                for order, fn in ((1, af.diff1), (2, af.diff2)):
                    want = orc.diff(x.view(np.int8) if code == po.B8 and x.dtype == np.bool_ else x, dim, order)
                    got = fn(g, dim)
                    assert got.dtype == code
                    if want.size == 0:
                        assert got.elements == 0, (shape, dim, order)
                    else:
                        same_bits(got.to_numpy(), want, "diff%d %s dim %d" % (order, shape, dim))


def test_set_unique_sorted(af, port_oracle):
    orc = port_oracle
    rng = np.random.default_rng(8400)
    for dt in (np.float32, np.float64, np.int32, np.uint32, np.int64, np.uint64, np.int16, np.uint16, np.int8, np.uint8):
        for n in (1, 2, 1000, 300001):
            x = np.sort(rng.integers(0, max(2, n // 3), n)).astype(dt)
            got = af.set_unique(af.Array.from_numpy(x), True)
            want = orc.set_unique_sorted(x)
            assert got.elements == want.size
            assert got.to_numpy().reshape(-1).tobytes() == want.tobytes()
    with pytest.raises(af.AfError) as ei:
        af.set_unique(af.Array.from_numpy(np.array([3, 1, 2], dtype=np.int32)), False)
    assert ei.value.code == 301


@pytest.mark.parametrize("code", [po.F32, po.F64, po.S32, po.U32, po.U8, po.S16])
def test_reference_sat_integral_image(af, code):
    """SAT.IntegralImage (/root/reference/test/sat.cpp:41-51): sat(a) == accum(accum(a, 0), 1) on a 530 x 671 array, bit for bit"""
    rng = np.random.default_rng(8800 + code)
    x = np.asfortranarray(rng.integers(0, 9, (530, 671)).astype(po.NP_OF[code]))
    g = af.Array.from_numpy(x, dtype=code)
    s = af.sat(g)
    c = af.accum(af.accum(g, 0), 1)
    assert s.dtype == c.dtype and s.dims == c.dims
    assert s.to_numpy().tobytes() == c.to_numpy().tobytes()


def test_reference_diff_large_dim(af):
    """Diff1.DiffLargeDim / Diff2.DiffLargeDim (/root/reference/test/diff1.cpp:186-211, diff2.cpp:179-206): a constant line of
    65535 * 32 + 1 elements along each of the four dims differentiates to zeros"""
    n = 65535 * 32 + 1
    for dim in range(4):
        dims = [1, 1, 1, 1]
        dims[dim] = n
        one = af.Array.from_numpy(np.ones(dims, dtype=np.float32, order="F"))
        d1, d2 = af.diff1(one, dim), af.diff2(one, dim)
        assert d1.dims[dim] == n - 1 and d2.dims[dim] == n - 2
        assert af.sum_all(d1) == 0.0 and af.sum_all(d2) == 0.0
        assert af.count_all(d1) == 0.0 and af.count_all(d2) == 0.0


def test_reference_set_unique_snippets(af):
    """Set.SNIPPET_setUniqueSorted / setUniqueSortedDesc (test/set.cpp:161-199)"""
    from tests.golden import kat_stats as ks
    for x, want in ks.SET_UNIQUE_SORTED:
        got = af.set_unique(af.Array.from_numpy(x), True)
        assert got.elements == want.size and got.to_numpy().reshape(-1).tolist() == want.tolist()


def test_tile(af):
    """af_tile (api/c/tile.cpp:41-90; test/tile.cpp): every dtype size, repeats along each dim, views; numpy's tile on the
    Fortran-ordered array is the truth (pure data movement: bit-exact)"""
    rng = np.random.default_rng(8600)
    for dt in (np.uint8, np.int16, np.float32, np.float64, np.complex128, np.float16):
        for shape, reps in [((5,), (3, 1, 1, 1)), ((5,), (1, 4, 1, 1)), ((7, 3), (2, 3, 1, 1)), ((4, 1, 3), (1, 5, 2, 2)),
                            ((2, 3, 2, 2), (3, 1, 2, 1)), ((1000, 33), (1, 7, 1, 1)), ((1, 1), (100000, 3, 1, 1))]:
            x = np.asfortranarray((rng.random(shape) * 100).astype(dt))
            got = af.tile(af.Array.from_numpy(x), *reps)
            x4 = x.reshape(po.dims4(shape), order="F")
            want = np.tile(x4, reps)
            assert got.dims == tuple(want.shape), (shape, reps)
            assert got.to_numpy().reshape(want.shape, order="F").tobytes() == np.asfortranarray(want).tobytes(), (dt, shape, reps)
    x = np.asfortranarray(rng.random((40, 30)).astype(np.float32))
    sub = af.Array.from_numpy(x).index(af.seq(3, 20, 1), af.seq(5, 25, 2))
    got = af.tile(sub, 2, 3)
    want = np.tile(x[3:21, 5:26:2], (2, 3))
    assert got.to_numpy().reshape(want.shape, order="F").tobytes() == np.asfortranarray(want).tobytes()
    with pytest.raises(af.AfError) as ei:
        af.tile(af.Array.from_numpy(x), 0, 1)
    assert ei.value.code == 203   # AF_ERR_SIZE


def test_weighted_mean_broadcasts_its_weights(af):
    """af_mean_weighted repeats a weights dim of 1 up to the input's (api/c/mean.cpp:112-124); a mismatch that is not 1 is an
    argument error (mean.cpp:118)"""
    rng = np.random.default_rng(8601)
    x = np.asfortranarray(rng.standard_normal((300, 17, 3)).astype(np.float32))
    x64 = x.astype(np.float64)
    for wshape, dim in (((300, 1, 1), 0), ((300, 1, 3), 0), ((1, 17, 1), 1), ((1, 1, 3), 2), ((300, 17, 1), 1)):
        w = np.asfortranarray((rng.random(wshape) + 0.5).astype(np.float32))
        wb = np.broadcast_to(w.astype(np.float64), x.shape)
        got = af.mean_weighted(af.Array.from_numpy(x), af.Array.from_numpy(w), dim).to_numpy()
        want = (x64 * wb).sum(axis=dim) / wb.sum(axis=dim)
        assert np.abs(got.reshape(-1, order="F") - want.reshape(-1, order="F")).max() <= 1e-4, (wshape, dim)
    with pytest.raises(af.AfError) as ei:
        af.mean_weighted(af.Array.from_numpy(x), af.Array.from_numpy(np.ones((150, 17, 3), np.float32, order="F")), 0)
    assert ei.value.code == 202
