"""CPU tests of the statistics oracle.

 * the restated running mean (oracle/oracle.cpp) is pinned BIT for BIT to the reference's own kernel template
   (src/backend/cpu/kernel/mean.hpp compiled in place into oracle/_ref/libafref.so);
 * mean / var / stdev / norm / corrcoef / cov restatements are checked against float64 truth at the tolerances the
   reference's own tests use (test/mean.cpp, var.cpp, stdev.cpp, norm.cpp: ASSERT_NEAR), and on the inline known-answer
   cases those tests hold (constant arrays, Var.ISSUE2117, MeanAll.SubArray, StandardDev.InvalidType, Norm identity).
"""
import numpy as np
import pytest

from oracle import pyoracle as po

ALL = [po.F32, po.F64, po.C32, po.C64, po.S32, po.U32, po.S64, po.U64, po.S16, po.U16, po.S8, po.U8, po.B8, po.F16]
SHAPES = [(1,), (7,), (1000,), (5, 3), (1, 9), (33, 1, 4), (4, 5, 3, 2), (257, 3)]


def rand(rng, code, shape):
    n = int(np.prod(shape))
    dt = po.NP_OF[code]
    if code in (po.F32, po.F64, po.F16):
        a = (rng.standard_normal(n) * 4 + 3).astype(dt)
    elif code in (po.C32, po.C64):
        a = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(dt)
    elif code == po.B8:
        a = rng.integers(0, 2, n).astype(dt)
    else:
        info = np.iinfo(dt)
        a = rng.integers(max(info.min, -100), min(info.max, 100), n, endpoint=True).astype(dt)
    return np.asfortranarray(a.reshape(shape, order="F"))


@pytest.mark.parametrize("code", ALL)
def test_mean_restatement_vs_reference_kernel(port_oracle, ref_oracle, code):
    rng = np.random.default_rng(500 + code)
    for shape in SHAPES:
        x = rand(rng, code, shape)
        for view in (x, x[tuple(slice(s // 3, s) for s in x.shape)]):
            if view.size == 0:
                continue
            for dim in range(4):
                a, b = port_oracle.o.mean(view, dim, code), ref_oracle.o.mean(view, dim, code)
                assert a.tobytes() == b.tobytes(), (shape, dim)
            a, b = port_oracle.o.mean_all(view, code), ref_oracle.o.mean_all(view, code)
            assert a == b or (a != a and b != b), shape


def tol_of(code):
    return 1e-12 if code in (po.F64, po.C64, po.S64, po.U64) else (2e-3 if code == po.F16 else 1e-5)


@pytest.mark.parametrize("code", ALL)
def test_moments_vs_float64_truth(port_oracle, code):
    orc = port_oracle
    rng = np.random.default_rng(900 + code)
    cplx = code in (po.C32, po.C64)
    for shape in [(1000,), (64, 9), (7, 5, 3)]:
        x = rand(rng, code, shape)
        x64 = x.astype(np.complex128 if cplx else np.float64)
        scale = np.abs(x64).max() + 1.0
        for dim in range(len(shape)):
            _, m = orc.mean(x, dim, code)
            want = x64.mean(axis=dim, keepdims=True)
            assert np.abs(m.reshape(want.shape).astype(want.dtype) - want).max() <= tol_of(code) * scale * 4
            n = shape[dim]
            if n > 1:
                for bias, ddof in ((po.AF_VARIANCE_POPULATION, 0), (po.AF_VARIANCE_SAMPLE, 1), (po.AF_VARIANCE_DEFAULT, 1)):
                    _, v = orc.var(x, dim, bias, code)
                    d = x64 - want
                    wv = (d * d).sum(axis=dim, keepdims=True) / (n - ddof)      # complex: plain product, as var.cpp does
                    assert np.abs(v.reshape(wv.shape).astype(wv.dtype) - wv).max() <= tol_of(code) * scale * scale * 8
                if not cplx and code != po.F16:
                    _, sd = orc.stdev(x, dim, po.AF_VARIANCE_POPULATION, code)
                    assert np.abs(sd.reshape(want.shape) - x64.std(axis=dim, keepdims=True)).max() <= tol_of(code) * scale * 8
        assert abs(orc.mean_all(x, code) - x64.mean()) <= tol_of(code) * scale * 4
        d = x64 - x64.mean()
        assert abs(orc.var_all(x, po.AF_VARIANCE_SAMPLE, code) - (d * d).sum() / (x.size - 1)) <= tol_of(code) * scale * scale * 8


def test_reference_known_answers(port_oracle):
    orc = port_oracle
    # test/mean.cpp:221-241 MeanAll: a constant array averages to the constant
    for dt, val, dims in ((np.float64, 2.1, (10, 10)), (np.float32, 2.1, (10, 5, 2)), (np.int32, 2, (5, 5, 2, 2)),
                          (np.uint32, 2, (100,)), (np.int16, 2, (5, 5, 2, 2)), (np.uint8, 2, (100,))):
        x = np.full(dims, val, dtype=dt, order="F")
        assert abs(orc.mean_all(x) - float(dt(val))) <= 1e-6
    x = np.full((10, 5, 2), 2.1 + 0j, dtype=np.complex64, order="F")
    assert abs(orc.mean_all(x) - complex(np.complex64(2.1))) <= 1e-6
    # test/var.cpp:175-190 Var.ISSUE2117: the variance of a constant is exactly zero
    _, v = orc.var(np.ones((100, 300), dtype=np.float32, order="F"), 1, po.AF_VARIANCE_SAMPLE)
    assert float(np.abs(v).sum()) == 0.0
    # test/mean.cpp:354-368 MeanAll.SubArray
    rng = np.random.default_rng(3)
    big = np.asfortranarray(rng.random((10, 10, 10, 10)).astype(np.float32))
    sub = big[0:1, :, :, :]
    assert abs(orc.mean_all(sub) - float(sub.astype(np.float64).mean())) <= np.finfo(np.float32).eps * sub.size
    # test/stdev.cpp:148-150 StandardDev.InvalidType
    with pytest.raises(po.AfTypeError):
        orc.stdev(np.full(10, 1 - 1j, dtype=np.complex128), 0)
    # test/norm.cpp:157-168, 263-273: identity matrices have matrix norms 1
    eye = np.asfortranarray(np.eye(17, dtype=np.float32))
    assert orc.norm(eye, po.AF_NORM_MATRIX_1) == 1.0 and orc.norm(eye, po.AF_NORM_MATRIX_INF) == 1.0


def test_norm_corrcoef_cov_vs_numpy(port_oracle):
    orc = port_oracle
    rng = np.random.default_rng(11)
    for dt in (np.float32, np.float64, np.complex64, np.complex128):
        a = rng.standard_normal((40, 30)).astype(dt)
        if np.iscomplexobj(a):
            a = (a + 1j * rng.standard_normal((40, 30))).astype(dt)
        a = np.asfortranarray(a)
        tol = 1e-12 if dt in (np.float64, np.complex128) else 2e-5
        A = np.abs(a.astype(np.complex128))
        assert abs(orc.norm(a) - np.sqrt((A * A).sum())) <= tol * 100
        assert abs(orc.norm(a, po.AF_NORM_VECTOR_1) - A.sum()) <= tol * 2000
        assert abs(orc.norm(a, po.AF_NORM_VECTOR_INF) - A.max()) <= tol * 10
        assert abs(orc.norm(a, po.AF_NORM_VECTOR_P, 3.5) - (A ** 3.5).sum() ** (1 / 3.5)) <= tol * 100
        assert abs(orc.norm(a, po.AF_NORM_MATRIX_1) - A.sum(0).max()) <= tol * 100
        assert abs(orc.norm(a, po.AF_NORM_MATRIX_INF) - A.sum(1).max()) <= tol * 100
        assert abs(orc.norm(a, po.AF_NORM_MATRIX_L_PQ, 2.0, 3.0) - ((A ** 2).sum(0) ** 0.5 ** 1).__pow__(3.0).sum() ** (1 / 3.0)) <= tol * 100
    x = np.asfortranarray(rng.standard_normal((500, 4)).astype(np.float32))
    y = np.asfortranarray((0.3 * x + rng.standard_normal((500, 4))).astype(np.float32))
    assert abs(orc.corrcoef(x, y) - np.corrcoef(x.ravel(order="F"), y.ravel(order="F"))[0, 1]) <= 1e-4
    _, c = orc.cov(x, y, po.AF_VARIANCE_SAMPLE)
    x64, y64 = x.astype(np.float64), y.astype(np.float64)
    want = ((x64 - x64.mean()) * (y64 - y64.mean())).sum(0) / (500 - 1)
    assert np.abs(c.reshape(-1) - want).max() <= 1e-4


def test_meanvar_large_known_answers(port_oracle):
    """MeanVar*.Large (test/meanvar.cpp:239-266, 299-311): 0..1999 as 2000x1x1 / 1x2000 / 1x1x2000 along its long dim ->
    mean 999.5, sample variance 333500; as 500x4 along dim 0 -> means 249.5 + 500 k, sample variance 20875 each."""
    orc = port_oracle
    ramp = np.arange(2000)
    for code in (po.F32, po.F64, po.S32, po.U32, po.S64, po.U64, po.S16, po.U16):
        x = ramp.astype(po.NP_OF[code])
        for dim, shape in enumerate([(2000,), (1, 2000), (1, 1, 2000)]):
            _, (m, v) = orc.var(np.asfortranarray(x.reshape(shape)), dim, po.AF_VARIANCE_SAMPLE, want_mean=True)
            assert abs(float(m.reshape(-1)[0]) - 999.5) <= 1e-3 and abs(float(v.reshape(-1)[0]) - 333500) <= 333500 * 1e-5, (code, dim)
        _, (m, v) = orc.var(np.asfortranarray(x.reshape((500, 4), order="F")), 0, po.AF_VARIANCE_SAMPLE, want_mean=True)
        assert np.abs(m.reshape(-1) - np.asarray([249.5, 749.5, 1249.5, 1749.5])).max() <= 1e-3
        assert np.abs(v.reshape(-1) - 20875).max() <= 20875 * 1e-5


def test_meanvar_weighted_restatement(port_oracle):
    """the weighted af_meanvar / af_var_weighted restatement (api/c/var.cpp:104-142): unit weights reproduce the unweighted pair
    (population: 1 / N; sample: 1 / (N - 1)); in general the mean is weighted, the squared deviations are not"""
    orc = port_oracle
    rng = np.random.default_rng(77)
    for dt in (np.float32, np.float64, np.int32, np.int64, np.uint8):
        for shape in [(2000,), (500, 4), (20, 7, 3), (4, 5, 3, 2)]:
            x = (rng.standard_normal(shape) * 3 + 1).astype(dt) if np.issubdtype(dt, np.floating) else rng.integers(0, 50, shape).astype(dt)
            x = np.asfortranarray(x)
            f64 = dt in (np.float64, np.int64)
            wdt = np.float64 if f64 else np.float32
            tol = 1e-12 if f64 else 2e-5
            x64 = x.astype(np.float64)
            scale = float(np.abs(x64).max()) + 1.0
            for dim in range(len(shape)):
                for bias in (po.AF_VARIANCE_POPULATION, po.AF_VARIANCE_SAMPLE):
                    if bias == po.AF_VARIANCE_SAMPLE and shape[dim] == 1:
                        continue
                    m1, v1 = orc.meanvar_weighted(x, np.ones(shape, wdt), dim, bias)
                    _, (m0, v0) = orc.var(x, dim, bias, want_mean=True)
                    assert m1.dtype == wdt and v1.dtype == wdt
                    assert np.abs(m1.reshape(-1) - m0.reshape(-1)).max() <= tol * scale * 4
                    assert np.abs(v1.reshape(-1) - v0.reshape(-1)).max() <= tol * scale * scale * 16
                    w = np.asfortranarray((rng.random(shape) + 0.5).astype(wdt))
                    w64 = w.astype(np.float64)
                    m, v = orc.meanvar_weighted(x, w, dim, bias)
                    wm = (x64 * w64).sum(axis=dim, keepdims=True) / w64.sum(axis=dim, keepdims=True)
                    want = ((x64 - wm) ** 2).sum(axis=dim) / (w64.sum(axis=dim) - (1 if bias == po.AF_VARIANCE_SAMPLE else 0))
                    assert np.abs(m.reshape(-1) - wm.reshape(-1)).max() <= tol * scale * 4
                    assert np.abs(v.reshape(-1) - want.reshape(-1)).max() <= tol * scale * scale * 16 * max(1.0, float(np.abs(want).max()) / (scale * scale))
    with pytest.raises(po.AfArgError):
        orc.meanvar_weighted(np.ones(4, np.float32), np.ones(4, np.float32), 4)
    with pytest.raises(po.AfArgError):
        orc.meanvar_weighted(np.ones(4, np.float32), np.ones(4, np.int32), 0)


def test_var_and_set_known_answers(port_oracle):
    from tests.golden import kat_stats as ks
    orc = port_oracle
    for code in (po.F32, po.F64, po.S32, po.U32, po.S64, po.U64, po.S16, po.U16, po.S8, po.U8, po.C32, po.C64):
        dt = po.NP_OF[code]
        for dims in ks.VAR_CONST_DIMS[:2]:
            x = np.full(dims, ks.VAR_CONST_VALUE, dtype=dt, order="F")
            for bias in (po.AF_VARIANCE_POPULATION, po.AF_VARIANCE_SAMPLE):
                assert abs(orc.var_all(x, bias)) <= ks.VAR_TOL
        r = ks.VAR_RAMP.astype(dt)
        assert abs(orc.var_all(r, po.AF_VARIANCE_POPULATION) - ks.VAR_RAMP_POPULATION) <= ks.VAR_TOL
        assert abs(orc.var_all(r, po.AF_VARIANCE_SAMPLE) - ks.VAR_RAMP_SAMPLE) <= ks.VAR_TOL
    for x, want in ks.SET_UNIQUE_SORTED:
        assert orc.set_unique_sorted(x).tolist() == want.tolist()
