"""`-m gpu`: SURVEY.md section 8(f)-4 -- unevaluated element-wise expressions fused into whole-array reductions
(api_expr.cpp + kernels/expr.cu).  The flagship case is the reference's pi benchmark (examples/benchmarks/pi.cpp:28-31):
sum<float>(sqrt(x*x + y*y) < 1).  Every operator is rounded on its own (no contraction across interpreter steps), so the
results are bit-identical to numpy evaluating the same expression in the same precision."""
import numpy as np
import pytest

from tests.gpu_util import gpu_available

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def af():
    if not gpu_available():
        pytest.skip("no CUDA device")
    import arrayfire_b200
    return arrayfire_b200


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_pi_benchmark_expression_is_one_fused_launch(af, dt):
    rng = np.random.default_rng(31415)
    for n in (1, 1000, 3 * (1 << 20) + 17):
        x, y = rng.random(n).astype(dt), rng.random(n).astype(dt)
        gx, gy = af.Array.from_numpy(x), af.Array.from_numpy(y)
        af.launch_count(reset=True)
        inside = af.lt(af.sqrt(af.add(af.mul(gx, gx), af.mul(gy, gy))), 1.0)
        assert af.launch_count() == 0                     # nothing has been evaluated yet
        got = af.sum_all(inside)                          # sum of a b8 expression counts, as in the reference
        assert af.launch_count() == 1, af.launch_count()  # ONE kernel: the operands are streamed once, no temporary
        want = int(np.count_nonzero(np.sqrt(x * x + y * y) < dt(1)))
        assert got == want
        assert af.count_all(inside) == want and af.any_true_all(inside) == (want > 0) and af.all_true_all(inside) == (want == n)
        assert abs(4.0 * got / n - np.pi) < (0.2 if n >= 1000 else 4.1)


def test_expression_values_match_numpy_bit_for_bit(af):
    rng = np.random.default_rng(27)
    n = 100003
    x = rng.standard_normal(n).astype(np.float32)
    y = (rng.standard_normal(n) + 2).astype(np.float32)
    d = rng.standard_normal(n)
    gx, gy, gd = af.Array.from_numpy(x), af.Array.from_numpy(y), af.Array.from_numpy(d)
    cases = [
        (af.add(gx, gy), x + y), (af.sub(gx, gy), x - y), (af.mul(gx, gy), x * y), (af.div(gx, gy), x / y),
        (af.sqrt(af.abs(gx)), np.sqrt(np.abs(x))), (af.minof(gx, gy), np.minimum(x, y)), (af.maxof(gx, gy), np.maximum(x, y)),
        (af.add(af.mul(gx, 2.5), 1.0), x * np.float32(2.5) + np.float32(1.0)),
        (af.mul(gd, gx), d * x.astype(np.float64)),                                 # f64 (x) f32 -> f64
    ]
    for expr, want in cases:
        got = expr.to_numpy().reshape(-1)                                            # materialises the expression
        assert got.dtype == want.dtype and got.tobytes() == want.tobytes()
    bools = [(af.lt(gx, gy), x < y), (af.ge(gx, 0.0), x >= 0), (af.eq(gx, gx), x == x), (af.neq(gx, gy), x != y),
             (af.logical_and(af.gt(gx, 0.0), af.le(gy, 2.0)), (x > 0) & (y <= 2)),
             (af.logical_or(af.lt(gx, -1.0), af.logical_not(af.lt(gy, 2.0))), (x < -1) | ~(y < 2))]
    for expr, want in bools:
        assert expr.dtype == af.b8
        assert af.count_all(expr) == int(np.count_nonzero(want))                     # fused
        assert np.array_equal(expr.to_numpy().reshape(-1) != 0, want)                # evaluated
    # fused sum / min / max of a value expression: the sum within f32 tolerance of float64 truth, min / max exact
    e = af.add(af.mul(gx, gx), gy)
    v = x * x + y
    assert abs(af.sum_all(e) - float(v.astype(np.float64).sum())) <= 1e-5 * float(np.abs(v).sum())
    assert af.min_all(e) == float(v.min()) and af.max_all(e) == float(v.max())
    # a consumer outside the fused set evaluates the expression first and then runs as usual
    w = af.where(af.gt(gx, 1.0)).to_numpy().reshape(-1)
    assert np.array_equal(w, np.flatnonzero(x > 1).astype(np.uint32))
    assert np.array_equal(af.accum(af.mul(gx, 2.0), 0).to_numpy().reshape(-1)[:5], np.cumsum((x * np.float32(2))[:5], dtype=np.float32))


def test_lazy_constant_and_limits(af):
    c = af.constant(3.0, (1000, 7), af.f32)
    af.launch_count(reset=True)
    assert af.sum_all(c) == 21000.0 and af.launch_count() == 1
    assert c.to_numpy().shape[:2] == (1000, 7) and float(c.to_numpy().reshape(-1)[123]) == 3.0
    x = af.Array.from_numpy(np.ones(5000, np.float32))
    e = x
    for _ in range(100):                       # far beyond one interpreter program: sub-trees are evaluated on the way
        e = af.add(e, 1.0)
    assert af.sum_all(e) == 5000.0 * 101
    with pytest.raises(af.AfError) as ei:      # integer operands are outside the supported set
        af.add(af.Array.from_numpy(np.arange(4, dtype=np.int32)), af.Array.from_numpy(np.arange(4, dtype=np.int32)))
    assert ei.value.code == 301
    with pytest.raises(af.AfError) as ei:      # no broadcasting
        af.add(af.Array.from_numpy(np.ones(4, np.float32)), af.Array.from_numpy(np.ones(5, np.float32)))
    assert ei.value.code == 301


def test_integer_operands_in_comparisons(af):
    """8- to 32-bit integer operands in comparisons / logic (the reference compares in implicit(l, r), api/c/implicit.cpp:20-45,
    binary.cpp): exact in the double interpreter, fused into count / sum / any / all, and feeding af_where"""
    rng = np.random.default_rng(99)
    n = 3 * (1 << 18) + 11
    for dt in (np.int32, np.uint32, np.int16, np.uint16, np.int8, np.uint8):
        info = np.iinfo(dt)
        x = rng.integers(max(info.min, -40), min(info.max, 40), n, endpoint=True).astype(dt)
        y = rng.integers(max(info.min, -40), min(info.max, 40), n, endpoint=True).astype(dt)
        x[:8] = [info.min, info.max, 0, 1, info.min, info.max, 3, 3]
        y[:8] = [info.max, info.min, 0, 1, info.min, info.max, 3, 4]
        gx, gy = af.Array.from_numpy(x), af.Array.from_numpy(y)
        af.launch_count(reset=True)
        e = af.eq(gx, 3)
        assert af.launch_count() == 0
        assert af.count_all(e) == int(np.count_nonzero(x == 3)) and af.launch_count() == 1
        for fn, ref in ((af.lt, np.less), (af.gt, np.greater), (af.le, np.less_equal), (af.ge, np.greater_equal), (af.eq, np.equal),
                        (af.neq, np.not_equal)):
            got = fn(gx, gy)
            assert got.dtype == af.b8
            assert af.sum_all(got) == int(np.count_nonzero(ref(x, y))), (dt, fn)
            assert got.to_numpy().reshape(-1).astype(bool).tobytes() == ref(x, y).tobytes(), (dt, fn)
        assert af.count_all(af.logical_and(gx, gy)) == int(np.count_nonzero((x != 0) & (y != 0)))
        assert af.count_all(af.logical_or(gx, gy)) == int(np.count_nonzero((x != 0) | (y != 0)))
        assert af.count_all(af.logical_not(gx)) == int(np.count_nonzero(x == 0))
        assert af.any_true_all(af.gt(gx, 39)) == bool((x > 39).any()) and bool(af.all_true_all(af.ge(gx, int(info.min))))
        hits = af.where(af.logical_and(af.ge(gx, 2), af.lt(gx, 5)))
        assert hits.to_numpy().reshape(-1).tobytes() == np.flatnonzero((x >= 2) & (x < 5)).astype(np.uint32).tobytes()
    # s32 / u32 against f32: the integer is rounded to float first (16777217 == 16777216.0f); against f64 it is exact
    xi = np.asarray([16777217, 16777216, -16777217, 5, 2147483647, -7], np.int32)
    yf = np.asarray([16777216, 16777216, -16777216, 5.5, 2147483648.0, -7], np.float32)
    gi, gf, gd = af.Array.from_numpy(xi), af.Array.from_numpy(yf), af.Array.from_numpy(yf.astype(np.float64))
    assert af.eq(gi, gf).to_numpy().reshape(-1).astype(bool).tolist() == (xi.astype(np.float32) == yf).tolist()
    assert af.lt(gf, gi).to_numpy().reshape(-1).astype(bool).tolist() == (yf < xi.astype(np.float32)).tolist()
    assert af.eq(gi, gd).to_numpy().reshape(-1).astype(bool).tolist() == (xi.astype(np.float64) == yf.astype(np.float64)).tolist()
    assert af.gt(gi, 5.5).to_numpy().reshape(-1).astype(bool).tolist() == (xi.astype(np.float32) > np.float32(5.5)).tolist()
    # narrower signed against wider signed; b8 against an integer
    s16 = np.asarray([-3, 7, 100, -32768], np.int16)
    s32 = np.asarray([-3, 8, 99, -32769], np.int32)
    assert af.le(af.Array.from_numpy(s16), af.Array.from_numpy(s32)).to_numpy().reshape(-1).astype(bool).tolist() == (s16 <= s32).tolist()
    flags = af.gt(af.Array.from_numpy(s16), 0)
    assert af.eq(flags, af.Array.from_numpy(np.asarray([0, 1, 1, 0], np.int32))).to_numpy().reshape(-1).astype(bool).all()
    # lazy integer constants evaluate to what the stored type holds
    c = af.constant(5, (1000, 3), af.s32)
    assert c.dtype == af.s32 and af.sum_all(c) == 15000.0 and int(c.to_numpy().reshape(-1)[77]) == 5
    assert int(af.constant(200, (10,), af.u8).to_numpy().reshape(-1)[0]) == 200
    # outside the supported set: integer arithmetic, mixed signedness (the reference would wrap the negatives), 64-bit integers
    for a, b, fn in ((xi, xi, af.mul), (np.asarray([1], np.int32), np.asarray([1], np.uint32), af.lt),
                     (np.asarray([1], np.int64), np.asarray([1], np.int64), af.eq)):
        with pytest.raises(af.AfError) as ei:
            fn(af.Array.from_numpy(a), af.Array.from_numpy(b))
        assert ei.value.code == 301
