"""`-m gpu`: the remaining BASELINE.md section 3 configurations at FULL size (the rest live in
tests/test_gpu_parity.py::test_full_size_configs): C2 8192 x 8192 along both dims, C3 sum_all f32 / s32 at 2^30 including the
full-range wrap-around case, C3 imax with NaNs sprinkled in, C4 with geometric(mean 8) run lengths and one run longer
than 2^27, C1 accuracy set (uniform[0,1), 2^26) within 1e-5 of an f64 truth with the reference CPU's own error printed
beside it.  Size-independent checks (numpy truths) where the serial oracle would take minutes."""
import numpy as np
import pytest

from oracle import pyoracle as po
from tests.gpu_util import gpu_available

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def af():
    if not gpu_available():
        pytest.skip("no CUDA device")
    import arrayfire_b200
    return arrayfire_b200


def test_c2_matrix_8192(af, port_oracle):
    rng = np.random.default_rng(0xAF02)
    m = np.asfortranarray(rng.integers(-4, 5, (8192, 8192)).astype(np.float32))   # parity set: exact partial sums
    g = af.Array.from_numpy(m)
    for dim in (0, 1):
        got = af.sum(g, dim).to_numpy().reshape(-1)
        want = m.sum(axis=dim, dtype=np.float64).astype(np.float32)
        assert got.tobytes() == want.tobytes(), "sum dim%d" % dim
        # ... and bit-identical to the serial oracle on a slab the CPU finishes quickly
    t, want = port_oracle.sum(m[:, :256], 0)
    assert af.sum(g.index(af.span, af.seq(0, 255, 1)), 0).to_numpy().reshape(-1).tobytes() == want.reshape(-1).tobytes()
    u = np.asfortranarray(rng.random((8192, 8192), dtype=np.float32))
    gu = af.Array.from_numpy(u)
    for dim in (0, 1):
        got = af.max(gu, dim).to_numpy().reshape(-1)
        assert got.tobytes() == u.max(axis=dim).tobytes(), "max dim%d" % dim
        got = af.min(gu, dim).to_numpy().reshape(-1)
        assert got.tobytes() == u.min(axis=dim).tobytes(), "min dim%d" % dim
    # index-returning forms on the same matrix: first maximum / last minimum per line
    v, i = af.imax(gu, 0)
    assert i.to_numpy().reshape(-1).tolist() == np.argmax(u, axis=0).tolist()
    v, i = af.imin(gu, 1)
    last_min = u.shape[1] - 1 - np.argmin(u[:, ::-1], axis=1)
    assert i.to_numpy().reshape(-1).tolist() == last_min.tolist()


def test_c3_sum_all_2p30(af):
    rng = np.random.default_rng(0xAF03)
    n = 1 << 30
    x = rng.integers(-4, 5, n, dtype=np.int8).astype(np.float32)       # parity set
    truth = int(x.sum(dtype=np.float64))
    assert abs(truth) < 2 ** 24
    assert af.sum_all(af.Array.from_numpy(x)) == float(truth)
    del x
    i3 = rng.integers(-1000, 1001, n, dtype=np.int32)
    assert af.sum_all(af.Array.from_numpy(i3)) == float(int(i3.sum(dtype=np.int64)))
    assert af.min_all(af.Array.from_numpy(i3)) == -1000.0 and af.max_all(af.Array.from_numpy(i3)) == 1000.0
    del i3
    w = rng.integers(-2 ** 31, 2 ** 31, n, dtype=np.int64).astype(np.int32)   # full range: the sum wraps modulo 2^32
    g = af.Array.from_numpy(w)
    wrapped = np.int64(w.sum(dtype=np.int64)).astype(np.int32)
    assert af.sum_all(g) == float(wrapped)
    arr = af.sum_all_array(g)
    assert arr.dtype == po.S32 and int(arr.to_numpy().reshape(-1)[0]) == int(wrapped)
    del w, g


def test_c3_imax_with_nans_2p30(af):
    rng = np.random.default_rng(0xAF13)
    n = 1 << 30
    u = (rng.integers(0, 1 << 24, n, dtype=np.int32).astype(np.float32) / np.float32(1 << 24))
    nan_at = rng.integers(0, n, n // 1000)            # p = 1e-3
    u[nan_at] = np.nan
    u[0] = np.nan                                     # a NaN seed must not win either
    g = af.Array.from_numpy(u)
    v, i = af.imax_all(g)
    assert (v, i) == (float(np.nanmax(u)), int(np.nanargmax(u)))
    v, i = af.imin_all(g)
    mn = np.nanmin(u)
    last = int(np.flatnonzero(u == mn)[-1])
    assert (v, i) == (float(mn), last)
    assert af.max_all(g) == float(np.nanmax(u)) and af.min_all(g) == float(mn)


def test_c4_geometric_runs_with_a_giant_run(af, port_oracle):
    rng = np.random.default_rng(0xAF04)
    n = 1 << 28
    vals = rng.integers(-15, 15, n, dtype=np.int32)
    lens = rng.geometric(1.0 / 8.0, n // 6)
    keys = np.repeat(np.arange(lens.size, dtype=np.int32), lens)[:n]
    assert keys.size == n
    lo = (1 << 26) + 12345
    keys[lo:lo + (1 << 27) + 4097] = -7                # one run > 2^27 (spans several shards in the sharded path)
    gk, gv = af.Array.from_numpy(keys), af.Array.from_numpy(vals)
    for incl in (False, True):
        got = af.scan_by_key(gk, gv, 0, af.AF_BINARY_ADD, incl).to_numpy().reshape(-1)
        want = port_oracle.scan_by_key(keys, vals, 0, 0, incl)[1].reshape(-1)
        assert got.tobytes() == want.tobytes(), "inclusive=%s" % incl
        del got, want
    # the by-key REDUCTION of the same input (section 8f-1) against numpy
    ko, vo = af.sum_by_key(gk, gv, 0)
    heads = np.flatnonzero(np.concatenate([[True], keys[1:] != keys[:-1]]))
    assert ko.to_numpy().reshape(-1).tobytes() == keys[heads].tobytes()
    want = np.add.reduceat(vals.astype(np.int64), heads).astype(np.int32)
    assert vo.to_numpy().reshape(-1).tobytes() == want.tobytes()


def test_c1_accuracy_set_2p26(af, port_oracle, capsys):
    rng = np.random.default_rng(0xAF01)
    n = 1 << 26
    x = rng.random(n, dtype=np.float32)
    truth = float(x.sum(dtype=np.float64))
    g = af.Array.from_numpy(x)
    got = af.sum_all(g)
    ref_cpu = float(port_oracle.sum_all(x)[1][0])      # the reference's serial f32 loop (stagnates at 2^24)
    with capsys.disabled():
        print("\n[C1 accuracy 2^26] truth %.6e  afb200 rel.err %.2e  reference CPU rel.err %.2e" % (
            truth, abs(got - truth) / truth, abs(ref_cpu - truth) / truth))
    assert abs(got - truth) <= 1e-5 * truth
    sc = af.accum(g, 0).to_numpy().reshape(-1)
    want = np.cumsum(x, dtype=np.float64)
    rel = np.abs(sc.astype(np.float64) - want) / want
    assert rel.max() <= 1e-5, rel.max()
    # f64 at 1e-12
    xd = rng.random(1 << 24)
    gd = af.Array.from_numpy(xd)
    t = float(np.sum(xd.astype(np.longdouble)))
    assert abs(af.sum_all(gd) - t) <= 1e-12 * t
    scd = af.accum(gd, 0).to_numpy().reshape(-1)
    wd = np.cumsum(xd.astype(np.longdouble))
    assert float(np.max(np.abs(scd - wd) / wd)) <= 1e-12
