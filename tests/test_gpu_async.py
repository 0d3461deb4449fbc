"""`-m gpu`: the issue / collect forms (afx_*_async + afx_pending_*, include/afb200.h) against the oracle -- many ops are
enqueued back to back, their scalars / indices / hit counts are collected afterwards from the pinned host records."""
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


def test_issue_collect_matches_oracle(af, port_oracle):
    orc = port_oracle
    rng = np.random.default_rng(77)
    n = 3 * (1 << 20) + 17
    x = rng.integers(-4, 5, n).astype(np.float32)
    iv = rng.integers(-1000, 1001, n).astype(np.int32)
    u = (rng.integers(0, 50, n) / 64.0).astype(np.float32)
    m = (rng.random(n) < 0.3).astype(np.uint8) * rng.integers(1, 256, n).astype(np.uint8)
    c = (rng.integers(-4, 5, n) + 1j * rng.integers(-4, 5, n)).astype(np.complex64)
    gx, gi, gu, gm, gc = (af.Array.from_numpy(a) for a in (x, iv, u, m, c))
    pend = []
    for rep in range(4):                       # 4 x 10 = 40 ops in flight (the ring holds 63)
        pend.append(("sum", af.reduce_all_async(gx, af.OP_ADD)))
        pend.append(("sum_i", af.reduce_all_async(gi, af.OP_ADD)))
        pend.append(("min", af.reduce_all_async(gi, af.OP_MIN)))
        pend.append(("max", af.reduce_all_async(gu, af.OP_MAX)))
        pend.append(("count", af.reduce_all_async(gm, af.OP_COUNT)))
        pend.append(("any", af.reduce_all_async(gm, af.OP_ANY)))
        pend.append(("csum", af.reduce_all_async(gc, af.OP_ADD)))
        pend.append(("imax", af.ireduce_all_async(gu, af.OP_MAX)))
        pend.append(("imin", af.ireduce_all_async(gu, af.OP_MIN, None, 1000)))
        pend.append(("where", af.where_async(gm, None, 5)))
    want_where = orc.where(m).astype(np.uint32) + 5
    for name, p in pend:
        r = p.result()
        if name == "sum":
            assert r == float(orc.sum_all(x)[1][0])
        elif name == "sum_i":
            assert r == float(orc.sum_all(iv)[1][0])
        elif name == "min":
            assert r == float(iv.min())
        elif name == "max":
            assert r == float(u.max())
        elif name == "count":
            assert r == float(np.count_nonzero(m))
        elif name == "any":
            assert r == 1.0
        elif name == "csum":
            assert r == complex(orc.sum_all(c, scalar_api=False)[1][0])
        elif name == "imax":
            wv, wi = orc.imax_all(u)
            assert r == (float(wv), int(wi))
        elif name == "imin":
            wv, wi = orc.imin_all(u)
            assert r == (float(wv), int(wi) + 1000)
        else:
            arr, off, tot = r
            assert (off, tot) == (0, want_where.size)
            assert arr.to_numpy().reshape(-1).tobytes() == want_where.tobytes()


def test_pending_lifecycle(af):
    x = af.Array.from_numpy(np.arange(1 << 16, dtype=np.float32))
    p = af.reduce_all_async(x, af.OP_ADD)
    assert p.result() == float(np.arange(1 << 16, dtype=np.float64).sum())
    with pytest.raises(af.AfError):
        p.result()                            # a pending op is collected once
    q = af.where_async(x)
    del q                                     # dropped without collecting: released, the output goes back to the pool
    w = af.where_async(x)
    with pytest.raises(af.AfError):           # wrong collector for a where
        import ctypes
        re, im, idx = ctypes.c_double(), ctypes.c_double(), ctypes.c_uint()
        h, w.handle = w.handle, None
        af._check(af.lib.afx_pending_scalar(h, ctypes.byref(re), ctypes.byref(im), ctypes.byref(idx)))
    assert af.sum_all(x) == float(np.arange(1 << 16, dtype=np.float64).sum())


def test_uploads_and_downloads_from_two_threads(af):
    """af_create_array / af_get_data_ptr run on dedicated copy streams: a reader thread downloads results while the main
    thread uploads the next inputs (what bench.py's end-to-end arm does); every byte must still be right"""
    from concurrent.futures import ThreadPoolExecutor
    rng = np.random.default_rng(5)
    n = 1 << 22
    inputs = [rng.integers(-100, 100, n).astype(np.int32) for _ in range(6)]
    with ThreadPoolExecutor(max_workers=1) as reader:
        futs = []
        for x in inputs * 3:
            g = af.Array.from_numpy(x)                       # upload (blocking for this thread only)
            s = af.accum(g, 0)                               # kernel on the compute stream
            futs.append((x, reader.submit(s.to_numpy)))      # download on the reader thread
            del g, s                                         # buffers go back to the pool while copies are in flight
        for x, f in futs:
            assert np.array_equal(f.result().reshape(-1), np.cumsum(x, dtype=np.int64).astype(np.int32))
