"""Shared helpers of the `-m gpu` parity tests: build device arrays / sub-array views through the C-ABI and
compare results with the oracle."""
import numpy as np

from oracle import pyoracle as po


def gpu_available():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def rand_data(rng, code, shape, nan_frac=0.0, small=False, exact=False):
    """random column-major array of af dtype `code`; exact=True keeps float sums exactly representable."""
    n = int(np.prod(shape))
    dt = po.NP_OF[code]
    if code in (po.F32, po.F64, po.F16):
        if exact or small:
            a = rng.integers(-4, 5, n).astype(dt)
        else:
            a = (rng.standard_normal(n) * 4).astype(dt)
    elif code in (po.C32, po.C64):
        if exact or small:
            a = (rng.integers(-4, 5, n) + 1j * rng.integers(-4, 5, n)).astype(dt)
        else:
            a = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(dt)
    elif code == po.B8:
        a = (rng.integers(0, 2, n) * rng.choice([1, 1, 2, -3], n)).astype(dt)
    else:
        info = np.iinfo(dt)
        lo, hi = (max(info.min, -6), min(info.max, 6)) if small else (info.min, info.max)
        a = rng.integers(lo, hi, n, dtype=np.int64 if dt != np.uint64 else np.uint64, endpoint=True).astype(dt)
    if nan_frac > 0 and code in (po.F32, po.F64, po.F16, po.C32, po.C64):
        a[rng.random(n) < nan_frac] = np.nan
    return np.asfortranarray(a.reshape(shape, order="F"))


def random_subview(rng, shape):
    """slices describing a sub-array with an offset (what af_index with seqs produces)"""
    sl = []
    for s in shape:
        lo = int(rng.integers(0, max(1, s // 3)))
        hi = int(rng.integers(max(lo + 1, s - s // 3), s + 1))
        sl.append(slice(lo, hi))
    return tuple(sl)


def to_gpu(af, base, slices=None, dtype=None):
    """device copy of `base`; with `slices`, an af_index view of it sharing the parent buffer"""
    a = af.Array.from_numpy(base, dtype=dtype)
    if slices is None:
        return a
    seqs = [af.seq(s.start, s.stop - 1, 1) for s in slices]
    return a.index(*seqs)


def same_bits(x, y, what=""):
    x, y = np.asarray(x), np.asarray(y)
    assert x.dtype == y.dtype, "%s dtype %s vs %s" % (what, x.dtype, y.dtype)
    assert tuple(po.dims4(x.shape)) == tuple(po.dims4(y.shape)), "%s dims %s vs %s" % (what, x.shape, y.shape)
    xb, yb = x.reshape(-1, order="F"), y.reshape(-1, order="F")
    if xb.tobytes() != yb.tobytes():
        # -0.0 vs +0.0 of a float min/max reduction is the one documented bit-level freedom
        neq = np.flatnonzero(~((xb == yb) | ((xb != xb) & (yb != yb))))
        assert neq.size == 0, "%s: %d mismatches, first at %d: got %r want %r" % (
            what, neq.size, neq[0], xb[neq[0]], yb[neq[0]])


def close(x, y, rtol, what=""):
    x, y = np.asarray(x), np.asarray(y)
    assert x.dtype == y.dtype and tuple(po.dims4(x.shape)) == tuple(po.dims4(y.shape)), what
    xf, yf = x.reshape(-1, order="F"), y.reshape(-1, order="F")
    np.testing.assert_allclose(xf, yf, rtol=rtol, atol=0, equal_nan=True, err_msg=what)
