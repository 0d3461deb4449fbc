"""More known answers / properties of the reference's own reduction tests (/root/reference/test/reduce.cpp) that
tests/golden/kat.py does not hold yet, as data + small generators.  Used by tests/test_oracle_golden_reduce.py (CPU, pins the
oracle) and tests/test_gpu_golden.py (`-m gpu`, the CUDA path)."""
import numpy as np

NAN = np.nan

# MinMax.MinCplxNaN / MaxCplxNaN (test/reduce.cpp:1171-1234): af::min / af::max along dim 0 of a 5 x 2 complex array with NaNs in
# either part; an element with a NaN part is ignored even when its other part is huge (-1205.4)
_RE = [0.005, NAN, -6.3, NAN, -0.5, NAN, NAN, 0.2, -1205.4, 8.9]
_IM = [NAN, NAN, -9.0, -0.005, -0.3, 0.007, NAN, 0.1, NAN, 4.5]
CPLX_NAN = (np.asarray(_RE, np.float32) + 1j * np.asarray(_IM, np.float32)).astype(np.complex64).reshape((5, 2), order="F")
CPLX_NAN_MIN = np.asarray([-0.5 - 0.3j, 0.2 + 0.1j], np.complex64)
CPLX_NAN_MAX = np.asarray([-6.3 - 9.0j, 8.9 + 4.5j], np.complex64)


def nan_vector(seed=7, num=10000, frac=0.25):
    """MinMax.MinMaxNaN / Sum.NaN (test/reduce.cpp:1147-1169,1247-1265): uniform [0,1) with the values below `frac` replaced by NaN"""
    a = np.random.default_rng(seed).random(num).astype(np.float32)
    a[a < frac] = NAN
    return a


def count_nan_pair(seed=8, num=10000):
    """Count.NaN (test/reduce.cpp:1236-1245): round(5 * randu); replacing the 2s by NaN does not change the count (NaN != 0)"""
    b = np.round(5 * np.random.default_rng(seed).random(num)).astype(np.float32)
    a = b.copy()
    a[a == 2] = NAN
    return a, b


def anyall_nan_pair(seed=9, num=10000):
    """AnyAll.NaN (test/reduce.cpp:1287-1299): a 0/1 array A and B = A with the zeros replaced by NaN"""
    a = (np.random.default_rng(seed).random(num) > 0.5).astype(np.float32)
    b = a.copy()
    b[b == 0] = NAN
    return a, b


# Reduce.nanval_issue_3255 (test/reduce.cpp:2382-2412): one NaN at every position in turn; *_by_key_nan(nanval = 1) never yields NaN
ISSUE_3255_KEYS = np.asarray([0, 0, 1, 1, 1, 2, 2, 0], np.uint32)
ISSUE_3255_VALS = np.asarray([1, 2, 3, 4, 5, 6, 7, 8], np.float64)


def reduce_half_cases():
    """ReduceHalf.Sum / Product (test/reduce.cpp:1375-1488): constant f16 arrays reduced along `dim` (-1: the first non-singleton
    dim) give value * n (sum) / value ** n (product, skipped when it overflows) as f32.  Yields (value, dims, dim)."""
    for n in (10, 2048, 2049, 8192):
        for tail in ((), (10,), (10, 10), (10, 10, 10)):
            yield 1.0, (n,) + tail, -1
    for n in (10, 2048, 2049, 8192):
        for tail in ((), (10,), (10, 10)):
            yield 1.25, (10, n) + tail, 1
    for n in (10, 2048, 2049, 8192):
        for tail in ((), (10,)):
            yield 1.25, (10, 10, n) + tail, 2


HALF_RAMP = np.asarray([1, 2, 3, 4, 5, 6, 7], np.float16)   # ReduceHalf.Min / Max / Count / AnyTrue / AllTrue (test/reduce.cpp:1490-1548)
