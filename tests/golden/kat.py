"""Known-answer vectors transcribed from the reference's own tests (/root/reference/test/*.cpp).

Each case: (name, citation, call, inputs, expected).  `call` names an AfOracle / arrayfire_b200
method; arrays are column-major (Fortran order), exactly as the reference test constructs them.
These pin BOTH the oracle (tests/test_oracle_golden.py, CPU) and the CUDA path (tests/test_gpu_golden.py).
"""
import numpy as np

NAN = float("nan")
F = lambda data, shape: np.asfortranarray(np.array(data, dtype=np.float32).reshape(shape, order="F"))  # noqa: E731


def C(re, im, shape):
    return np.asfortranarray((np.array(re, np.float32) + 1j * np.array(im, np.float32)).astype(np.complex64).reshape(shape, order="F"))


_SUBA = F([0.0168, 0.0278, 0.0317, 0.0248, 0.0131, 0.0197, 0.0321, 0.0362, 0.0279, 0.0141, 0.0218, 0.0353, 0.0394,
           0.0297, 0.0143, 0.0224, 0.0363, 0.0104, 0.0302, 0.0142, 0.0217, 0.0409, 0.0398, 0.0302, 0.0144], (5, 5))
SUBA_VIEW = _SUBA[1:4, 2:5]  # a(seq(1,3), seq(2,4)), test/ireduce.cpp:424-431

_re = [0.005, NAN, -6.3, NAN, -0.5, NAN, NAN, 0.2, -1205.4, 8.9]
_im = [NAN, NAN, -9.0, -0.005, -0.3, 0.007, NAN, 0.1, NAN, 4.5]

# (name, citation, op, input, dim, expected values, expected indices)
IREDUCE_DIM = [
    ("MinNaN", "test/ireduce.cpp:196-221", "imin", F([1, NAN, 5, 0.1, NAN, -0.5, NAN, 0], (4, 2)), 0,
     np.array([0.1, -0.5], np.float32), [3, 1]),
    ("MaxNaN", "test/ireduce.cpp:223-248", "imax", F([1, NAN, 5, 0.1, NAN, -0.5, NAN, 0], (4, 2)), 0,
     np.array([5.0, 0.0], np.float32), [2, 3]),
    ("MinCplxNaN", "test/ireduce.cpp:250-284", "imin", C(_re, _im, (5, 2)), 0,
     np.array([-0.5 - 0.3j, 0.2 + 0.1j], np.complex64), [4, 2]),
    ("MaxCplxNaN", "test/ireduce.cpp:286-320", "imax", C(_re, _im, (5, 2)), 0,
     np.array([-6.3 - 9.0j, 8.9 + 4.5j], np.complex64), [2, 4]),
    ("MinPreferLargerIdxIfEqual", "test/ireduce.cpp:322-342", "imin", F([0, 50, 50, 0], (4,)), 0,
     np.array([0.0], np.float32), [3]),
    ("MaxPreferSmallerIdxIfEqual", "test/ireduce.cpp:344-364", "imax", F([0, 50, 50, 0], (4,)), 0,
     np.array([50.0], np.float32), [1]),
    ("MinCplxPreferLargerIdxIfEqual", "test/ireduce.cpp:366-393", "imin", C([0, 50, 50, 0], [0, 50, 50, 0], (4,)), 0,
     np.array([0j], np.complex64), [3]),
    ("MaxCplxPreferSmallerIdxIfEqual", "test/ireduce.cpp:395-422", "imax", C([0, 50, 50, 0], [0, 50, 50, 0], (4,)), 0,
     np.array([50 + 50j], np.complex64), [1]),
    ("max_subarray_0", "test/ireduce.cpp:461-479", "imax", SUBA_VIEW, 0, np.array([0.0394, 0.0363, 0.0409], np.float32), [1, 0, 0]),
    ("min_subarray_0", "test/ireduce.cpp:481-499", "imin", SUBA_VIEW, 0, np.array([0.0297, 0.0104, 0.0302], np.float32), [2, 1, 2]),
    ("max_subarray_1", "test/ireduce.cpp:501-519", "imax", SUBA_VIEW, 1, np.array([0.0409, 0.0398, 0.0302], np.float32), [2, 2, 1]),
    ("min_subarray_1", "test/ireduce.cpp:521-539", "imin", SUBA_VIEW, 1, np.array([0.0353, 0.0104, 0.0297], np.float32), [0, 1, 0]),
]

# (name, citation, op, input, expected value, expected index)
IREDUCE_ALL = [
    ("max_subarray_all", "test/ireduce.cpp:433-445", "imax_all", SUBA_VIEW, np.float32(0.0409), 6),
    ("min_subarray_all", "test/ireduce.cpp:447-459", "imin_all", SUBA_VIEW, np.float32(0.0104), 4),
    ("reduce_jit_array", "test/ireduce.cpp:542-551", "imin_all", F([2, 3], (2,)), np.float32(2.0), 0),
]

_iota23 = F([0, 1, 2, 3, 4, 5], (2, 3))
# (name, citation, op, input, dim, expected (column-major flat), expected dims)
DIMWISE = [
    ("sum_dim0", "test/reduce.cpp:2414-2450", "sum", _iota23, 0, [1, 5, 9], (1, 3, 1, 1)),
    ("sum_dim1", "test/reduce.cpp:2414-2450", "sum", _iota23, 1, [6, 9], (2, 1, 1, 1)),
    ("sum_dim2", "test/reduce.cpp:2414-2450", "sum", _iota23, 2, [0, 1, 2, 3, 4, 5], (2, 3, 1, 1)),
    ("sum_dim3", "test/reduce.cpp:2414-2450", "sum", _iota23, 3, [0, 1, 2, 3, 4, 5], (2, 3, 1, 1)),
    ("accum_1d", "test/scan.cpp:212-232", "accum", F([0, 1, 2, 3, 4], (5,)), 0, [0, 1, 3, 6, 10], (5, 1, 1, 1)),
    ("accum_2d_dim0", "test/scan.cpp:234-259", "accum", F(range(9), (3, 3)), 0, [0, 1, 3, 3, 7, 12, 6, 13, 21], (3, 3, 1, 1)),
    ("accum_2d_dim1", "test/scan.cpp:234-259", "accum", F(range(9), (3, 3)), 1, [0, 1, 2, 3, 5, 7, 9, 12, 15], (3, 3, 1, 1)),
]


def exclusive_ones_cases():
    """test/scan.cpp:261-348 ExclusiveSum1D / 2D_Dim0..3: exclusive add-scan of 80000 ones along every dim."""
    n = 80000
    ramp = np.arange(n, dtype=np.int32)
    yield "ExclusiveSum1D", np.ones((n,), np.int32, order="F"), 0, ramp.reshape((n, 1, 1, 1), order="F")
    g = np.stack([ramp, ramp], axis=1)  # [n, 2]
    yield "ExclusiveSum2D_Dim0", np.ones((n, 2), np.int32, order="F"), 0, np.asfortranarray(g).reshape((n, 2, 1, 1), order="F")
    yield "ExclusiveSum2D_Dim1", np.ones((2, n), np.int32, order="F"), 1, np.asfortranarray(g.T).reshape((2, n, 1, 1), order="F")
    yield "ExclusiveSum2D_Dim2", np.ones((1, 2, n), np.int32, order="F"), 2, np.asfortranarray(g.T.reshape(1, 2, n)).reshape((1, 2, n, 1), order="F")
    yield "ExclusiveSum2D_Dim3", np.ones((1, 1, 2, n), np.int32, order="F"), 3, np.asfortranarray(g.T.reshape(1, 1, 2, n))


class CRand:
    """glibc rand()/srand() (TYPE_3, r[i] = r[i-3] + r[i-31]) so the reference's `std::srand(0)` key/data
    generators of test/scan_by_key.cpp:36-77 are reproduced bit-for-bit on any host."""

    def __init__(self, seed):
        seed = seed or 1
        r = [0] * 34
        r[0] = seed
        for i in range(1, 31):
            hi, lo = divmod(r[i - 1], 127773)
            w = 16807 * lo - 2836 * hi
            r[i] = w + 2147483647 if w < 0 else w
        for i in range(31, 34):
            r[i] = r[i - 31]
        self.r = [x & 0xFFFFFFFF for x in r]
        for _ in range(34, 344):
            self._step()

    def _step(self):
        v = (self.r[-3] + self.r[-31]) & 0xFFFFFFFF
        self.r.append(v)
        self.r.pop(0)
        return v

    def rand(self):
        return self._step() >> 1


def scan_by_key_case(dims, scan_dim, node_lengths, key_start, key_end, data_start, data_end, np_dtype):
    """createScanKey + createScanData of test/scan_by_key.cpp:40-77 (note: ONE rand stream: srand(0) in
    createScanKey, then createScanData continues it)."""
    rng = CRand(0)
    d = list(dims) + [1] * (4 - len(dims))
    n = int(np.prod(d))
    key = np.zeros(n, np.int32)
    stride = int(np.prod(d[:scan_dim]))
    for start in range(stride):
        keyval = 0
        i = 0
        for index in range(start, n, stride):
            is_node = any(i % nl == 0 for nl in node_lengths)
            if is_node and (rng.rand() % 2):
                keyval = key_start + rng.rand() % (key_end - key_start)
            key[index] = keyval
            i = (i + 1) % d[scan_dim]
    data = np.array([data_start + rng.rand() % (data_end - data_start) for _ in range(n)], dtype=np_dtype)
    return key.reshape(d, order="F"), data.reshape(d, order="F")


def scan_by_key_verify(dims, data, key, out, scan_dim, op, inclusive, eps):
    """verify<> of test/scan_by_key.cpp:79-109 (the reference test's own checker), vectorised per line."""
    d = list(dims) + [1] * (4 - len(dims))
    n = int(np.prod(d))
    stride = int(np.prod(d[:scan_dim]))
    fin, fkey, fout = data.reshape(-1, order="F"), key.reshape(-1, order="F"), out.reshape(-1, order="F")
    init = {0: 0, 2: np.inf if fout.dtype.kind == "f" else np.iinfo(fout.dtype).max,
            3: -np.inf if fout.dtype.kind == "f" else np.iinfo(fout.dtype).min}[op]
    f = {0: lambda a, b: a + b, 2: min, 3: max}[op]
    bad = 0
    for start in range(0, stride, max(1, stride // 64)):  # sample lines: python-loop cost
        keyval = fkey[start]
        gold = init
        i = 0 if inclusive else 1
        index = start + (0 if inclusive else stride)
        while index < n:
            if fkey[index] != keyval or i == 0:
                keyval = fkey[index]
                if inclusive:
                    gold = fin[index]
                    bad += abs(float(gold) - float(fout[index])) > eps
                else:
                    gold = init
            else:
                gold = f(gold, fin[index - (0 if inclusive else stride)])
                bad += abs(float(gold) - float(fout[index])) > eps
            index += stride
            i = (i + 1) % d[scan_dim]
    return bad
