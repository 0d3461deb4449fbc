"""Known-answer vectors of the reference's own by-key / ragged tests (/root/reference/test/reduce.cpp), transcribed as data:
they pin the oracle (tests/test_oracle_golden_bykey.py, CPU) and the CUDA path (tests/test_gpu_golden.py, `-m gpu`).
Every entry: (name, op, keys, vals, dim, nanval or None, expected keys, expected vals) -- vals / expected are given as small
integers and cast to each tested dtype (the reference runs them TYPED over its type list)."""
import numpy as np

KEYS_A = [0, 2, 2, 9, 5, 5, 5, 8]          # test/reduce.cpp:735,754,771,865: runs 0 | 2 2 | 9 | 5 5 5 | 8
VALS_A = [0, 7, 1, 6, 2, 5, 3, 4]
RKEYS_A = [0, 2, 9, 5, 8]

# (name, function, keys, vals, expected keys, expected vals, reference lines)
TYPED = [
    ("productReduceByKey", "product_by_key", KEYS_A, VALS_A, RKEYS_A, [0, 7, 6, 30, 4], "test/reduce.cpp:732-749"),
    ("minReduceByKey", "min_by_key", KEYS_A, VALS_A, RKEYS_A, [0, 1, 6, 2, 4], "test/reduce.cpp:751-766"),
    ("maxReduceByKey", "max_by_key", KEYS_A, VALS_A, RKEYS_A, [0, 7, 6, 5, 4], "test/reduce.cpp:768-783"),
    ("allTrueReduceByKey", "all_true_by_key", KEYS_A, [0, 1, 1, 1, 0, 1, 1, 1], RKEYS_A, [0, 1, 1, 0, 1], "test/reduce.cpp:785-800"),
    ("anyTrueReduceByKey", "any_true_by_key", [0, 2, 2, 9, 5, 5, 8, 8], [0, 1, 1, 1, 0, 1, 0, 0], [0, 2, 9, 5, 8], [0, 1, 1, 1, 0],
     "test/reduce.cpp:802-818"),
    ("countReduceByKey", "count_by_key", [0, 2, 2, 9, 5, 5, 5, 5], [0, 1, 1, 1, 0, 1, 1, 1], [0, 2, 9, 5], [0, 2, 1, 3],
     "test/reduce.cpp:820-835"),
    ("sumReduceByKey (nDim0, one column)", "sum_by_key", KEYS_A, VALS_A, RKEYS_A, [0, 8, 6, 10, 4], "test/reduce.cpp:862-889"),
]

# ReduceByKeyNans (test/reduce.cpp:837-860): product with nanval = 1 over {0, 7, NaN, 6, 2, 5, 3, 4} -> {0, 7, 6, 30, 4}
NAN_CASE = ("ReduceByKeyNans", "product_by_key", KEYS_A, [0, 7, np.nan, 6, 2, 5, 3, 4], 1.0, RKEYS_A, [0, 7, 6, 30, 4])

# RaggedMax (test/reduce.cpp:1951-1987): (array as column-major list, dims, ragged lengths, their dims, dim, values, indices)
RAGGED = [
    ("RaggedMax.simple", [1, 2, 3, 4, 5, 6], (3, 2), [9, 2], (1, 2), 0, [3, 5], [2, 1]),
    ("RaggedMax.simpleDim1", [1, 2, 3, 4, 5, 6, 7, 8], (2, 4), [8, 2], (2, 1), 1, [7, 4], [3, 1]),
]


def issue_2955():
    """ReduceByKey.ISSUE_2955 (test/reduce.cpp:2134-2144): 256 keys 0..126 then 129 x 1 -> 128 runs"""
    key = np.arange(256, dtype=np.int32)
    key[127:] = 1
    return key, 128


def issue_3062():
    """ReduceByKey.ISSUE_3062 (test/reduce.cpp:2158-2180): 129 ones under one key -> 129 (sum and count), along dim 0 and dim 1"""
    return np.zeros(129, np.uint32), np.ones(129, np.uint32), 129
