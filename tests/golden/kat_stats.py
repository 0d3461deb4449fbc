"""Inline known answers of the reference's statistics / set tests (no data file needed) -- they pin the oracle
(tests/test_oracle_stats.py) and the CUDA path (tests/test_gpu_stats.py, tests/test_gpu_next.py)."""
import numpy as np

# Var.AllCPPSmall / Medium / Large (test/var.cpp:52-116): a constant array has variance 0 under both biases; 0..4 has population
# variance 2 and sample variance 2.5; tolerance 1e-3
VAR_CONST_DIMS = [(10, 10), (100, 100), (1000, 1000)]
VAR_CONST_VALUE = 2
VAR_RAMP = np.arange(5)
VAR_RAMP_POPULATION, VAR_RAMP_SAMPLE = 2.0, 2.5
VAR_TOL = 1.0e-3

# Set.SNIPPET_setUniqueSorted / setUniqueSortedDesc (test/set.cpp:161-199): (input, unique); descending input stays descending
SET_UNIQUE_SORTED = [(np.asarray([1, 2, 2, 3, 3, 3], np.int32), np.asarray([1, 2, 3], np.int32)),
                     (np.asarray([3, 3, 3, 2, 2, 1], np.int32), np.asarray([3, 2, 1], np.int32))]
