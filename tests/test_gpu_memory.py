"""`-m gpu`: the memory manager's garbage-collection thresholds (reference: common/DefaultMemoryManager.cpp:88-89,117-126,
175-187): AF_MAX_BUFFERS caps the number of buffers (in use + cached); reaching it frees the cached ones before the next
allocation.  Runs in a child process because the limit is read once per process."""
import os
import subprocess
import sys

import pytest

from tests.gpu_util import gpu_available

pytestmark = pytest.mark.gpu

CHILD = r'''
import ctypes, sys
import numpy as np
import arrayfire_b200 as af

def mem_info():
    v = [ctypes.c_size_t() for _ in range(4)]
    af._check(af.lib.af_device_mem_info(*[ctypes.byref(x) for x in v]))
    return [x.value for x in v]          # alloc_bytes, alloc_buffers, lock_bytes, lock_buffers

keep = af.Array.from_numpy(np.zeros(1000, np.float32))
base = mem_info()
arrs = [af.Array.from_numpy(np.zeros(1000 * (i + 2), np.float32)) for i in range(6)]   # six distinct sizes
del arrs
cached = mem_info()
assert cached[1] - base[1] == 6, ("freed buffers stay cached", base, cached)
a = af.Array.from_numpy(np.zeros(77777, np.float32))      # 1 + 6 + ... >= AF_MAX_BUFFERS = 6: GC runs before this allocation
after = mem_info()
assert after[1] == base[1] + 1, ("cached buffers were released by the threshold GC", base, cached, after)
assert af.sum_all(a) == 0.0
print("OK")
'''


def test_max_buffers_threshold_triggers_gc():
    if not gpu_available():
        pytest.skip("no CUDA device")
    env = dict(os.environ, AF_MAX_BUFFERS="6")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", CHILD], capture_output=True, text=True, env=env, cwd=root, timeout=300)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout + r.stderr


def test_array_moves_to_the_active_device():
    """an array used on another device is moved there, like the reference's checkAndMigrate (cuda/Array.cpp:61-74);
    needs two GPUs (skipped on a one-GPU box; run with `gpurun --gpus 2`)"""
    if not gpu_available():
        pytest.skip("no CUDA device")
    import numpy as np
    import arrayfire_b200 as af
    if af.device_count() < 2:
        pytest.skip("needs two GPUs")
    rng = np.random.default_rng(5)
    x = rng.integers(-9, 10, 300001).astype(np.float32)
    m = (rng.random(70000) < 0.3).astype(np.uint8)
    try:
        af.set_device(0)
        gx, gm = af.Array.from_numpy(x), af.Array.from_numpy(m)
        view = gx.index(af.seq(5, 200004, 1))
        af.set_device(1)
        assert af.sum_all(gx) == float(x.sum(dtype=np.float64))            # moved on first use
        assert gx.device == 1
        assert af.sum_all(view) == float(x[5:200005].sum(dtype=np.float64))  # a view moves with its whole buffer
        w = af.where(gm)
        assert w.device == 1
        assert np.array_equal(w.to_numpy().reshape(-1), np.flatnonzero(m).astype(np.uint32))
        acc = af.accum(gx, 0).to_numpy().reshape(-1)
        assert np.array_equal(acc, np.cumsum(x.astype(np.float64)).astype(np.float32))
        af.set_device(0)
        assert af.max_all(gx) == float(x.max())                            # ... and back
        assert gx.device == 0
    finally:
        af.set_device(0)
