"""`-m gpu` tests of the shard-local primitives behind the multi-GPU layer (afx_* of include/afb200.h), on ONE
GPU: a global array is cut into shards, every shard goes through LocalGpu exactly as a rank would run it, the
payloads are "all-gathered" by stacking them, and the combine rules of arrayfire_b200/sharded.py produce the
global answer, checked bit-for-bit against the oracle on the whole array.  (The same ShardedOps code over a
real process group is covered by tests/test_sharded_gloo.py on CPU and by bench.py --gpus N on the box.)"""
import numpy as np
import pytest

from oracle import pyoracle as po
from tests.gpu_util import gpu_available

pytestmark = pytest.mark.gpu


class StackGroup:
    """a process group of `world` ranks simulated in one process: rank r's payload is recorded, and the
    gather returns the stack once every rank has contributed (two-pass protocol driven by the test)."""

    def __init__(self, world):
        self.world, self.rank = world, 0
        self.payloads = {}
        self.replay = False

    def all_gather_bytes(self, payload):
        import torch
        key = self.key
        self.key += 1      # an op may gather more than once: every collective of a call gets its own slot
        if not self.replay:
            self.payloads.setdefault(key, {})[self.rank] = payload.clone()
            # first pass: the other ranks are unknown yet -> answer with copies of our own payload
            return torch.stack([payload] * self.world)
        rows = self.payloads[key]
        return torch.stack([rows[r] for r in range(self.world)])


@pytest.fixture(scope="module")
def env():
    if not gpu_available():
        pytest.skip("no CUDA device")
    import arrayfire_b200 as af
    from arrayfire_b200 import sharded as sh
    return af, sh, po.AfOracle("port")


def run_sharded(af, sh, world, call):
    """call(ops, rank) -> result; executed twice per rank (record pass, then replay pass with all payloads)"""
    grp = StackGroup(world)
    ops = sh.ShardedOps(None, sh.LocalGpu())
    ops.group, ops.world = grp, world
    out = [None] * world
    for replay in (False, True):
        grp.replay = replay
        for r in range(world):
            grp.rank = ops.rank = r
            grp.key = 0
            ops.__dict__.pop("_bases", None)   # the simulated ranks share one ShardedOps: no cross-rank caching
            out[r] = call(ops, r)
    return out


@pytest.mark.parametrize("world", [2, 5])
def test_sharded_scan_by_key_single_gpu(env, world):
    af, sh, orc = env
    rng = np.random.default_rng(77 + world)
    n = 3 * (1 << 20) + 12345
    vals = rng.integers(-15, 15, n).astype(np.int32)
    cases = {
        "runs_1024": np.repeat(np.arange(n // 200 + 2), rng.integers(1, 1025, n // 200 + 2))[:n],
        "short": np.repeat(np.arange(n // 4 + 2), rng.integers(1, 9, n // 4 + 2))[:n],
        "one_run": np.zeros(n),
        "giant_middle": np.where((np.arange(n) > n // 7) & (np.arange(n) < n - n // 9), 7, np.arange(n) % 3),
    }
    for cname, keys in cases.items():
        keys = keys.astype(np.int32)
        for binop, incl in ((0, False), (0, True), (3, False), (2, True)):
            def call(ops, r):
                lo, hi = sh.shard_bounds(n, world, r)
                return ops.scan_by_key(af.Array.from_numpy(keys[lo:hi]), af.Array.from_numpy(vals[lo:hi]), binop, incl).to_numpy().reshape(-1)
            got = np.concatenate(run_sharded(af, sh, world, call))
            want = orc.scan_by_key(keys, vals, 0, binop, incl)[1].reshape(-1)
            assert got.tobytes() == want.tobytes(), (cname, binop, incl, np.flatnonzero(got != want)[:5])


@pytest.mark.parametrize("world", [3])
def test_sharded_reduce_scan_where_single_gpu(env, world):
    af, sh, orc = env
    rng = np.random.default_rng(99)
    n = (1 << 21) + 77
    x = rng.integers(-4, 5, n).astype(np.float32)
    ties = (rng.integers(0, 50, n) / 64.0).astype(np.float32)
    u8 = (rng.random(n) < 0.3).astype(np.uint8) * rng.integers(1, 256, n).astype(np.uint8)

    def shard(a, r):
        lo, hi = sh.shard_bounds(n, world, r)
        return af.Array.from_numpy(a[lo:hi]), lo

    got = run_sharded(af, sh, world, lambda ops, r: ops.reduce_all(sh.OP_SUM, shard(x, r)[0]))
    assert all(np.float32(g).tobytes() == orc.sum_all(x, scalar_api=False)[1][:1].tobytes() for g in got)
    got = run_sharded(af, sh, world, lambda ops, r: ops.reduce_all(sh.OP_COUNT, shard(u8, r)[0]))
    assert all(int(g) == int(orc.count_all(u8)[1][0]) for g in got)
    for op, fn in ((sh.OP_MAX, orc.imax_all), (sh.OP_MIN, orc.imin_all)):
        got = run_sharded(af, sh, world, lambda ops, r: ops.ireduce_all(op, *shard(ties, r)))
        wv, wi = fn(ties)
        assert all((float(v), int(i)) == (float(wv), int(wi)) for v, i in got), (got, wv, wi)
    for incl in (True, False):
        got = run_sharded(af, sh, world, lambda ops, r: ops.scan(shard(x, r)[0], 0, incl).to_numpy().reshape(-1))
        want = orc.scan(x, 0, 0, incl)[1].reshape(-1)
        assert np.concatenate(got).tobytes() == want.tobytes()

    def wh(ops, r):
        a, lo = shard(u8, r)
        idx, off, total = ops.where(a, lo)
        return idx.to_numpy().reshape(-1), off, total
    got = run_sharded(af, sh, world, wh)
    want = orc.where(u8)
    assert got[0][2] == want.size
    assert np.concatenate([g[0] for g in got]).tobytes() == want.tobytes()
    for idx, off, total in got:
        assert idx.tobytes() == want[off:off + idx.size].tobytes()


def test_peer_exchange_single_rank(env):
    """afx_peer_exchange with world = 1: the rank's own buffer plays every role (store, flag, wait, host landing);
    the multi-rank protocol itself is exercised by tools/check_sharded.py under torchrun."""
    import ctypes

    import torch
    af, sh, _ = env
    buf = torch.zeros(2048, dtype=torch.uint8, device="cuda")
    ptrs = (ctypes.c_ulonglong * 8)(buf.data_ptr(), 0, 0, 0, 0, 0, 0, 0)
    out = (ctypes.c_ubyte * 512)()
    for epoch in range(1, 6):
        payload = torch.arange(64, dtype=torch.uint8, device="cuda") + epoch
        torch.cuda.synchronize()
        af._check(af.lib.afx_peer_exchange(out, ctypes.c_void_p(payload.data_ptr()), ptrs, ctypes.c_int(0), ctypes.c_int(1),
                                           ctypes.c_uint(epoch)))
        got = np.frombuffer(out, dtype=np.uint8, count=64)
        assert got.tolist() == [(i + epoch) % 256 for i in range(64)], epoch
    # argument checking
    with pytest.raises(af.AfError):
        af._check(af.lib.afx_peer_exchange(out, ctypes.c_void_p(payload.data_ptr()), ptrs, ctypes.c_int(0), ctypes.c_int(9), ctypes.c_uint(1)))
    with pytest.raises(af.AfError):
        af._check(af.lib.afx_peer_exchange(out, ctypes.c_void_p(payload.data_ptr()), ptrs, ctypes.c_int(0), ctypes.c_int(1), ctypes.c_uint(0)))
