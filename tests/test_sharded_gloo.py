"""world_size-2 (and 3) `gloo` tests of the multi-GPU layer, on CPU.

arrayfire_b200/sharded.py's ShardedOps is run with an ORACLE-backed stand-in for the shard-local GPU
primitives (same payload layout as LocalGpu / the afx_* entry points), so the packing, the all-gather and the
rank-ordered combine rules -- the code that runs unchanged over NCCL on the GPU box -- are checked here against
the oracle's answer on the WHOLE array: bit-exact for every op (the float inputs are integer-valued, so any
association order gives the same bits).
"""
import contextlib
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


class LocalOracle:
    """shard-local primitives computed by the oracle; mirrors arrayfire_b200.sharded.LocalGpu"""

    def __init__(self):
        import torch

        from oracle import pyoracle as po
        self.torch, self.po = torch, po
        self.orc = po.AfOracle("port")

    def payload(self):
        return self.torch.zeros(64, dtype=self.torch.uint8)

    def on_stream(self):
        return contextlib.nullcontext()

    def elements(self, a):
        return int(a.size)

    def dtype(self, a):
        return self.po.code_of(a)

    def _put(self, buf, off, raw):
        buf[off:off + len(raw)] = self.torch.frombuffer(bytearray(raw), dtype=self.torch.uint8)

    def reduce_partial(self, buf, a, op):
        fn = {0: lambda: self.orc.sum_all(a, scalar_api=False), 1: lambda: self.orc.product_all(a, scalar_api=False),
              2: lambda: self.orc.min_all(a), 3: lambda: self.orc.max_all(a), 4: lambda: self.orc.count_all(a),
              5: lambda: self.orc.all_true_all(a), 6: lambda: self.orc.any_true_all(a)}[op]
        code, v = fn()
        self._put(buf, 0, np.asarray(v).reshape(-1)[:1].tobytes())
        return code

    def put_base(self, buf, base):
        self._put(buf, 48, np.uint64(base).tobytes())

    def ireduce_partial(self, buf, a, op):
        v, i = (self.orc.imin_all if op == 2 else self.orc.imax_all)(a)
        self._put(buf, 0, np.asarray(v, dtype=a.dtype).tobytes())
        self._put(buf, 16, np.uint32(i).tobytes())

    @staticmethod
    def _join(binop, carry, x):
        if binop == 0:
            return (x + carry).astype(x.dtype)
        if binop == 1:
            return (x * carry).astype(x.dtype)
        if binop == 2:
            return np.minimum(x, carry)
        return np.maximum(x, carry)

    def scan_with_carry(self, a, binop, inclusive, carry_bytes):
        code, out = self.orc.scan(a, 0, binop, inclusive)
        out = out.reshape(-1)
        if carry_bytes is not None:
            out = self._join(binop, np.frombuffer(carry_bytes, dtype=out.dtype)[0], out)
        return out

    def scan_by_key_local(self, buf, key, a, binop, inclusive):
        code, out = self.orc.scan_by_key(key, a, 0, binop, inclusive)
        out = out.reshape(-1).copy()
        heads = np.flatnonzero(key[1:] != key[:-1]) + 1
        start = int(heads[-1]) if heads.size else 0
        tail_code, tail = {0: self.orc.sum_all, 1: self.orc.product_all}.get(binop, None)(a[start:].astype(out.dtype), scalar_api=False) \
            if binop < 2 else ((self.orc.min_all if binop == 2 else self.orc.max_all)(a[start:].astype(out.dtype)))
        self._put(buf, 0, np.asarray(tail).reshape(-1)[:1].astype(out.dtype).tobytes())
        self._put(buf, 16, np.uint32(1 if heads.size else 0).tobytes())
        self._put(buf, 24, np.uint64(heads[0] if heads.size else 0xffffffffffffffff).tobytes())
        self._put(buf, 32, key[:1].tobytes())
        self._put(buf, 40, key[-1:].tobytes())
        return out

    def scan_fixup(self, out, carry_bytes, count, binop):
        c = np.frombuffer(carry_bytes, dtype=out.dtype)[0]
        out[:count] = self._join(binop, c, out[:count])

    def where_shard(self, a, base):
        idx = self.orc.where(a)
        return (idx + np.uint32(base)).astype(np.uint32), int(idx.size)

    def host_scalar_payload(self, raw):
        buf = self.payload()
        self._put(buf, 0, raw)
        return buf


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, outdir):
    import torch.distributed as dist

    from arrayfire_b200 import sharded as sh
    from oracle import pyoracle as po
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        orc = po.AfOracle("port")
        ops = sh.ShardedOps(dist, LocalOracle())
        rng = np.random.default_rng(0xAF00)  # the same global arrays on every rank
        n = 100003
        lo, hi = sh.shard_bounds(n, world, rank)
        res = {}

        x = rng.integers(-4, 5, n).astype(np.float32)
        i32 = rng.integers(-2**31, 2**31 - 1, n).astype(np.int32)       # wraps: compare modulo 2^32
        u8 = (rng.random(n) < 0.3).astype(np.uint8) * rng.integers(1, 256, n).astype(np.uint8)
        ties = (rng.integers(0, 50, n) / 64.0).astype(np.float32)          # many exact ties for imin/imax
        for name, arr, op, want in (
                ("sum_f32", x, sh.OP_SUM, orc.sum_all(x, scalar_api=False)[1][0]),
                ("sum_s32", i32, sh.OP_SUM, orc.sum_all(i32, scalar_api=False)[1][0]),
                ("min_f32", x, sh.OP_MIN, orc.min_all(x)[1][0]),
                ("max_s32", i32, sh.OP_MAX, orc.max_all(i32)[1][0]),
                ("count_u8", u8, sh.OP_COUNT, orc.count_all(u8)[1][0]),
                ("sum_u8", u8, sh.OP_SUM, orc.sum_all(u8, scalar_api=False)[1][0]),
                ("any_u8", u8, sh.OP_ANY, orc.any_true_all(u8)[1][0]),
                ("all_u8", u8, sh.OP_ALL, orc.all_true_all(u8)[1][0])):
            with np.errstate(over="ignore"):
                got = ops.reduce_all(op, arr[lo:hi])
            assert np.asarray(got).tobytes() == np.asarray(want).tobytes(), (name, got, want)
        for op, fn in ((sh.OP_MAX, orc.imax_all), (sh.OP_MIN, orc.imin_all)):
            for arr in (ties, i32):
                v, i = ops.ireduce_all(op, arr[lo:hi], lo)
                wv, wi = fn(arr)
                assert (float(v), int(i)) == (float(wv), int(wi)), (op, v, i, wv, wi)

        for binop in range(4):
            for incl in (True, False):
                src = x if binop != 1 else rng.integers(1, 3, n).astype(np.int32) // 2 + 1  # products of 1s and 2s... keep small
                if binop == 1:
                    src = np.where(rng.random(n) < 0.0002, 2, 1).astype(np.int32)
                got = ops.scan(src[lo:hi], binop, incl)
                want = orc.scan(src, 0, binop, incl)[1].reshape(-1)[lo:hi]
                assert np.asarray(got).tobytes() == want.tobytes(), ("scan", binop, incl)
        # narrow input: totals are promoted to the scan's output type before they become the carry
        got = ops.scan(u8[lo:hi], 0, True)
        assert got.tobytes() == orc.scan(u8, 0, 0, True)[1].reshape(-1)[lo:hi].tobytes()

        vals = rng.integers(-15, 15, n).astype(np.int32)
        cases = {
            "short_runs": np.repeat(np.arange(n // 4 + 2), rng.integers(1, 9, n // 4 + 2))[:n],
            "runs_1024": np.repeat(np.arange(n // 200 + 2), rng.integers(1, 1025, n // 200 + 2))[:n],
            "one_run": np.zeros(n),                                         # a run spanning every shard
            "break_at_boundary": (np.arange(n) >= sh.shard_bounds(n, world, 1)[0]).astype(np.int64),
            "giant_middle": np.where((np.arange(n) > n // 7) & (np.arange(n) < n - n // 9), 7, np.arange(n) % 3),
        }
        for cname, keys in cases.items():
            keys = keys.astype(np.int32)
            for binop in (0, 2, 3):
                for incl in (True, False):
                    got = ops.scan_by_key(keys[lo:hi], vals[lo:hi], binop, incl)
                    want = orc.scan_by_key(keys, vals, 0, binop, incl)[1].reshape(-1)[lo:hi]
                    assert np.asarray(got).tobytes() == want.tobytes(), ("scan_by_key", cname, binop, incl)
        fkeys = cases["runs_1024"].astype(np.int64)
        got = ops.scan_by_key(fkeys[lo:hi], x[lo:hi], 0, False)
        assert got.tobytes() == orc.scan_by_key(fkeys, x, 0, 0, False)[1].reshape(-1)[lo:hi].tobytes()

        for mask in (u8, np.zeros(n, np.uint8), (rng.random(n) < 0.01).astype(np.uint8)):
            idx, off, total = ops.where(mask[lo:hi], lo)
            want = orc.where(mask)
            assert total == want.size
            assert idx.tobytes() == want[off:off + idx.size].tobytes()
        res["ok"] = True
        with open(os.path.join(outdir, "rank%d.ok" % rank), "w") as f:
            f.write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_ops_gloo(world, tmp_path):
    import torch.multiprocessing as mp
    from oracle import pyoracle as po
    po.build()
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert os.path.exists(os.path.join(str(tmp_path), "rank%d.ok" % r))


def test_shard_bounds_cover():
    from arrayfire_b200 import sharded as sh
    for n in (0, 1, 7, 100003, 1 << 30):
        for world in (1, 2, 3, 8):
            prev = 0
            for r in range(world):
                lo, hi = sh.shard_bounds(n, world, r)
                assert lo == prev and hi >= lo
                prev = hi
            assert prev == n


def test_segmented_carry_rules():
    from arrayfire_b200 import sharded as sh
    # run spanning shards 0..2, shard 1 is a single segment: its carry is passed through
    tails = np.array([5, 7, 11, 13], np.int32)
    single = [False, True, False, False]
    first = [1, 9, 9, 4]
    last = [9, 9, 3, 4]
    assert sh.segmented_carry(0, tails, single, first, last, 0) is None
    assert int(sh.segmented_carry(0, tails, single, first, last, 1)) == 5
    assert int(sh.segmented_carry(0, tails, single, first, last, 2)) == 12
    assert sh.segmented_carry(0, tails, single, first, last, 3) is None
