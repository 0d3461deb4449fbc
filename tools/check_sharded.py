"""multi-GPU parity check (run under torchrun, one rank per GPU): every sharded whole-array op of
arrayfire_b200/sharded.py over a REAL process group (NCCL / peer exchange) against the oracle on the global array.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_sharded.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    os.environ.setdefault("AF_CUDA_DEFAULT_DEVICE", str(local))
    import arrayfire_b200 as af
    from arrayfire_b200 import sharded as sh
    from oracle import pyoracle as po
    af.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ops = sh.ShardedOps(dist)
    orc = po.AfOracle("port")
    rng = np.random.default_rng(4242)          # same seed on every rank: everybody builds the same global arrays
    n = 5 * (1 << 20) + 12345
    x = rng.integers(-4, 5, n).astype(np.float32)
    ties = (rng.integers(0, 50, n) / 64.0).astype(np.float32)
    iv = rng.integers(-15, 15, n).astype(np.int32)
    keys = np.repeat(np.arange(n // 200 + 2), rng.integers(1, 1025, n // 200 + 2))[:n].astype(np.int32)
    keys[n // 3: n // 3 + 2 * (n // world)] = 7          # one run spanning several shards
    u8 = (rng.random(n) < 0.3).astype(np.uint8) * rng.integers(1, 256, n).astype(np.uint8)
    lo, hi = sh.shard_bounds(n, world, rank)
    A = af.Array.from_numpy
    bad = []

    def check(name, ok):
        if not ok:
            bad.append(name)

    # other dtypes through the fused / exchanged reductions (f64, s64 wrap-free, u8 min / max, complex sum)
    xd = rng.integers(-4, 5, n).astype(np.float64)
    xl = rng.integers(-2 ** 40, 2 ** 40, n).astype(np.int64)
    xc = (rng.integers(-4, 5, n) + 1j * rng.integers(-4, 5, n)).astype(np.complex64)
    check("sum_f64", float(ops.reduce_all(sh.OP_SUM, A(xd[lo:hi]))) == float(orc.sum_all(xd)[1][0]))
    check("sum_s64", int(ops.reduce_all(sh.OP_SUM, A(xl[lo:hi]))) == int(float(orc.sum_all(xl)[1][0])))
    check("min_u8", int(ops.reduce_all(sh.OP_MIN, A(u8[lo:hi]))) == int(u8.min()) and int(ops.reduce_all(sh.OP_MAX, A(u8[lo:hi]))) == int(u8.max()))
    check("sum_c32", complex(ops.reduce_all(sh.OP_SUM, A(xc[lo:hi]))) == complex(orc.sum_all(xc, scalar_api=False)[1][0]))
    check("any_all", int(ops.reduce_all(sh.OP_ANY, A(u8[lo:hi]))) == 1 and int(ops.reduce_all(sh.OP_ALL, A(u8[lo:hi]))) == 0)
    # NaN seeds at shard starts must not become candidates (only global element 0 is the seed)
    tn = ties.copy()
    for r in range(world):
        tn[sh.shard_bounds(n, world, r)[0]] = np.nan
    for op, fn in ((sh.OP_MAX, orc.imax_all), (sh.OP_MIN, orc.imin_all)):
        v, i = ops.ireduce_all(op, A(tn[lo:hi]), lo)
        wv, wi = fn(tn)
        check("ireduce_nan%d" % op, (float(v), int(i)) == (float(wv), int(wi)))
    # scans: min / max by key, f32 plain max
    for binop in (2, 3):
        got = ops.scan_by_key(A(keys[lo:hi]), A(iv[lo:hi]), binop, True).to_numpy().reshape(-1)
        check("sbk_op%d" % binop, got.tobytes() == orc.scan_by_key(keys, iv, 0, binop, True)[1].reshape(-1)[lo:hi].tobytes())
    got = ops.scan(A(x[lo:hi]), 3, True).to_numpy().reshape(-1)
    check("scan_max", got.tobytes() == orc.scan(x, 0, 3, True)[1].reshape(-1)[lo:hi].tobytes())
    k1 = np.zeros(n, np.int32)                 # ONE run over all shards: the carry chains through every rank
    got = ops.scan_by_key(A(k1[lo:hi]), A(iv[lo:hi]), 0, True).to_numpy().reshape(-1)
    check("sbk_one_run", got.tobytes() == np.cumsum(iv, dtype=np.int64).astype(np.int32)[lo:hi].tobytes())

    for rep in range(3):   # repeated: epochs / persistent buffers must stay consistent
        check("sum", np.float32(ops.reduce_all(sh.OP_SUM, A(x[lo:hi]))).tobytes() == orc.sum_all(x, scalar_api=False)[1][:1].tobytes())
        check("count", int(ops.reduce_all(sh.OP_COUNT, A(u8[lo:hi]))) == int(orc.count_all(u8)[1][0]))
        for op, fn in ((sh.OP_MAX, orc.imax_all), (sh.OP_MIN, orc.imin_all)):
            v, i = ops.ireduce_all(op, A(ties[lo:hi]), lo)
            wv, wi = fn(ties)
            check("ireduce%d" % op, (float(v), int(i)) == (float(wv), int(wi)))
        for incl in (True, False):
            got = ops.scan(A(x[lo:hi]), 0, incl).to_numpy().reshape(-1)
            check("scan%d" % incl, got.tobytes() == orc.scan(x, 0, 0, incl)[1].reshape(-1)[lo:hi].tobytes())
            got = ops.scan_by_key(A(keys[lo:hi]), A(iv[lo:hi]), 0, incl).to_numpy().reshape(-1)
            check("sbk%d" % incl, got.tobytes() == orc.scan_by_key(keys, iv, 0, 0, incl)[1].reshape(-1)[lo:hi].tobytes())
        idx, off, total = ops.where(A(u8[lo:hi]), lo)
        want = orc.where(u8)
        got = idx.to_numpy().reshape(-1)
        check("where", total == want.size and got.tobytes() == want[off:off + got.size].tobytes())
    t = torch.tensor([len(bad)], device="cuda")
    dist.all_reduce(t)
    if rank == 0:
        print("check_sharded world=%d peer=%s fused=%s: %s" % (world, getattr(ops.local, "_peer", None) is not None,
                                                               getattr(ops.local, "_comm", None) is not None,
                                                               "OK (every sharded op bit-exact against the oracle on the global array)"
                                                               if int(t.item()) == 0 else "FAILED"), flush=True)
    if bad:
        print("  rank %d failures: %s" % (rank, bad), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    return 0 if int(t.item()) == 0 else 1


if __name__ == "__main__":
    sys.exit(main())
