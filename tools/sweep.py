"""GPU micro-sweeps of libafb200's launch knobs (run under gpurun).  Prints one line per (op, setting):
median ms over `reps` launches timed with CUDA events on the library's stream, and effective GB/s."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import arrayfire_b200 as af  # noqa: E402

dev = torch.device("cuda", 0)
stream = torch.cuda.ExternalStream(af.get_stream(0), device=dev)


def timeit(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        fn()
        e1.record(stream)
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), float(np.min(ts))


def report(name, setting, nbytes, fn, **kw):
    med, mn = timeit(fn, **kw)
    print("%-28s %-34s med %8.4f ms  min %8.4f ms  %8.1f GB/s (min %8.1f)" % (
        name, setting, med, mn, nbytes / med / 1e6, nbytes / mn / 1e6), flush=True)


def main():
    which = sys.argv[1:] or ["reduce", "scan", "where", "dim"]
    g = torch.Generator(device=dev)
    g.manual_seed(1)
    n = 1 << 30
    if "reduce" in which:
        x = torch.randint(-4, 5, (n,), generator=g, device=dev, dtype=torch.int8).to(torch.float32)
        ax = af.Array.from_torch(x)
        u8 = torch.randint(0, 2, (n,), generator=g, device=dev, dtype=torch.uint8)
        au8 = af.Array.from_torch(u8)
        for ctas in (2, 4, 6, 8, 12, 16):
            os.environ["AFB200_REDUCE_CTAS_PER_SM"] = str(ctas)
            for cl in (1, 2):
                os.environ["AFB200_CLUSTER"] = str(cl)
                report("sum_all_array f32 2^30", "ctas/sm=%d cluster=%d" % (ctas, cl), 4 * n, lambda: af.sum_all_array(ax))
            report("imax_all f32 2^30", "ctas/sm=%d" % ctas, 4 * n, lambda: af.imax_all(ax))
            report("count_all_array u8 2^30", "ctas/sm=%d" % ctas, n, lambda: af.count_all_array(au8))
        os.environ.pop("AFB200_REDUCE_CTAS_PER_SM")
        os.environ.pop("AFB200_CLUSTER")
        x1 = x[: 1 << 26]
        a1 = af.Array.from_torch(x1)
        report("sum_all_array f32 2^26", "default", 4 << 26, lambda: af.sum_all_array(a1))
        report("sum_all (scalar) f32 2^26", "default", 4 << 26, lambda: af.sum_all(a1))
        del x, ax, u8, au8
    if "scan" in which:
        x = torch.randint(-4, 5, (1 << 28,), generator=g, device=dev, dtype=torch.int8).to(torch.float32)
        v = torch.randint(-15, 15, (1 << 28,), generator=g, device=dev, dtype=torch.int32)
        lens = torch.randint(1, 1025, ((1 << 20) + 2,), generator=g, device=dev)
        k = torch.repeat_interleave(torch.arange((1 << 20) + 2, device=dev, dtype=torch.int32), lens)[: 1 << 28].contiguous()
        ax, av, ak = af.Array.from_torch(x), af.Array.from_torch(v), af.Array.from_torch(k)
        a26 = af.Array.from_torch(x[: 1 << 26])
        for ctas in (1, 2):
            os.environ["AFB200_CHAIN_CTAS_PER_SM"] = str(ctas)
            report("accum f32 2^28", "chain ctas/sm=%d" % ctas, 8 << 28, lambda: af.accum(ax, 0))
            report("accum f32 2^26", "chain ctas/sm=%d" % ctas, 8 << 26, lambda: af.accum(a26, 0))
            report("scan_by_key s32 2^28 excl", "chain ctas/sm=%d" % ctas, 12 << 28,
                   lambda: af.scan_by_key(ak, av, 0, af.AF_BINARY_ADD, False))
        os.environ.pop("AFB200_CHAIN_CTAS_PER_SM")
        os.environ["AFB200_NO_CHAIN"] = "1"
        report("accum f32 2^28", "one-tile-per-CTA look-back", 8 << 28, lambda: af.accum(ax, 0))
        report("scan_by_key s32 2^28 excl", "one-tile-per-CTA look-back", 12 << 28,
               lambda: af.scan_by_key(ak, av, 0, af.AF_BINARY_ADD, False))
        os.environ.pop("AFB200_NO_CHAIN")
        del x, v, k, ax, av, ak
    if "where" in which:
        for p in (0.5, 0.01, 0.0, 1.0):
            m = (torch.rand(n, generator=g, device=dev) < p).to(torch.uint8) * 7
            am = af.Array.from_torch(m)
            cnt = int(af.count_all(am))
            for ctas in (1, 2):
                os.environ["AFB200_CHAIN_CTAS_PER_SM"] = str(ctas)
                report("where u8 2^30 p=%.2f" % p, "chain ctas/sm=%d" % ctas, n + 4 * cnt, lambda: af.where(am), reps=10)
            os.environ.pop("AFB200_CHAIN_CTAS_PER_SM")
            os.environ["AFB200_NO_CHAIN"] = "1"
            report("where u8 2^30 p=%.2f" % p, "one-tile-per-CTA", n + 4 * cnt, lambda: af.where(am), reps=10)
            os.environ.pop("AFB200_NO_CHAIN")
            del m, am
    if "dim" in which:
        m = torch.randint(-4, 5, (8192 * 8192,), generator=g, device=dev, dtype=torch.int8).to(torch.float32)
        am = af.Array.from_device_ptr(m.data_ptr(), (8192, 8192), af.f32, keepalive=m)
        nb = 4 * 8192 * 8192
        for name, fn in (("sum dim0", lambda: af.sum(am, 0)), ("sum dim1", lambda: af.sum(am, 1)),
                         ("max dim0", lambda: af.max(am, 0)), ("max dim1", lambda: af.max(am, 1)),
                         ("imax dim0", lambda: af.imax(am, 0)), ("imax dim1", lambda: af.imax(am, 1)),
                         ("accum dim0", lambda: af.accum(am, 0)), ("accum dim1", lambda: af.accum(am, 1))):
            report("8192x8192 f32 " + name, "default", nb * (2 if "accum" in name else 1), fn)


if __name__ == "__main__":
    main()
