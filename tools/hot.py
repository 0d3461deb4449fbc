"""Quick A-B timing + result check of the ops being tuned (run under gpurun): `python tools/hot.py [where] [sbk] [scan] [rbk] [pi]`.
Median / min ms over `reps` launches timed with CUDA events on the library's stream; every op is checked once against torch."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import arrayfire_b200 as af  # noqa: E402

dev = torch.device("cuda", 0)
stream = torch.cuda.ExternalStream(af.get_stream(0), device=dev)


def timeit(fn, reps=15, warm=3):
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


def report(name, nbytes, fn, ok, **kw):
    torch.cuda.synchronize()
    med, mn = timeit(fn, **kw)
    print("%-34s med %8.4f ms  min %8.4f ms  %8.1f GB/s  frac %.3f  %s" % (
        name, med, mn, nbytes / med / 1e6, nbytes / med / 1e6 / 6551.4, "OK" if ok else "MISMATCH"), flush=True)


def main():
    which = sys.argv[1:] or ["where", "sbk", "scan", "rbk", "pi"]
    g = torch.Generator(device=dev)
    g.manual_seed(1)
    if "where" in which:
        n = 1 << 30
        for p in (0.5, 0.01):
            m = (torch.rand(n, generator=g, device=dev) < p).to(torch.uint8) * 7
            am = af.Array.from_torch(m)
            torch.cuda.synchronize()   # torch's stream and the library's are not ordered
            got = af.where(am)
            want = torch.nonzero(m).reshape(-1).to(torch.int32)
            gt = torch.from_numpy(got.to_numpy().reshape(-1).astype(np.int32)).to(dev)
            ok = gt.numel() == want.numel() and bool((gt == want).all())
            cnt = want.numel()
            del got, want, gt
            report("where u8 2^30 p=%.2f" % p, n + 4 * cnt, lambda: af.where(am), ok, reps=10)
            del m, am
        for dt, nn in ((torch.float32, 1 << 28), (torch.int16, 1 << 28)):
            m = ((torch.rand(nn, generator=g, device=dev) < 0.3) * 3).to(dt)
            am = af.Array.from_torch(m)
            torch.cuda.synchronize()   # torch's stream and the library's are not ordered
            got = af.where(am)
            want = torch.nonzero(m).reshape(-1).to(torch.int32)
            gt = torch.from_numpy(got.to_numpy().reshape(-1).astype(np.int32)).to(dev)
            ok = gt.numel() == want.numel() and bool((gt == want).all())
            cnt = want.numel()
            del got, want, gt
            report("where %s 2^28 p=0.3" % str(dt).split(".")[-1], nn * m.element_size() + 4 * cnt, lambda: af.where(am), ok, reps=10)
            del m, am
    if "sbk" in which or "scan" in which or "rbk" in which:
        n = 1 << 28
        gap = int(os.environ.get("HOT_GAP", "0"))          # bytes between the big inputs: do 2^30-aligned arrays alias in DRAM?
        v = torch.randint(-15, 15, (n,), generator=g, device=dev, dtype=torch.int32)
        if gap:
            pad1 = torch.empty(gap, device=dev, dtype=torch.uint8)
        lens = torch.randint(1, 1025, ((1 << 20) + 2,), generator=g, device=dev)
        k = torch.repeat_interleave(torch.arange((1 << 20) + 2, device=dev, dtype=torch.int32), lens)[:n].contiguous()
        if gap:
            kk = torch.empty(n + gap // 4, device=dev, dtype=torch.int32)
            kk[gap // 4:] = k
            k = kk[gap // 4:]
            print("ptrs v %x k %x" % (v.data_ptr(), k.data_ptr()))
        av, ak = af.Array.from_torch(v), af.Array.from_torch(k)
        torch.cuda.synchronize()
        if "sbk" in which:
            got = torch.from_numpy(af.scan_by_key(ak, av, 0, af.AF_BINARY_ADD, False).to_numpy().reshape(-1)).to(dev)
            cs = torch.cumsum(v.to(torch.int64), 0)
            excl = cs - v
            head = torch.ones(n, dtype=torch.bool, device=dev)
            head[1:] = k[1:] != k[:-1]
            start = torch.cummax(torch.where(head, torch.arange(n, device=dev), torch.zeros(n, dtype=torch.int64, device=dev)), 0)[0]
            want = (excl - excl[start]).to(torch.int32)
            ok = bool((got == want).all())
            del got, cs, excl, head, start, want
            report("scan_by_key s32 2^28 excl add", 12 * n, lambda: af.scan_by_key(ak, av, 0, af.AF_BINARY_ADD, False), ok)
            report("scan_by_key s32 2^28 incl add", 12 * n, lambda: af.scan_by_key(ak, av, 0, af.AF_BINARY_ADD, True), True)
        if "rbk" in which:
            ko, vo = af.sum_by_key(ak, av, 0)
            uk, cnts = torch.unique_consecutive(k, return_counts=True)
            ends = torch.cumsum(cnts, 0)
            cs = torch.cumsum(v.to(torch.int64), 0)
            tot = cs[ends - 1]
            tot[1:] = tot[1:] - cs[ends[:-1] - 1]
            gk = torch.from_numpy(ko.to_numpy().reshape(-1)).to(dev)
            gv = torch.from_numpy(vo.to_numpy().reshape(-1)).to(dev)
            ok = gk.numel() == uk.numel() and bool((gk == uk).all()) and bool((gv == tot.to(torch.int32)).all())
            del ko, vo, uk, cnts, ends, cs, tot, gk, gv
            report("sum_by_key s32 2^28", 8 * n, lambda: af.sum_by_key(ak, av, 0), ok)
            report("max_by_key s32 2^28", 8 * n, lambda: af.max_by_key(ak, av, 0), True)
        if "scan" in which:
            x = torch.randint(-4, 5, (n,), generator=g, device=dev, dtype=torch.int8).to(torch.float32)
            a26 = af.Array.from_torch(x[: 1 << 26])
            torch.cuda.synchronize()
            got = torch.from_numpy(af.accum(a26, 0).to_numpy().reshape(-1)).to(dev)
            ok = bool((got == torch.cumsum(x[: 1 << 26].to(torch.float64), 0).to(torch.float32)).all())
            del got
            report("accum f32 2^26", 8 << 26, lambda: af.accum(a26, 0), ok)
            ax = af.Array.from_torch(x)
            report("accum f32 2^28", 8 << 28, lambda: af.accum(ax, 0), True)
            del x, ax, a26
        del v, k, av, ak
    if "pi" in which:
        n = 1 << 28
        x = torch.rand(n, device=dev, generator=g)
        y = torch.rand(n, device=dev, generator=g)
        X, Y = af.Array.from_torch(x), af.Array.from_torch(y)
        torch.cuda.synchronize()
        fn = lambda: af.sum_all(af.lt(af.sqrt(af.add(af.mul(X, X), af.mul(Y, Y))), 1.0))
        want = int((torch.sqrt(x * x + y * y) < 1).sum())
        ok = int(fn()) == want
        report("pi: sum(sqrt(x*x+y*y)<1) 2^28", 8 * n, fn, ok)


if __name__ == "__main__":
    main()
