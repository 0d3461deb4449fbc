"""per-op duration vs size on one GPU, COLD: `reps` ops enqueued back to back (launch latency hidden), each on a different
slice of a buffer much larger than the 126 MB L2 (so nothing is served from cache); CUDA events around the batch.
Shows the fixed cost (ramp, tail, last-arriver fold, host publish) of each kernel family.  Usage: floor.py [name-filter ...]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import arrayfire_b200 as af  # noqa: E402

dev = torch.device("cuda", 0)
stream = torch.cuda.ExternalStream(af.get_stream(0), device=dev)
reps = int(os.environ.get("REPS", "32"))
EXPS = [int(e) for e in os.environ.get("EXPS", "22,24,25,26,27,28,29,30").split(",")]


def timeit(fns):
    """fns: list of callables over distinct slices; cycled"""
    for f in fns[:2]:
        r = f()
        if hasattr(r, "result"):
            r.result()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    keep = []
    e0.record(stream)
    for i in range(reps):
        r = fns[i % len(fns)]()
        if hasattr(r, "result"):
            keep.append(r)          # array results are dropped at once: their buffers go back to the pool (stream-ordered)
        del r
    e1.record(stream)
    for r in keep:
        if hasattr(r, "result"):
            r.result()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3   # us


def main():
    only = sys.argv[1:] or None
    print("%-26s %s" % ("op \\ log2(bytes)", " ".join("%8d" % e for e in EXPS)))
    g = torch.Generator(device=dev)
    g.manual_seed(1)
    TOT = 1 << 32                                             # 4 GiB per source buffer
    bigf = torch.randint(-4, 5, (TOT // 4,), device=dev, dtype=torch.int32, generator=g).float()
    bigi = torch.randint(-15, 15, (TOT // 4,), device=dev, dtype=torch.int32, generator=g)
    bigu8 = (torch.rand(TOT, device=dev, generator=g) < 0.5).to(torch.uint8)
    nk = (TOT // 4) // 256 + 2
    keys = torch.repeat_interleave(torch.arange(nk, device=dev, dtype=torch.int32),
                                   torch.randint(1, 1025, (nk,), device=dev, generator=g))[:TOT // 4].contiguous()
    A = af.Array.from_torch

    def slices(src, n):
        k = max(1, min(16, src.numel() // n))
        return [src[i * n:(i + 1) * n] for i in range(k)]

    def mat(t, rows):
        return af.Array.from_device_ptr(t.data_ptr(), (rows, t.numel() // rows), af.f32, keepalive=t)

    ops = [
        ("sum_all f32 (host scalar)", bigf, 4, lambda t: (lambda a: (lambda: af.reduce_all_async(a, af.OP_ADD)))(A(t))),
        ("sum_all_array f32 (dev)", bigf, 4, lambda t: (lambda a: (lambda: af.sum_all_array(a)))(A(t))),
        ("imax_all f32", bigf, 4, lambda t: (lambda a: (lambda: af.ireduce_all_async(a, af.OP_MAX)))(A(t))),
        ("count u8", bigu8, 1, lambda t: (lambda a: (lambda: af.reduce_all_async(a, af.OP_COUNT)))(A(t))),
        ("where u8 50%", bigu8, 1, lambda t: (lambda a: (lambda: af.where_async(a)))(A(t))),
        ("accum f32", bigf, 4, lambda t: (lambda a: (lambda: af.accum(a, 0)))(A(t))),
        ("accum dim1 (8192 rows)", bigf, 4, lambda t: (lambda a: (lambda: af.accum(a, 1)))(mat(t, 8192))),
        ("accum dim1 (1024 rows)", bigf, 4, lambda t: (lambda a: (lambda: af.accum(a, 1)))(mat(t, 1024))),
        ("accum dim1 (32 rows)", bigf, 4, lambda t: (lambda a: (lambda: af.accum(a, 1)))(mat(t, 32))),
        ("mean_all f32", bigf, 4, lambda t: (lambda a: (lambda: af.mean_all(a)))(A(t))),
        ("var_all f32", bigf, 4, lambda t: (lambda a: (lambda: af.var_all(a)))(A(t))),
        ("var dim0 (8192 rows)", bigf, 4, lambda t: (lambda a: (lambda: af.var(a, 0)))(mat(t, 8192))),
        ("norm2 f32", bigf, 4, lambda t: (lambda a: (lambda: af.norm(a)))(A(t))),
        ("sum dim0 (8192 rows)", bigf, 4, lambda t: (lambda a: (lambda: af.sum(a, 0)))(mat(t, 8192))),
        ("sum dim1 (8192 rows)", bigf, 4, lambda t: (lambda a: (lambda: af.sum(a, 1)))(mat(t, 8192))),
        ("sum dim1 (1024 rows)", bigf, 4, lambda t: (lambda a: (lambda: af.sum(a, 1)))(mat(t, 1024))),
        ("max dim1 (1024 rows)", bigf, 4, lambda t: (lambda a: (lambda: af.max(a, 1)))(mat(t, 1024))),
    ]
    for name, src, item, mk in ops:
        if only and not any(o in name for o in only):
            continue
        row = []
        for e in EXPS:
            n = (1 << e) // item
            fns = [mk(t) for t in slices(src, n)]
            row.append("%8.1f" % timeit(fns))
        print("%-26s %s" % (name[:26], " ".join(row)), flush=True)
    if not only or any("sbk" in o for o in only):
        row = []
        for e in EXPS:
            n = (1 << e) // 4
            fns = []
            for kk, vv in zip(slices(keys, n), slices(bigi, n)):
                k, x = A(kk), A(vv)
                fns.append((lambda k, x: (lambda: af.scan_by_key(k, x, 0, af.AF_BINARY_ADD, False)))(k, x))
            row.append("%8.1f" % timeit(fns))
        print("%-26s %s" % ("sbk s32 (bytes of vals)", " ".join(row)), flush=True)


main()
