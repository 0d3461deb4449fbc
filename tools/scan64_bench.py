"""timing of the generic (8-byte) scan path: f64 inclusive add, s64 by-key"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import arrayfire_b200 as af
stream = torch.cuda.ExternalStream(af.get_stream(0))
n = 1 << 26
x = torch.randint(-4, 5, (n,), device="cuda", dtype=torch.int32).to(torch.float64)
a = af.Array.from_torch(x)
k = torch.repeat_interleave(torch.arange(n // 256 + 2, device="cuda", dtype=torch.int64), torch.randint(1, 1025, (n // 256 + 2,), device="cuda"))[:n].contiguous()
v = torch.randint(-15, 15, (n,), device="cuda", dtype=torch.int64)
gk, gv = af.Array.from_torch(k), af.Array.from_torch(v)
for name, fn, nbytes in (("accum f64 2^26", lambda: af.accum(a, 0), 16 * n), ("scan_by_key s64/s64 2^26", lambda: af.scan_by_key(gk, gv, 0, 0, False), 24 * n)):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); fn(); e1.record(stream); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    print("%-28s %.4f ms  %.0f GB/s" % (name, ms, nbytes / ms / 1e6))
