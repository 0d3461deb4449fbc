"""a few launches of sum_by_key (2^28 s32 keys + values) and of the fused pi expression (2^28 f32 x, y): run under
`ncu --metrics gpu__time_duration.sum` to see the per-kernel split"""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import arrayfire_b200 as af

dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(3)
n = 1 << 28
nk = n // 256 + 2
keys = torch.repeat_interleave(torch.arange(nk, device=dev, dtype=torch.int32), torch.randint(1, 1025, (nk,), device=dev, generator=g))[:n].contiguous()
vals = torch.randint(-15, 15, (n,), device=dev, dtype=torch.int32, generator=g)
x = torch.rand(n, device=dev, generator=g); y = torch.rand(n, device=dev, generator=g)
K, V, X, Y = af.Array.from_torch(keys), af.Array.from_torch(vals), af.Array.from_torch(x), af.Array.from_torch(y)
for rep in range(3):
    t0 = time.perf_counter()
    ko, vo = af.sum_by_key(K, V, 0)
    af.sync()
    t1 = time.perf_counter()
    c = af.sum_all(af.lt(af.sqrt(af.add(af.mul(X, X), af.mul(Y, Y))), 1.0))
    t2 = time.perf_counter()
    print("rep %d: sum_by_key %.3f ms (%d runs)   pi %.3f ms (pi ~ %.5f)" % (rep, (t1 - t0) * 1e3, ko.elements, (t2 - t1) * 1e3, 4 * c / n), flush=True)
