"""one launch per op for an `ncu --set full` capture: python tools/ncu_one.py where|sbk|scan|rbk|pi"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import arrayfire_b200 as af
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(1)
which = sys.argv[1:] or ["where"]
if "where" in which:
    n = 1 << 30
    for p in (0.5, 0.01):
        m = (torch.rand(n, generator=g, device=dev) < p).to(torch.uint8) * 7
        am = af.Array.from_torch(m); torch.cuda.synchronize()
        for _ in range(2): r = af.where(am)
        af.sync(); del r, am, m
if "sbk" in which or "rbk" in which or "scan" in which:
    n = 1 << 28
    v = torch.randint(-15, 15, (n,), generator=g, device=dev, dtype=torch.int32)
    lens = torch.randint(1, 1025, ((1 << 20) + 2,), generator=g, device=dev)
    k = torch.repeat_interleave(torch.arange((1 << 20) + 2, device=dev, dtype=torch.int32), lens)[:n].contiguous()
    av, ak = af.Array.from_torch(v), af.Array.from_torch(k); torch.cuda.synchronize()
    if "sbk" in which:
        for _ in range(2): r = af.scan_by_key(ak, av, 0, af.AF_BINARY_ADD, False)
    if "rbk" in which:
        for _ in range(2): r = af.sum_by_key(ak, av, 0)
    if "scan" in which:
        x = v.to(torch.float32)[: 1 << 26].contiguous(); ax = af.Array.from_torch(x); torch.cuda.synchronize()
        for _ in range(2): r = af.accum(ax, 0)
    af.sync()
if "pi" in which:
    n = 1 << 28
    x = torch.rand(n, device=dev, generator=g); y = torch.rand(n, device=dev, generator=g)
    X, Y = af.Array.from_torch(x), af.Array.from_torch(y); torch.cuda.synchronize()
    for _ in range(2): c = af.sum_all(af.lt(af.sqrt(af.add(af.mul(X, X), af.mul(Y, Y))), 1.0))
