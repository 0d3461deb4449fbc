"""debug harness: per-CTA cycle counters of pipe_scan_kernel (AFB200_PIPE_DEBUG=1)"""
import os, sys, time
os.environ["AFB200_PIPE_DEBUG"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import arrayfire_b200 as af
n = 1 << 26
x = torch.randint(-4, 5, (n,), device="cuda", dtype=torch.int32).to(torch.float32)
g = af.Array.from_torch(x)
for _ in range(3):
    r = af.accum(g, 0)
k = torch.repeat_interleave(torch.arange(n // 256 + 2, device="cuda", dtype=torch.int32), torch.randint(1, 1025, (n // 256 + 2,), device="cuda"))[:n].contiguous()
v = torch.randint(-15, 15, (n,), device="cuda", dtype=torch.int32)
gk, gv = af.Array.from_torch(k), af.Array.from_torch(v)
for _ in range(2):
    r = af.scan_by_key(gk, gv, 0, af.AF_BINARY_ADD, False)
