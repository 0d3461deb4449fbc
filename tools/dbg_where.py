"""per-CTA cycle counters of pipe_where_kernel (AFB200_PIPE_DEBUG=1): 2^30 u8 mask at 50 % and 1 %"""
import os, sys
os.environ["AFB200_PIPE_DEBUG"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import arrayfire_b200 as af
n = 1 << 30
for p in (0.5, 0.01):
    m = (torch.rand(n, device="cuda") < p).to(torch.uint8)
    g = af.Array.from_torch(m)
    print("density", p, file=sys.stderr)
    for _ in range(2):
        r = af.where(g)
    del r, g, m
