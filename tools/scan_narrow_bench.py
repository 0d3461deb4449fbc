"""timing of the promoting scans: af_accum of u8 (-> u32) and s16 (-> s32), 2^28 elements"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import arrayfire_b200 as af
stream = torch.cuda.ExternalStream(af.get_stream(0))
n = 1 << 28
for name, t, isz in (("accum u8 -> u32", torch.randint(0, 256, (n,), device="cuda", dtype=torch.int16).to(torch.uint8), 1),
                     ("accum s16 -> s32", torch.randint(-100, 100, (n,), device="cuda", dtype=torch.int16), 2)):
    a = af.Array.from_torch(t)
    for _ in range(3):
        af.accum(a, 0)
    ts = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); af.accum(a, 0); e1.record(stream); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    print("%-20s %.4f ms  %.0f GB/s" % (name, ms, (isz + 4) * n / ms / 1e6))
