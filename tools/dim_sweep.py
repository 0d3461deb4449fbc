"""micro-benchmark of dim_kernel (reduction along dim 1) for several row counts and vector widths"""
import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    import torch
    import arrayfire_b200 as af
    stream = torch.cuda.ExternalStream(af.get_stream(0))
    for rows in (8192, 4096, 1024, 512):
        x = torch.randint(-4, 5, (rows * 8192,), device="cuda", dtype=torch.int32).to(torch.float32)
        a = af.Array.from_device_ptr(x.data_ptr(), (rows, 8192), af.f32, keepalive=x)
        big = torch.empty(1 << 28, device="cuda", dtype=torch.float32)
        for _ in range(3):
            af.sum(a, 1)
        ts = []
        for _ in range(10):
            big.zero_()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream); af.sum(a, 1); e1.record(stream); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[len(ts) // 2]
        print("vec=%s rows=%5d  %.4f ms  %.0f GB/s" % (os.environ.get("AFB200_DIM_VEC", "auto"), rows, ms, rows * 8192 * 4 / ms / 1e6))
else:
    for v in ("0", "4", "2", "1"):
        env = dict(os.environ); env["AFB200_DIM_VEC"] = v
        subprocess.run([sys.executable, __file__, "run"], env=env)
