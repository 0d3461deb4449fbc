"""bench_gpu.py -- the GPU arm of bench.py (kept separate so `bench.py --impl reference` never imports torch.cuda).

Per rank: build the synthetic inputs of BASELINE.json's configs C1..C5 in HBM, run W warm-up steps and K timed
steps of the suite through libafb200's C-ABI (arrayfire_b200), time them with CUDA events recorded on the
library's own stream (afcu_get_stream), take the max over ranks, print ONE JSON line on rank 0.
"""
import json
import os
import sys
import time

import numpy as np

from bench import (OP_ORDER, SEED, WORKLOAD, ClockSampler, algorithmic_bytes, log, run_cpu, workload)

ROOT = os.path.dirname(os.path.abspath(__file__))

# ops whose dominant launch is span_kernel (the whole-array / long-column streaming reduction)
SPAN_OPS = ["C3.sum_f32", "C3.sum_s32", "C3.imax_f32", "C1.sum_f32", "C5.count"]


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, STREAM-style copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


# ------------------------------------------------------------------------------------------------
# synthetic inputs, generated on the device (per rank: its shard of the sharded configs)
# ------------------------------------------------------------------------------------------------
def make_inputs(torch, w, rank, world, dev):
    """returns dict of torch tensors (device).  Whole-array configs hold the rank's contiguous shard of the
    GLOBAL array (strong scaling); C2 holds the rank's column / row block."""
    from arrayfire_b200.sharded import shard_bounds
    g = torch.Generator(device=dev)
    g.manual_seed(SEED + 7919 * rank)

    def small_ints(n):  # integer-valued floats in {-4..4}: every partial sum is exact
        out = torch.empty(n, dtype=torch.float32, device=dev)
        step = 1 << 28
        for lo in range(0, n, step):
            m = min(step, n - lo)
            out[lo:lo + m] = torch.randint(-4, 5, (m,), generator=g, device=dev, dtype=torch.int8).to(torch.float32)
        return out

    def shard(n):
        lo, hi = shard_bounds(n, world, rank)
        return hi - lo, lo

    d = {}
    n1, b1 = shard(w["C1"]["n"])
    d["x1"], d["x1_base"] = small_ints(n1), b1
    rows, cols = w["C2"]["rows"], w["C2"]["cols"]
    # dim-0 reductions keep whole columns: the rank owns a block of columns; dim-1 reductions a block of rows.
    c_lo, c_hi = shard_bounds(cols, world, rank)
    r_lo, r_hi = shard_bounds(rows, world, rank)
    d["m2_cols"] = small_ints(rows * (c_hi - c_lo))          # column-major [rows, c_hi-c_lo]
    d["m2_cols_dims"] = (rows, c_hi - c_lo)
    d["m2_rows"] = small_ints((r_hi - r_lo) * cols)          # column-major [r_hi-r_lo, cols]
    d["m2_rows_dims"] = (r_hi - r_lo, cols)
    n3, b3 = shard(w["C3"]["n"])
    d["x3"], d["x3_base"] = small_ints(n3), b3
    i3 = torch.empty(n3, dtype=torch.int32, device=dev)
    u3 = torch.empty(n3, dtype=torch.float32, device=dev)
    step = 1 << 28
    for lo in range(0, n3, step):
        m = min(step, n3 - lo)
        i3[lo:lo + m] = torch.randint(-1000, 1001, (m,), generator=g, device=dev, dtype=torch.int32)
        u3[lo:lo + m] = torch.randint(0, 1 << 24, (m,), generator=g, device=dev, dtype=torch.int32).to(torch.float32) / float(1 << 24)
    d["i3"], d["u3"] = i3, u3
    n4, b4 = shard(w["C4"]["n"])
    d["v4"] = torch.randint(-15, 15, (n4,), generator=g, device=dev, dtype=torch.int32)
    nruns = n4 // 256 + 2
    lens = torch.randint(1, 1025, (nruns,), generator=g, device=dev, dtype=torch.int64)
    # run ids are globally distinct per rank except that the first run of rank r continues the last key of
    # rank r-1 (a run crossing the shard boundary), so the cross-GPU carry path is exercised
    ids = torch.arange(nruns, device=dev, dtype=torch.int32) + (rank * (nruns - 1))
    d["k4"] = torch.repeat_interleave(ids, lens)[:n4].contiguous()
    del lens, ids
    n5, b5 = shard(w["C5"]["n"])
    d["c5_base"] = b5
    for name, p in (("m50", 0.5), ("m01", 0.01)):
        out = torch.empty(n5, dtype=torch.uint8, device=dev)
        for lo in range(0, n5, step):
            m = min(step, n5 - lo)
            nz = torch.randint(1, 256, (m,), generator=g, device=dev, dtype=torch.int16).to(torch.uint8)
            keep = torch.rand(m, generator=g, device=dev) < p
            out[lo:lo + m] = torch.where(keep, nz, torch.zeros_like(nz))
        d[name] = out
    torch.cuda.synchronize()
    return d


# ------------------------------------------------------------------------------------------------
# the suite: name -> callable returning (result object or None); device-resident inputs
# ------------------------------------------------------------------------------------------------
def build_suite(af, torch, d, rank, world, group, pipelined=True, ops=None):
    """name -> callable that ENQUEUES the op and returns its result: an af.Array (device result, as in the reference API),
    or -- pipelined -- an af.Pending whose .result() collects the scalar / index / hit count later.  pipelined=False uses
    the blocking reference calls (af_sum_all ... af_where: one host round trip per op)."""
    A = af.Array
    x1 = A.from_torch(d["x1"])
    m2c = A.from_device_ptr(d["m2_cols"].data_ptr(), d["m2_cols_dims"], af.f32, keepalive=d["m2_cols"])
    m2r = A.from_device_ptr(d["m2_rows"].data_ptr(), d["m2_rows_dims"], af.f32, keepalive=d["m2_rows"])
    x3, i3, u3 = A.from_torch(d["x3"]), A.from_torch(d["i3"]), A.from_torch(d["u3"])
    k4, v4 = A.from_torch(d["k4"]), A.from_torch(d["v4"])
    m50, m01 = A.from_torch(d["m50"]), A.from_torch(d["m01"])
    keep = [x1, m2c, m2r, x3, i3, u3, k4, v4, m50, m01]
    if world == 1:
        if pipelined:
            red = lambda op, a: af.reduce_all_async(a, op)                      # noqa: E731
            ired = lambda a, base: af.ireduce_all_async(a, af.OP_MAX, None, 0)  # noqa: E731
            whr = lambda a, base: af.where_async(a, None, 0)                    # noqa: E731
        else:
            fn = {af.OP_ADD: af.sum_all, af.OP_COUNT: af.count_all}
            red = lambda op, a: fn[op](a)                                       # noqa: E731
            ired = lambda a, base: af.imax_all(a)                               # noqa: E731
            whr = lambda a, base: af.where(a)                                   # noqa: E731
        scn = lambda a: af.accum(a, 0)                                          # noqa: E731
        sbk = lambda k, v: af.scan_by_key(k, v, 0, af.AF_BINARY_ADD, False)     # noqa: E731
        OP_SUM, OP_COUNT = af.OP_ADD, af.OP_COUNT
    else:
        from arrayfire_b200 import sharded as sh
        if ops is None:
            ops = sh.ShardedOps(group)
        if pipelined:
            red = lambda op, a: ops.reduce_all_async(op, a)                     # noqa: E731
            ired = lambda a, base: ops.ireduce_all_async(sh.OP_MAX, a, base)    # noqa: E731
            whr = lambda a, base: ops.where_async(a, base)                      # noqa: E731
        else:
            red = lambda op, a: ops.reduce_all(op, a)                           # noqa: E731
            ired = lambda a, base: ops.ireduce_all(sh.OP_MAX, a, base)          # noqa: E731
            whr = lambda a, base: ops.where(a, base)                            # noqa: E731
        scn = lambda a: ops.scan(a, af.AF_BINARY_ADD, True)                     # noqa: E731
        sbk = lambda k, v: ops.scan_by_key(k, v, af.AF_BINARY_ADD, False)       # noqa: E731
        OP_SUM, OP_COUNT = sh.OP_SUM, sh.OP_COUNT
        keep.append(ops)
    suite = {
        "C1.sum_f32": lambda: red(OP_SUM, x1),
        "C1.scan_f32": lambda: scn(x1),
        "C2.sum_dim0": lambda: af.sum(m2c, 0),       # N > 1: partitioned along the non-reduced axis, no communication
        "C2.sum_dim1": lambda: af.sum(m2r, 1),
        "C2.max_dim0": lambda: af.max(m2c, 0),
        "C2.max_dim1": lambda: af.max(m2r, 1),
        "C3.sum_f32": lambda: red(OP_SUM, x3),
        "C3.sum_s32": lambda: red(OP_SUM, i3),
        "C3.imax_f32": lambda: ired(u3, d["x3_base"]),
        "C4.scan_by_key_s32": lambda: sbk(k4, v4),
        "C5.where_50": lambda: whr(m50, d["c5_base"]),
        "C5.where_01": lambda: whr(m01, d["c5_base"]),
        "C5.count": lambda: red(OP_COUNT, m50),
    }
    return suite, keep, ops


def as_torch(torch, arr, dtype, dev):
    """zero-copy torch view of a dense af.Array"""
    n = int(arr.elements)
    if n == 0:
        return torch.empty(0, dtype=dtype, device=dev)
    item = torch.empty(0, dtype=dtype).element_size()
    ptr = arr.device_ptr()
    arr.unlock()

    class _Holder:
        __cuda_array_interface__ = {"shape": (n * item,), "typestr": "|u1", "data": (ptr, False), "version": 3, "strides": None}
    return torch.as_tensor(_Holder(), device=dev).view(dtype)


def check_results(af, torch, dist, d, last, rank, world, dev):
    """Every op of the step against an INDEPENDENT value computed by torch from the same device-resident inputs (all-reduced /
    all-gathered over the ranks for the sharded ops).  All inputs are integer valued, so every comparison is exact.
    Returns the list of failed op names (empty = parity holds on this rank)."""
    bad = []

    def allsum(v):
        t = torch.tensor([v], dtype=torch.float64 if isinstance(v, float) else torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(t)
        return t.item()

    def gather(v, dtype=torch.int64):
        t = torch.tensor([v], dtype=dtype, device=dev)
        if world == 1:
            return [t.item()]
        out = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return [o.item() for o in out]

    def chk(name, ok):
        if not bool(ok):
            bad.append(name)

    torch.cuda.synchronize()
    # whole-array reductions
    chk("C1.sum_f32", float(last["C1.sum_f32"]) == allsum(float(d["x1"].sum(dtype=torch.float64).item())))
    chk("C3.sum_f32", float(last["C3.sum_f32"]) == allsum(float(d["x3"].sum(dtype=torch.float64).item())))
    tot = allsum(int(d["i3"].sum(dtype=torch.int64).item()))
    tot = ((tot + (1 << 31)) % (1 << 32)) - (1 << 31)          # s32 accumulator: wraps like the reference's
    chk("C3.sum_s32", int(last["C3.sum_s32"]) == tot)
    chk("C5.count", int(last["C5.count"]) == allsum(int(torch.count_nonzero(d["m50"]).item())))
    # imax: first index of the global maximum
    mx = d["u3"].max()
    first = int((d["u3"] == mx).nonzero()[0].item()) + int(d["x3_base"])
    vals, idxs = gather(float(mx.item()), torch.float64), gather(first)
    gmax = max(vals)
    gidx = min(i for v, i in zip(vals, idxs) if v == gmax)
    v, i = last["C3.imax_f32"]
    chk("C3.imax_f32", float(v) == gmax and int(i) == gidx)
    # C2: column-major blocks; flat buffer viewed row-major as [cols, rows]
    rows_c, cols_c = d["m2_cols_dims"]
    mc = d["m2_cols"].view(cols_c, rows_c)
    chk("C2.sum_dim0", torch.equal(as_torch(torch, last["C2.sum_dim0"], torch.float32, dev), mc.sum(dim=1, dtype=torch.float64).float()))
    chk("C2.max_dim0", torch.equal(as_torch(torch, last["C2.max_dim0"], torch.float32, dev), mc.max(dim=1).values))
    rows_r, cols_r = d["m2_rows_dims"]
    mr = d["m2_rows"].view(cols_r, rows_r)
    chk("C2.sum_dim1", torch.equal(as_torch(torch, last["C2.sum_dim1"], torch.float32, dev), mr.sum(dim=0, dtype=torch.float64).float()))
    chk("C2.max_dim1", torch.equal(as_torch(torch, last["C2.max_dim1"], torch.float32, dev), mr.max(dim=0).values))
    # C1 inclusive scan: local cumsum (exact: |partial sums| << 2^24) + the totals of the lower ranks
    tots = gather(float(d["x1"].sum(dtype=torch.float64).item()), torch.float64)
    want = torch.cumsum(d["x1"].double(), 0) + sum(tots[:rank])
    chk("C1.scan_f32", torch.equal(as_torch(torch, last["C1.scan_f32"], torch.float32, dev), want.float()))
    del want
    # C4 exclusive scan by key: segmented from the plain cumsum; the first run of rank r continues rank r-1's last key
    k, v4 = d["k4"], d["v4"].long()
    n4 = k.numel()
    head = torch.ones(n4, dtype=torch.bool, device=dev)
    head[1:] = k[1:] != k[:-1]
    ex = torch.cumsum(v4, 0) - v4                                 # exclusive prefix of the whole shard
    start = torch.cummax(torch.where(head, torch.arange(n4, device=dev), torch.zeros((), dtype=torch.long, device=dev)), 0).values
    want = ex - ex[start]
    firsts, lasts = gather(int(k[0].item())), gather(int(k[-1].item()))
    last_head = int(start[-1].item())
    tails = gather(int(v4[last_head:].sum().item()))
    nheads = gather(int(head.sum().item()))
    carry, r = 0, rank
    while r > 0 and lasts[r - 1] == firsts[r]:                    # chain through shards that are ONE run of the same key
        carry += tails[r - 1]
        if nheads[r - 1] > 1:
            break
        r -= 1
    if carry:
        first_len = int(head[1:].nonzero()[0].item()) + 1 if nheads[rank] > 1 else n4
        want[:first_len] += carry
    chk("C4.scan_by_key_s32", torch.equal(as_torch(torch, last["C4.scan_by_key_s32"], torch.int32, dev), want.int()))
    del want, ex, start, head
    # where: this rank's slice = local nonzero positions + base; offsets / totals from the counts of all ranks
    for name, key in (("C5.where_50", "m50"), ("C5.where_01", "m01")):
        want = d[key].nonzero().flatten() + int(d["c5_base"])
        cnts = gather(int(want.numel()))
        r = last[name]
        if isinstance(r, tuple):
            arr, off, total = r
            ok = int(off) == sum(cnts[:rank]) and int(total) == sum(cnts)
        else:
            arr, ok = r, True
        ok = ok and int(arr.elements) == int(want.numel()) and torch.equal(as_torch(torch, arr, torch.int32, dev).long() & 0xffffffff, want)
        chk(name, ok)
        del want
    torch.cuda.synchronize()
    return bad


def where_count(res):
    """hit count of a where result (single-GPU Array or sharded (Array, offset, total) tuple)"""
    if isinstance(res, tuple):
        return int(res[2])
    return int(res.elements)


# ------------------------------------------------------------------------------------------------
# end-to-end arm: HOST buffers in, results read back to the host, all inside the timed region
# ------------------------------------------------------------------------------------------------
class E2E:
    """The same suite through the public API with pinned HOST inputs: af_create_array (H2D) -> op ->
    af_get_data_ptr / scalar (D2H) every step.  Array results are read back by ONE helper thread while the main thread
    uploads the next input: the library runs uploads and downloads on dedicated copy streams, so the two PCIe directions
    overlap.  Every call is still the blocking reference call; a step ends when all of its read-backs have landed."""

    def __init__(self, af, torch, d, world, ops=None):
        self.af, self.torch = af, torch
        self.ops, self.d = ops, d      # ops: arrayfire_b200.sharded.ShardedOps when world > 1 (cross-GPU combine included)
        self.h = {}
        for k in ("x1", "m2_cols", "m2_rows", "x3", "i3", "u3", "k4", "v4", "m50", "m01"):
            t = torch.empty(d[k].shape, dtype=d[k].dtype, pin_memory=True)
            t.copy_(d[k])
            self.h[k] = t.numpy()
        torch.cuda.synchronize()
        self.dims = {"m2_cols": d["m2_cols_dims"], "m2_rows": d["m2_rows_dims"]}
        n_out = max(d["v4"].numel(), d["m50"].numel(), d["x1"].numel())
        self.out = torch.empty(n_out, dtype=torch.int32, pin_memory=True).numpy()  # landing zone for array results
        self.h2d = 0
        self.d2h = 0
        from concurrent.futures import ThreadPoolExecutor
        self.reader = ThreadPoolExecutor(max_workers=1)
        self.pending = []

    def _up(self, k):
        a = self.h[k]
        self.h2d += a.nbytes
        if k in self.dims:
            a = a.reshape(self.dims[k], order="F")
        return self.af.Array.from_numpy(a)

    def _down(self, arr):
        n = arr.elements * 4
        self.d2h += n
        # one reader thread: read-backs land one after the other in the pinned zone while the main thread goes on uploading
        self.pending.append(self.reader.submit(arr.host_into, self.out.view(np.uint8)[:max(n, 1)]))

    def _drain(self):
        for f in self.pending:
            f.result()
        self.pending = []

    def step(self):
        if self.ops is not None:
            return self.step_sharded()
        af = self.af
        self.h2d = self.d2h = 0
        counts = {}
        m2c = self._up("m2_cols")            # every distinct input of the step is uploaded ONCE (sum and max share the matrix)
        self._down(af.sum(m2c, 0))
        af.sum_all(self._up("x3")); self.d2h += 4
        m2r = self._up("m2_rows")
        self._down(af.sum(m2r, 1))
        af.sum_all(self._up("i3")); self.d2h += 4
        self._down(af.max(m2c, 0))
        af.imax_all(self._up("u3")); self.d2h += 8
        self._down(af.max(m2r, 1))
        self._down(af.scan_by_key(self._up("k4"), self._up("v4"), 0, af.AF_BINARY_ADD, False))
        x1 = self._up("x1")
        af.sum_all(x1); self.d2h += 4
        m50 = self._up("m50")
        r = af.where(m50); counts["C5.where_50"] = r.elements; self._down(r)
        self._down(af.accum(x1, 0))
        r = af.where(self._up("m01")); counts["C5.where_01"] = r.elements; self._down(r)
        af.count_all(m50); self.d2h += 4
        self._drain()
        return counts

    def step_sharded(self):
        """N > 1: every rank uploads ITS shard from pinned host memory, the whole-array ops go through ShardedOps (kernel +
        cross-GPU combine), array results (the rank's slice) come back to the host"""
        from arrayfire_b200 import sharded as sh
        af, ops, d = self.af, self.ops, self.d
        self.h2d = self.d2h = 0
        counts = {}
        m2c = self._up("m2_cols")
        self._down(af.sum(m2c, 0))
        ops.reduce_all(sh.OP_SUM, self._up("x3")); self.d2h += 4
        m2r = self._up("m2_rows")
        self._down(af.sum(m2r, 1))
        ops.reduce_all(sh.OP_SUM, self._up("i3")); self.d2h += 4
        self._down(af.max(m2c, 0))
        ops.ireduce_all(sh.OP_MAX, self._up("u3"), d["x3_base"]); self.d2h += 8
        self._down(af.max(m2r, 1))
        self._down(ops.scan_by_key(self._up("k4"), self._up("v4"), af.AF_BINARY_ADD, False))
        x1 = self._up("x1")
        ops.reduce_all(sh.OP_SUM, x1); self.d2h += 4
        m50 = self._up("m50")
        r, _, total = ops.where(m50, d["c5_base"]); counts["C5.where_50"] = total; self._down(r)
        self._down(ops.scan(x1, af.AF_BINARY_ADD, True))
        r, _, total = ops.where(self._up("m01"), d["c5_base"]); counts["C5.where_01"] = total; self._down(r)
        ops.reduce_all(sh.OP_COUNT, m50); self.d2h += 4
        self._drain()
        return counts


def bind_to_gpu_numa_node(local):
    """One process per GPU: pin the process (and the threads it starts later) to the CPUs next to ITS GPU, before any pinned
    host buffer is allocated -- first-touch then places those buffers on the GPU's NUMA node, so that 8 ranks uploading at once
    do not funnel their PCIe traffic through one socket's memory controllers (N = 8 end to end ran at 21 GB/s per GPU
    against 64 GB/s at N = 1).  NVML knows the ideal CPU set; failure is not an error (single-socket hosts, containers
    without the permission)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = int(vis.split(",")[local]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else local
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * i + b for i, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if cpus and cpus != allowed:
            os.sched_setaffinity(0, cpus)
            return "%d cpus near gpu %d" % (len(cpus), idx)
        return "all %d cpus (one node)" % len(allowed)
    except Exception as e:   # noqa: BLE001
        return "unbound (%s)" % type(e).__name__


# ------------------------------------------------------------------------------------------------
def gpu_arm(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries the ONE JSON line and nothing else: libraries that print to fd 1 (NCCL's version banner at
    # communicator creation) are sent to stderr for the duration of the run
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            log("bench.py: --gpus %d needs torchrun (one rank per GPU); running 1 GPU" % args.gpus)
            args.gpus = 1
        else:
            args.gpus = world
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the GPU arm has no CPU fallback; use --impl reference)")
    numa = bind_to_gpu_numa_node(local) if world > 1 else None
    # one visible device per process keeps libafb200's device 0 == this rank's GPU
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("AF_CUDA_DEFAULT_DEVICE", str(local))
    import arrayfire_b200 as af
    af.set_device(local)
    group = None
    if world > 1:
        # NCCL's own banner / debug lines go to stdout by default; rank 0's stdout must carry the ONE JSON line only
        if os.environ.get("NCCL_DEBUG") and not os.environ.get("NCCL_DEBUG_FILE"):
            os.environ["NCCL_DEBUG_FILE"] = "/dev/stderr"
        dist.init_process_group("nccl", device_id=dev)
        group = dist
    lib_stream = torch.cuda.ExternalStream(af.get_stream(local), device=dev)

    w = workload(args.scale)
    t0 = time.time()
    d = make_inputs(torch, w, rank, world, dev)
    log("[rank %d] inputs ready in %.1f s" % (rank, time.time() - t0))
    pipelined = not args.sync_ops
    suite, keep, sh_ops = build_suite(af, torch, d, rank, world, group, pipelined)
    e2e_ops = sh_ops

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident arm -------------------------------------------------------------------
    # One step = ENQUEUE the 13 ops back to back, then COLLECT the host-side results (scalars, index, hit counts; the
    # where outputs are trimmed to their count).  Array results stay on the device, as in the reference API.
    counts = {}
    last = {}

    host_issue = [0.0, 0]   # host seconds spent enqueueing, steps

    def run_step(st, ev=None):
        pend = []
        t_issue = time.perf_counter()
        for i, name in enumerate(OP_ORDER):
            if ev is not None:
                ev[i][0].record(lib_stream)
            pend.append(st[name]())
            if ev is not None:
                ev[i][1].record(lib_stream)
        host_issue[0] += time.perf_counter() - t_issue
        host_issue[1] += 1
        for name, r in zip(OP_ORDER, pend):
            if hasattr(r, "result"):
                r = r.result()
            if name.startswith("C5.where"):
                counts[name] = where_count(r)
            last[name] = r

    def step(ev=None):
        run_step(suite, ev)

    for _ in range(args.warmup):
        step()
    barrier()
    # ---- parity: the results of the last warm-up step against torch on the same inputs, on every rank ----
    parity_failed = None
    if not args.no_check:
        bad = check_results(af, torch, dist, d, last, rank, world, dev)
        if world > 1:
            obj = [None] * world
            dist.all_gather_object(obj, bad)
            bad = sorted(set(sum(obj, [])))
        parity_failed = bad
        if bad:
            log("[rank %d] PARITY FAILED: %s" % (rank, bad))
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    af.launch_count(reset=True)
    # events are created up front: nothing but the two records per op is added to the timed region
    evs = [[(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in OP_ORDER]
           for _ in range(args.steps)]
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    host_issue[0], host_issue[1] = 0.0, 0
    start.record(lib_stream)
    for k in range(args.steps):
        step(evs[k])
    end.record(lib_stream)
    barrier()
    launches = af.launch_count()
    host_issue_ms = host_issue[0] / max(1, host_issue[1]) * 1e3
    events = [(name, e0, e1) for ev in evs for name, (e0, e1) in zip(OP_ORDER, ev)]
    if rank == 0 and world == 1 and len(sampler.rows) < 8:
        # the timed region (~150 ms) can be too short for the sampling thread: keep the same load running, untimed,
        # until the clock / throttle-reason record has a few samples taken UNDER LOAD
        t_end = time.time() + 0.5
        while time.time() < t_end and len(sampler.rows) < 16:
            step()
            time.sleep(0)
        torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = start.elapsed_time(end)
    per_op_ms = {}
    for name, e0, e1 in events:
        per_op_ms[name] = per_op_ms.get(name, 0.0) + e0.elapsed_time(e1)

    # max over ranks (device time), where counts summed over ranks
    if world > 1:
        t = torch.tensor([ms_total] + [per_op_ms[n] for n in OP_ORDER], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t[0])
        per_op_ms = {n: float(t[1 + i]) for i, n in enumerate(OP_ORDER)}
        lt = torch.tensor([launches], dtype=torch.int64, device=dev)
        dist.all_reduce(lt)
        launches = int(lt[0])
    ab = algorithmic_bytes(w, counts)   # global sizes; counts are global totals in the sharded path
    total_bytes = sum(ab.values())
    value = total_bytes * args.steps / (ms_total * 1e-3) / 1e9
    peak, peak_src = measured_peak()
    ops = {}
    for n in OP_ORDER:
        ms = per_op_ms[n] / args.steps
        gbs = ab[n] / (ms * 1e-3) / 1e9
        ops[n] = {"ms": round(ms, 4), "GBps": round(gbs, 1), "frac_of_peak_per_gpu": round(gbs / world / peak, 4)}
    span_bytes = sum(ab[n] for n in SPAN_OPS)
    span_ms = sum(per_op_ms[n] for n in SPAN_OPS) / args.steps
    achieved = span_bytes / world / (span_ms * 1e-3) / 1e9   # per GPU
    roofline = {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": None,
                "kernel": "span_kernel (whole-array / long-column streaming reduction): " + ", ".join(SPAN_OPS),
                "peak_source": peak_src, "frac_of_nominal_8000": round(achieved / 8000.0, 4),
                "note": "the peak is a COPY figure (read + write streams); a read-only stream reaches ~7.1 TB/s of DRAM throughput on "
                        "this part (ncu dram__throughput of span_kernel, profiles/r02a_ncu_full.md), so frac may exceed 1",
                "algorithmic_bytes_per_launch": round(span_bytes / world / len(SPAN_OPS)),
                "how": "mean over the 5 whole-array span_kernel launches of a step: algorithmic bytes / CUDA-event time of "
                       "those API calls on the library's stream (includes the scalar D2H); traffic = mean ncu dram read+write "
                       "bytes of the same 5 launches (profiles/r02_span_kernel_traffic.json)"}
    prof = os.path.join(ROOT, "profiles", "r02_span_kernel_traffic.json")
    if os.path.exists(prof):
        try:
            with open(prof) as f:
                t = json.load(f).get("traffic_bytes_per_launch")
                roofline["traffic"] = None if t is None else round(t / world)   # captured at N = 1; shards scale 1/N
        except Exception:
            pass

    # ---- the same suite through the BLOCKING reference calls (af_sum_all ... af_where: one host round trip per op) -------
    sync_api = None
    if pipelined and not args.no_sync_compare:
        suite_s, keep_s, _ = build_suite(af, torch, d, rank, world, group, False, ops=sh_ops)
        for _ in range(2):
            run_step(suite_s)
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        s0.record(lib_stream)
        for _ in range(args.steps):
            run_step(suite_s)
        s1.record(lib_stream)
        barrier()
        s_ms = s0.elapsed_time(s1)
        if world > 1:
            t = torch.tensor([s_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            s_ms = float(t[0])
        sync_api = {"value": round(total_bytes * args.steps / (s_ms * 1e-3) / 1e9, 1), "unit": "GB/s",
                    "ms_per_step": round(s_ms / args.steps, 4),
                    "how": "same suite, same inputs, every whole-array op through the blocking call (the host waits for each "
                           "scalar / hit count before it enqueues the next op)"}
        del suite_s, keep_s

    # ---- ops outside BASELINE's suite (SURVEY.md section 8f rows and the dim > 0 scan): timed on their own, NOT part of
    # `value`; N = 1 only.  Each: 2 warm-ups, then 5 launches bracketed by events on the library's stream.
    extras = None
    if world == 1 and not args.no_extras:
        A = af.Array
        m2 = A.from_device_ptr(d["m2_rows"].data_ptr(), d["m2_rows_dims"], af.f32, keepalive=d["m2_rows"])
        x3, u3, k4, v4 = A.from_torch(d["x3"]), A.from_torch(d["u3"]), A.from_torch(d["k4"]), A.from_torch(d["v4"])
        c2, c3, c4 = w["C2"]["rows"] * w["C2"]["cols"], w["C3"]["n"], w["C4"]["n"]
        xs = {
            "X.scan_dim1_f32 (C2 matrix, along dim 1)": (lambda: af.accum(m2, 1), 8 * c2),
            "X.sat_f32 (C2 matrix: scan dim 0 + scan dim 1)": (lambda: af.sat(m2), 16 * c2),
            "X.var_dim0_f32 (C2 matrix)": (lambda: af.var(m2, 0), 4 * c2),
            "X.mean_dim1_f32 (C2 matrix)": (lambda: af.mean(m2, 1), 4 * c2),
            "X.mean_all_f32 (C3)": (lambda: af.mean_all(x3), 4 * c3),
            "X.var_all_f32 (C3)": (lambda: af.var_all(x3), 4 * c3),
            "X.norm2_f32 (C3)": (lambda: af.norm(x3), 4 * c3),
            "X.corrcoef_f32 (C3 x, C3 u)": (lambda: af.corrcoef(x3, u3), 8 * c3),
            "X.sum_by_key_s32 (C4 keys + values)": (lambda: af.sum_by_key(k4, v4, 0), 8 * c4),
            "X.pi_count_f32 sum(sqrt(x*x+y*y)<1) (C3 x, C3 u), fused expression": (
                lambda: af.sum_all(af.lt(af.sqrt(af.add(af.mul(x3, x3), af.mul(u3, u3))), 1.0)), 8 * c3),
        }
        extras = {}
        for name, (fn, nbytes) in xs.items():
            try:
                for _ in range(2):
                    r = fn()
                    del r
                torch.cuda.synchronize()
                x0, x1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                x0.record(lib_stream)
                for _ in range(5):
                    r = fn()
                    del r
                x1.record(lib_stream)
                torch.cuda.synchronize()
                ms = x0.elapsed_time(x1) / 5
                extras[name] = {"ms": round(ms, 4), "GBps": round(nbytes / (ms * 1e-3) / 1e9, 1),
                                "frac_of_peak": round(nbytes / (ms * 1e-3) / 1e9 / peak, 4)}
            except Exception as ex:  # an extra never takes the headline down
                extras[name] = {"error": repr(ex)}
        del m2, x3, u3, k4, v4

    # ---- end-to-end arm (host buffers) ------------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        t0 = time.time()
        runner = E2E(af, torch, d, world, e2e_ops)
        log("[rank %d] pinned host copies ready in %.1f s" % (rank, time.time() - t0))
        ksteps = max(1, min(args.steps, 5))
        runner.step()   # warm the pool for the host-path allocations
        barrier()
        t0 = time.perf_counter()
        e_start, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e_start.record(lib_stream)
        for _ in range(ksteps):
            runner.step()
        e_end.record(lib_stream)
        barrier()
        e_ms = e_start.elapsed_time(e_end)
        if world > 1:
            t = torch.tensor([e_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e_ms = float(t[0])
        # under N > 1 every rank uploads its shard and the whole-array ops include the cross-GPU combine
        e_bytes = total_bytes
        e2e = {"value": round(e_bytes * ksteps / (e_ms * 1e-3) / 1e9, 2), "unit": "GB/s",
               "h2d_bytes_per_step": int(runner.h2d) * world, "d2h_bytes_per_step": int(runner.d2h) * world,
               "steps": ksteps, "ms_per_step": round(e_ms / ksteps, 2),
               "how": "pinned host inputs -> af_create_array (H2D, every distinct input once per step) -> op (N > 1: sharded op incl. the cross-GPU combine) -> "
                      "af_get_data_ptr / scalar (D2H), every step; blocking calls, array read-backs issued from one helper "
                      "thread so that they overlap the next upload (dedicated copy streams in the library)"}
        runner.reader.shutdown()
        del runner

    # ---- CPU baseline (rank 0, N=1 only) ----------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cpu, _, _ = run_cpu(2, 1, 1.0 / 8, "reference CPU path (serial, 1 worker thread like libafcpu)")
            cpu["value"] = round(cpu["value"], 3)
        except Exception as ex:  # the checker is optional for the GPU arm
            cpu = {"value": None, "unit": "GB/s", "cores": 1, "kind": "port", "sample": "failed: %r" % (ex,)}

    if rank == 0:
        line = {
            "metric": "effective_GBps_reduce_scan_where", "value": round(value, 1), "unit": "GB/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_total / args.steps, 4),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32/s32/u8",
            "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "scale": args.scale, "l2": "every input (>=256 MiB) exceeds the 126 MB L2 and ops are interleaved so "
                                                  "no input is re-read while resident",
                       "issue": ("pipelined: the 13 ops of a step are enqueued back to back (afx_*_async for the ops with a host "
                                 "result), the scalars / index / hit counts are collected from pinned host records at the end "
                                 "of the step") if pipelined else "blocking reference calls, one host round trip per op",
                       "sharding": "none" if world == 1 else "whole-array ops sharded along the flattened axis (global sizes "
                                   "fixed); the 64-byte per-shard partials are exchanged INSIDE the compute kernels over NVLink "
                                   "peer memory (symmetric buffers, kernels/peer_comm.cuh) and folded in rank order on every "
                                   "GPU -- no collective-library call on the data path (NCCL only carries the rendezvous, the "
                                   "barrier and the timing all-reduce); C2 partitioned along the non-reduced axis, no "
                                   "communication"},
            "parity_checked": parity_failed is not None and not parity_failed,
            "parity_failed_ops": parity_failed,
            "sync_api": sync_api,
            "extras": extras,
            "host_issue_ms_per_step": round(host_issue_ms, 4),
            "cpu_binding": numa,
            "bytes_per_step": int(total_bytes), "ops": ops, "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": int(launches), "clocks": clocks, "where_counts": counts,
        }
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0
