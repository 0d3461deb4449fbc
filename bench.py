#!/usr/bin/env python
"""bench.py -- effective GB/s of the reduce / scan / scan_by_key / where hot path on B200.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

One STEP = one pass of the hot path over one batch of synthetic input: every op of BASELINE.json's
configs C1..C5 once (the workload named in `config.workload`).  `value` is whole-job effective GB/s =
algorithmic bytes (compulsory input reads + output writes) of all ops of all steps divided
by the device time of the K timed steps, inputs resident in HBM.  `e2e` is the same metric through the
public C-ABI with HOST buffers (af_create_array H2D from pinned memory -> op -> result D2H) inside the
timed region.  `roofline` describes the dominant kernel (the whole-array span reduction), `ops` gives the
per-op breakdown, `cpu_baseline` times the reference's CPU implementation on a bounded sample.

With N > 1 (launched under torchrun, one rank per GPU) the whole-array ops C3/C4/C5 are sharded along the
flattened axis (strong scaling: the global sizes stay those of BASELINE.json), the 64-byte per-shard partials
are exchanged by one kernel over NVLink peer memory (afx_peer_exchange; NCCL all-gather as the fallback) and
folded in rank order (arrayfire_b200/sharded.py); the dimension-wise ops of C2 are partitioned along a
non-reduced axis with no communication.  The e2e arm then uploads every rank's shard from pinned host memory.

`--impl reference` times the reference's own CPU implementation (oracle/_ref when it was built from
/root/reference, else the oracle port) on rank 0, each step a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 0xAF00
WORKLOAD = ("C1..C5 suite of BASELINE.json, every op once per step: sum+inclusive scan 2^26 f32; sum/max dim0+dim1 8192x8192 f32; "
            "sum_all f32/s32 + imax_all f32 2^30; scan_by_key excl add 2^28 s32; where 50%/1% + count 2^30 u8")


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ------------------------------------------------------------------------------------------------
# workload description (global sizes; a rank processes 1/world of the sharded ones)
# ------------------------------------------------------------------------------------------------
def workload(scale=1.0):
    """scale < 1 shrinks every size (used for CPU samples and quick checks)."""
    def n(e):
        return max(1 << 12, int((1 << e) * scale))
    side = max(64, int(8192 * (scale ** 0.5)))
    return {
        "C1": {"n": n(26)},
        "C2": {"rows": side, "cols": side},
        "C3": {"n": n(30)},
        "C4": {"n": n(28)},
        "C5": {"n": n(30)},
    }


def algorithmic_bytes(w, counts=None):
    """bytes per op of one step.  counts: realised where-hit counts {name: count}."""
    c1, c2, c3, c4, c5 = w["C1"]["n"], w["C2"]["rows"] * w["C2"]["cols"], w["C3"]["n"], w["C4"]["n"], w["C5"]["n"]
    counts = counts or {}
    return {
        "C1.sum_f32": 4 * c1 + 4,
        "C1.scan_f32": 8 * c1,
        "C2.sum_dim0": 4 * c2 + 4 * w["C2"]["cols"],
        "C2.sum_dim1": 4 * c2 + 4 * w["C2"]["rows"],
        "C2.max_dim0": 4 * c2 + 4 * w["C2"]["cols"],
        "C2.max_dim1": 4 * c2 + 4 * w["C2"]["rows"],
        "C3.sum_f32": 4 * c3 + 4,
        "C3.sum_s32": 4 * c3 + 4,
        "C3.imax_f32": 4 * c3 + 8,
        "C4.scan_by_key_s32": 12 * c4,
        "C5.where_50": c5 + 4 * counts.get("C5.where_50", c5 // 2),
        "C5.where_01": c5 + 4 * counts.get("C5.where_01", c5 // 100),
        "C5.count": c5 + 4,
    }


# interleaved so that no input is re-read while it could still sit in the 126 MB L2
OP_ORDER = ["C2.sum_dim0", "C3.sum_f32", "C2.sum_dim1", "C3.sum_s32", "C2.max_dim0", "C3.imax_f32", "C2.max_dim1",
            "C4.scan_by_key_s32", "C1.sum_f32", "C5.where_50", "C1.scan_f32", "C5.where_01", "C5.count"]


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """samples SM clock + throttle reasons DURING the timed region: NVML in a thread (a timed region is only a
    few hundred ms, too short for `nvidia-smi -lms`), with nvidia-smi as the fallback."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []      # (sm_mhz, reasons bitmask, power_w)
        self.max_mhz = None
        self.stop_flag = False
        self.thread = None
        self.how = None

    def _nvml_loop(self, nv, h):
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                try:
                    pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
                except Exception:
                    pw = None
                self.rows.append((float(sm), int(rs), pw))
            except Exception:
                pass
            time.sleep(0.004)

    def _smi_loop(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                r = [c.strip() for c in out.strip().split(",")]
                bits = 0
                for col, bit in ((5, 0x8), (6, 0x40), (7, 0x20), (8, 0x4)):
                    if r[col].lower().startswith("active"):
                        bits |= bit
                self.rows.append((float(r[1]), bits, float(r[3])))
                self.max_mhz = float(r[2])
            except Exception:
                time.sleep(0.05)

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            idx = self.gpu
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:
                try:
                    idx = int(vis.split(",")[self.gpu])
                except Exception:
                    pass
            h = nv.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            self.how = "nvml"
            self.thread = threading.Thread(target=self._nvml_loop, args=(nv, h), daemon=True)
        except Exception:
            self.how = "nvidia-smi"
            self.thread = threading.Thread(target=self._smi_loop, daemon=True)
        self.thread.start()

    def stop(self):
        self.stop_flag = True
        if self.thread:
            self.thread.join(timeout=6)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["no samples"], "samples": 0}
        sm = [r[0] for r in self.rows]
        bits = 0
        for r in self.rows:
            bits |= r[1]
        names = []
        # NVML clocks-event-reason bits (nvml.h): 0x4 sw_power_cap, 0x8 hw_slowdown, 0x20 sw_thermal, 0x40 hw_thermal
        for bit, name in ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"),
                          (0x4, "sw_power_cap"), (0x80, "hw_power_brake_slowdown")):
            if bits & bit:
                names.append(name)
        pw = [r[2] for r in self.rows if r[2] is not None]
        return {"sm_mhz": float(np.median(sm)), "sm_min_mhz": float(min(sm)), "sm_max_mhz": self.max_mhz,
                "reasons": names, "samples": len(sm), "power_w_max": max(pw) if pw else None, "how": self.how}


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's CPU implementation on a bounded sample
# ------------------------------------------------------------------------------------------------
def cpu_inputs(w, rng):
    c2 = w["C2"]
    x1 = rng.integers(-4, 5, w["C1"]["n"]).astype(np.float32)
    m2 = np.asfortranarray(rng.integers(-4, 5, (c2["rows"], c2["cols"])).astype(np.float32))
    x3 = rng.integers(-4, 5, w["C3"]["n"]).astype(np.float32)
    i3 = rng.integers(-1000, 1001, w["C3"]["n"]).astype(np.int32)
    u3 = (rng.integers(0, 1 << 24, w["C3"]["n"]).astype(np.float32) / np.float32(1 << 24))
    n4 = w["C4"]["n"]
    v4 = rng.integers(-15, 15, n4).astype(np.int32)
    k4 = np.repeat(np.arange(n4 // 256 + 2, dtype=np.int32), rng.integers(1, 1025, n4 // 256 + 2))[:n4]
    n5 = w["C5"]["n"]
    nz = rng.integers(1, 256, n5).astype(np.uint8)
    m50 = np.where(rng.random(n5) < 0.5, nz, 0).astype(np.uint8)
    m01 = np.where(rng.random(n5) < 0.01, nz, 0).astype(np.uint8)
    return dict(x1=x1, m2=m2, x3=x3, i3=i3, u3=u3, v4=v4, k4=k4, m50=m50, m01=m01)


def cpu_step(orc, d):
    """one pass of the suite through the CPU checker's C-API level; returns where counts"""
    counts = {}
    orc.sum(d["m2"], 0); orc.sum_all(d["x3"]); orc.sum(d["m2"], 1); orc.sum_all(d["i3"])
    orc.max(d["m2"], 0); orc.imax_all(d["u3"]); orc.max(d["m2"], 1)
    orc.scan_by_key(d["k4"], d["v4"], 0, 0, False)
    orc.sum_all(d["x1"])
    counts["C5.where_50"] = int(orc.where(d["m50"]).size)
    orc.accum(d["x1"], 0)
    counts["C5.where_01"] = int(orc.where(d["m01"]).size)
    orc.count_all(d["m50"])
    return counts


def run_cpu(steps, warmup, scale, label):
    from oracle import pyoracle as po
    po.build()
    kind = "reference" if po.Oracle.available("reference") else "port"
    orc = po.AfOracle(kind)
    w = workload(scale)
    d = cpu_inputs(w, np.random.default_rng(SEED))
    counts = None
    for _ in range(warmup):
        counts = cpu_step(orc, d)
    t0 = time.perf_counter()
    for _ in range(steps):
        counts = cpu_step(orc, d)
    dt = time.perf_counter() - t0
    total = sum(algorithmic_bytes(w, counts).values())
    gbs = total * steps / dt / 1e9
    sample = "%s: every op of C1..C5 at 1/%d of the BASELINE sizes (%.0f MB algorithmic bytes per step), %d steps" % (
        label, round(1 / scale), total / 1e6, steps)
    return {"value": gbs, "unit": "GB/s", "cores": 1, "kind": kind, "sample": sample}, dt / steps * 1e3, w


def reference_scale(steps, warmup, budget_s=150.0, full_step_s=16.0):
    """largest power-of-two fraction of the BASELINE sizes (at most 1, at least 1/64) whose steps + warmup passes fit the
    time budget: the reference's serial CPU kernels stream ~1.5 GB/s, i.e. ~16 s for one full-size pass of the suite"""
    scale = 1.0
    while scale > 1.0 / 64 and (steps + warmup) * full_step_s * scale > budget_s:
        scale /= 2
    return scale


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    scale = args.scale if args.scale < 1.0 else reference_scale(args.steps, args.warmup)
    base, ms, w = run_cpu(args.steps, args.warmup, scale, "reference CPU path (serial, 1 worker thread like libafcpu)")
    line = {
        "impl": "reference", "metric": "effective_GBps_reduce_scan_where", "value": base["value"], "unit": "GB/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32/s32/u8", "data": "synthetic",
        # same op list as the GPU arm, every size multiplied by `scale` (1.0 = BASELINE sizes): NOT the same configuration
        # unless scale == 1
        "config": {"workload": WORKLOAD, "scale": scale, "same_config_as_gpu_arm": scale == 1.0, "sample": base["sample"]},
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="afb200")
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the workload (debugging only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--sync-ops", action="store_true", help="device-resident arm through the blocking calls (one host round trip per op)")
    ap.add_argument("--no-sync-compare", action="store_true", help="skip the extra timed pass through the blocking calls")
    ap.add_argument("--no-extras", action="store_true", help="skip the ops outside BASELINE's suite (timed on their own)")
    ap.add_argument("--no-check", action="store_true", help="skip the parity check of the warm-up step")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return reference_arm(args)

    from bench_gpu import gpu_arm  # GPU-only imports live there
    return gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
