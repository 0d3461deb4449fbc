#!/bin/bash
# N-GPU box: multi-GPU parity (check_sharded) + bench at N (pipelined + blocking comparison + in-bench parity check)
TAG=${1:-r2m}; N=${2:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR tools/check_sharded.py > gpurun_out/check_sharded_n${N}_$TAG.log 2>&1; echo "rc=$?" >> gpurun_out/check_sharded_n${N}_$TAG.log
grep -v "^\*\|OMP_NUM\|^$" gpurun_out/check_sharded_n${N}_$TAG.log | tail -4
timeout 600 $TR bench.py --gpus $N --no-cpu-baseline --steps 30 > gpurun_out/bench_n${N}_$TAG.json 2> gpurun_out/bench_n${N}_$TAG.err; echo "bench n$N rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_n${N}_$TAG.json"))
print(d["value"], d["ms_per_step"], "e2e", d.get("e2e") and d["e2e"].get("value"), "sync", d["sync_api"]["value"], "parity", d.get("parity_checked"), d.get("parity_failed_ops"), "host_issue", d["host_issue_ms_per_step"])
for k,v in d["ops"].items(): print("   %-20s %8.4f ms %8.1f GB/s %.3f" % (k, v["ms"], v["GBps"], v["frac_of_peak_per_gpu"]))
PY
