#!/bin/bash
TAG=${1:-r2l}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_bykey.py tests/test_gpu_stats.py tests/test_gpu_memory.py tests/test_cpp_wrapper.py -m gpu -x -q --timeout 900 > gpurun_out/pytest_$TAG.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_$TAG.log
tail -6 gpurun_out/pytest_$TAG.log
timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-sync-compare --steps 10 > gpurun_out/bench_n1_$TAG.json 2> gpurun_out/bench_n1_$TAG.err; echo "bench n1 rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_n1_$TAG.json"))
print(d["value"], d["ms_per_step"], "parity", d["parity_checked"], d["parity_failed_ops"])
for k,v in (d.get("extras") or {}).items(): print("   %-50s %s" % (k, v))
PY
