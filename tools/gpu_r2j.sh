#!/bin/bash
TAG=${1:-r2j}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_async.py tests/test_gpu_golden.py tests/test_gpu_bykey.py -m gpu -x -q --timeout 600 > gpurun_out/pytest_$TAG.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_$TAG.log
tail -5 gpurun_out/pytest_$TAG.log
out=gpurun_out/floor_$TAG.txt
: > $out
echo "== default" | tee -a $out; python tools/floor.py sum_all count dim 2>&1 | tee -a $out
echo "== AFB200_DIM_GSPLIT_CTAS=0" | tee -a $out; AFB200_DIM_GSPLIT_CTAS=0 python tools/floor.py dim1 2>&1 | tee -a $out
echo "== AFB200_DIM_GSPLIT_CTAS=4" | tee -a $out; AFB200_DIM_GSPLIT_CTAS=4 python tools/floor.py dim1 2>&1 | tee -a $out
echo "== AFB200_DIM_GSPLIT_CTAS=8" | tee -a $out; AFB200_DIM_GSPLIT_CTAS=8 python tools/floor.py dim1 2>&1 | tee -a $out
timeout 600 python bench.py --no-cpu-baseline --no-e2e --steps 20 > gpurun_out/bench_n1_$TAG.json 2> gpurun_out/bench_n1_$TAG.err; echo "bench n1 rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_n1_$TAG.json"))
print(d["value"], d["ms_per_step"], "sync", d["sync_api"]["value"], "parity", d["parity_checked"], d["parity_failed_ops"])
for k,v in d["ops"].items(): print("   %-20s %8.4f ms %8.1f GB/s %.3f" % (k, v["ms"], v["GBps"], v["frac_of_peak_per_gpu"]))
PY
