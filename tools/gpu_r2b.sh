#!/bin/bash
TAG=${1:-r2b}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bykey.py tests/test_gpu_golden.py -m gpu -q --timeout 600 > gpurun_out/pytest_new_$TAG.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_new_$TAG.log
tail -12 gpurun_out/pytest_new_$TAG.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 600 -k "where" > gpurun_out/pytest_where_$TAG.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_where_$TAG.log
tail -5 gpurun_out/pytest_where_$TAG.log
timeout 600 python bench.py --no-e2e --no-cpu-baseline --steps 20 > gpurun_out/bench_quick_$TAG.json 2> gpurun_out/bench_quick_$TAG.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_quick_$TAG.json"))
print(d["value"], d["ms_per_step"])
for k,v in d["ops"].items(): print("  ",k,v)
PY
AFB200_PIPE_DEBUG=1 timeout 300 python tools/dbg_where.py 2>&1 | tail -30
