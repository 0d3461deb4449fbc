#!/bin/bash
TAG=${1:-r2n}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_async.py tests/test_gpu_parity.py -m gpu -x -q --timeout 600 -k "async or threads or lifecycle or issue or pending or write" > gpurun_out/pytest_$TAG.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_$TAG.log
tail -4 gpurun_out/pytest_$TAG.log
timeout 900 python bench.py --no-cpu-baseline --no-sync-compare --no-extras --steps 10 > gpurun_out/bench_n1_$TAG.json 2> gpurun_out/bench_n1_$TAG.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_n1_$TAG.json"))
print(d["value"], d["ms_per_step"], "e2e", d["e2e"], "parity", d["parity_checked"])
PY
