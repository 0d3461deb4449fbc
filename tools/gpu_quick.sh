#!/bin/bash
# quick GPU pass: pipe-kernel tests first, then the full gpu suite, then a short device-only bench
TAG=${1:-q}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 200 -k "pipe" > gpurun_out/pytest_pipe_$TAG.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_pipe_$TAG.log
tail -5 gpurun_out/pytest_pipe_$TAG.log
timeout 900 python -m pytest tests -m gpu -x -q --timeout 200 > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
tail -5 gpurun_out/pytest_gpu_$TAG.log
timeout 600 python bench.py --no-e2e --no-cpu-baseline --steps 20 > gpurun_out/bench_quick_$TAG.json 2> gpurun_out/bench_quick_$TAG.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_quick_$TAG.json"))
print(d["value"], d["ms_per_step"])
for k,v in d["ops"].items(): print("  ",k,v)
PY
