#!/bin/bash
# round-2 first GPU pass: new parity tests (no -x: show every failure), the old suite, a device-only bench
TAG=${1:-r2a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi_$TAG.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_bykey.py tests/test_gpu_golden.py tests/test_cpp_wrapper.py -m gpu -q --timeout 600 > gpurun_out/pytest_new_$TAG.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_new_$TAG.log
tail -30 gpurun_out/pytest_new_$TAG.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -k "nan_variants or lifecycle or argument" > gpurun_out/pytest_new2_$TAG.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_new2_$TAG.log
tail -30 gpurun_out/pytest_new2_$TAG.log
timeout 1500 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -s --timeout 900 > gpurun_out/pytest_full_$TAG.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_full_$TAG.log
tail -30 gpurun_out/pytest_full_$TAG.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded.py -m gpu -x -q --timeout 600 -k "not nan_variants and not lifecycle" > gpurun_out/pytest_old_$TAG.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_old_$TAG.log
tail -8 gpurun_out/pytest_old_$TAG.log
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_$TAG.log
tail -3 gpurun_out/smoke_$TAG.log
timeout 600 python bench.py --no-e2e --no-cpu-baseline --steps 20 > gpurun_out/bench_quick_$TAG.json 2> gpurun_out/bench_quick_$TAG.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_quick_$TAG.json"))
print(d["value"], d["ms_per_step"])
for k,v in d["ops"].items(): print("  ",k,v)
PY
