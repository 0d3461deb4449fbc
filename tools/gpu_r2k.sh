#!/bin/bash
TAG=${1:-r2k}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_next.py tests/test_cpp_wrapper.py tests/test_gpu_parity.py -m gpu -x -q --timeout 900 -k "next or wrapper or scan or Scan or accum or sat or diff or unique" > gpurun_out/pytest_$TAG.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_$TAG.log
tail -12 gpurun_out/pytest_$TAG.log
out=gpurun_out/floor_$TAG.txt
: > $out
echo "== default" | tee -a $out; EXPS=24,26,28,29,30 python tools/floor.py "accum" "mean_all" "var" "norm2" "dim1" 2>&1 | tee -a $out
echo "== AFB200_NO_DIMSCAN=1" | tee -a $out; AFB200_NO_DIMSCAN=1 EXPS=24,26,28,29,30 python tools/floor.py "accum dim1" 2>&1 | tee -a $out
