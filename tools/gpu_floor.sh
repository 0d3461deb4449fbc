#!/bin/bash
TAG=${1:-r2i}
out=gpurun_out/floor_$TAG.txt
: > $out
echo "== default" | tee -a $out; python tools/floor.py 2>&1 | tee -a $out
echo "== AFB200_CLUSTER=1" | tee -a $out; AFB200_CLUSTER=1 python tools/floor.py sum_all count "dim0" 2>&1 | tee -a $out
echo "== AFB200_REDUCE_CTAS_PER_SM=4" | tee -a $out; AFB200_REDUCE_CTAS_PER_SM=4 python tools/floor.py sum_all count 2>&1 | tee -a $out
