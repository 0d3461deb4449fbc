#!/bin/bash
# N=2 box: full GPU test-suite, multi-GPU parity (fused and round-1 exchange paths), bench N=1 and N=2
TAG=${1:-r2c}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi_$TAG.txt
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 900 > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
tail -6 gpurun_out/pytest_gpu_$TAG.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR tools/check_sharded.py > gpurun_out/check_sharded_n2_fused_$TAG.log 2>&1; echo "rc=$?" >> gpurun_out/check_sharded_n2_fused_$TAG.log
tail -8 gpurun_out/check_sharded_n2_fused_$TAG.log
AFB200_NO_FUSED=1 timeout 300 $TR tools/check_sharded.py > gpurun_out/check_sharded_n2_exch_$TAG.log 2>&1; echo "rc=$?" >> gpurun_out/check_sharded_n2_exch_$TAG.log
tail -4 gpurun_out/check_sharded_n2_exch_$TAG.log
timeout 600 python bench.py --no-cpu-baseline --steps 20 > gpurun_out/bench_n1_$TAG.json 2> gpurun_out/bench_n1_$TAG.err; echo "bench n1 rc=$?"
timeout 600 $TR bench.py --gpus 2 --no-cpu-baseline --steps 20 > gpurun_out/bench_n2_$TAG.json 2> gpurun_out/bench_n2_$TAG.err; echo "bench n2 rc=$?"
AFB200_NO_FUSED=1 timeout 600 $TR bench.py --gpus 2 --no-e2e --no-cpu-baseline --steps 20 > gpurun_out/bench_n2_exch_$TAG.json 2> gpurun_out/bench_n2_exch_$TAG.err; echo "bench n2 exch rc=$?"
python - <<PY
import json
for f in ("bench_n1_$TAG","bench_n2_$TAG","bench_n2_exch_$TAG"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f))
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f, d["value"], d["ms_per_step"], d.get("e2e",{}) and d["e2e"].get("value"))
    for k,v in d["ops"].items(): print("   %-20s %8.4f ms %8.1f GB/s %.3f" % (k, v["ms"], v["GBps"], v["frac_of_peak_per_gpu"]))
PY
tail -5 gpurun_out/bench_n2_$TAG.err
