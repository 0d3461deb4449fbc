#!/bin/bash
# N-GPU box: async test, bench N=1 and N=$2 in pipelined mode (with the blocking comparison + parity check)
TAG=${1:-r2d}; N=${2:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_async.py tests/test_gpu_parity.py -k "async or reduce_all or ireduce or issue" -m gpu -x -q --timeout 600 > gpurun_out/pytest_async_$TAG.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_async_$TAG.log
tail -6 gpurun_out/pytest_async_$TAG.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR tools/check_sharded.py > gpurun_out/check_sharded_n${N}_$TAG.log 2>&1; echo "rc=$?" >> gpurun_out/check_sharded_n${N}_$TAG.log
tail -3 gpurun_out/check_sharded_n${N}_$TAG.log
timeout 600 python bench.py --no-cpu-baseline --steps 20 > gpurun_out/bench_n1_$TAG.json 2> gpurun_out/bench_n1_$TAG.err; echo "bench n1 rc=$?"
timeout 600 $TR bench.py --gpus $N --no-cpu-baseline --steps 20 > gpurun_out/bench_n${N}_$TAG.json 2> gpurun_out/bench_n${N}_$TAG.err; echo "bench n$N rc=$?"
python - <<PY
import json
for f in ("bench_n1_$TAG","bench_n${N}_$TAG"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f))
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f, d["value"], d["ms_per_step"], "e2e", d.get("e2e") and d["e2e"].get("value"), "sync", d.get("sync_api"), "parity", d.get("parity_checked"), d.get("parity_failed_ops"), "launches", d["gpu_launches"])
    for k,v in d["ops"].items(): print("   %-20s %8.4f ms %8.1f GB/s %.3f" % (k, v["ms"], v["GBps"], v["frac_of_peak_per_gpu"]))
PY
tail -5 gpurun_out/bench_n${N}_$TAG.err
