#!/bin/bash
# one GPU, the per-rank problem sizes of N=8 (scale 1/8): per-op event times + ncu launch list (cold-cache kernel durations)
TAG=${1:-r2h}
mkdir -p gpurun_out
timeout 300 python bench.py --scale 0.125 --no-cpu-baseline --no-e2e --no-sync-compare --steps 20 > gpurun_out/bench_s8_$TAG.json 2> gpurun_out/bench_s8_$TAG.err; echo "rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_s8_$TAG.json"))
print(d["value"], d["ms_per_step"], d["parity_checked"])
for k,v in d["ops"].items(): print("   %-20s %8.4f ms %8.1f GB/s" % (k, v["ms"], v["GBps"]))
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_s8_$TAG.csv python bench.py --scale 0.125 --no-cpu-baseline --no-e2e --no-sync-compare --no-check --steps 2 --warmup 3 > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/launches_s8_$TAG.csv")) if len(r)>10]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value"); ui=hdr.index("Metric Unit")
import re
last=rows[-40:]
for r in last:
    print("%-60s %10s %s" % (re.sub(r"\(.*","",r[ki])[:60], r[vi], r[ui]))
PY
