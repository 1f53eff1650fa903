#!/bin/bash
# Last validation of the round (split-K reduce loads unrolled): every -m gpu test + the default bench.
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -q > gpurun_out/r02i_pytest.log 2>&1
tail -3 gpurun_out/r02i_pytest.log
timeout 400 python bench.py > gpurun_out/r02i_bench.json 2> gpurun_out/r02i_bench.err
python - <<EOF2
import json
try:
    d = json.loads(open("gpurun_out/r02i_bench.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print(d["value"], d["ms_per_step"], r["frac"], d["e2e"]["value"], d["e2e"]["value_pageable_input"], d["latency_batch1_ms"],
          d["forward_plus_voting"]["value"], d["gpu_launches_per_step"], d["clocks"])
except Exception as e:
    print("ERR", e); print(open("gpurun_out/r02i_bench.err").read()[-2000:])
EOF2
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"splitk_reduce" -c 40 --csv --log-file gpurun_out/r02i_reduce_launches.csv python bench.py --batch 1 --steps 2 --warmup 3 --no-cpu-baseline --no-micro --no-graph --streams 1 > /dev/null 2>&1
python - <<EOF3
import csv
rows = [r for r in csv.reader(l for l in open("gpurun_out/r02i_reduce_launches.csv") if not l.startswith("==")) if len(r) > 5]
hdr = rows[0]; k = hdr.index("Kernel Name"); v = hdr.index("Metric Value"); g = hdr.index("Grid Size")
agg = {}
for r in rows[1:]:
    agg.setdefault((r[k][:40], r[g]), []).append(float(r[v].replace(",", "")))
for key, vals in agg.items():
    vals.sort(); print(key, len(vals), "median %.1f us" % (vals[len(vals) // 2] / 1000.0))
EOF3
