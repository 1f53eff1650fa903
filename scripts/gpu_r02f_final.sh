#!/bin/bash
# Final validation of the round on the box: every -m gpu test, the default bench line, the launch
# list of two eager steps.
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/r02f_pytest.log 2>&1
tail -9 gpurun_out/r02f_pytest.log
timeout 400 python bench.py > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err
python - <<EOF2
import json
try:
    d = json.loads(open("gpurun_out/r02f_bench.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print(d["value"], d["ms_per_step"], r["frac"], r["share_of_step"], d["e2e"]["value"], d["e2e"]["value_pageable_input"],
          d["latency_batch1_ms"], d["forward_plus_voting"]["value"], d["gpu_launches_per_step"], d["clocks"])
    print({k: (v.get("ms"), v.get("frac_of_hbm"), v.get("bit_exact")) for k, v in d["micro"].items()})
    print(d["cpu_baseline"]["value"], d["roofline_roi_warp"]["frac"])
except Exception as e:
    print("ERR", e)
    print(open("gpurun_out/r02f_bench.err").read()[-3000:])
EOF2
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "timed/" -c 400 --csv --log-file gpurun_out/r02f_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-micro --no-graph > /dev/null 2>&1
python scripts/launch_summary.py gpurun_out/r02f_launches.csv > gpurun_out/r02f_launches_summary.txt 2>&1; head -30 gpurun_out/r02f_launches_summary.txt
timeout 120 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02f_bench_reference_arm.json 2>/dev/null; tail -c 600 gpurun_out/r02f_bench_reference_arm.json
