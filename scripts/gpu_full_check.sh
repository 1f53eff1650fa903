#!/bin/bash
# Full GPU validation on the box (run under gpurun from the repo root): every -m gpu test, then the
# default bench line.
timeout 600 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/r02_pytest_final.log 2>&1
tail -12 gpurun_out/r02_pytest_final.log
timeout 400 python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err
python - <<EOF2
import json
try:
    d = json.loads(open("gpurun_out/r02_bench_final.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print(d["value"], d["ms_per_step"], r["frac"], r["share_of_step"], r["tensor_work_factor"], d["e2e"]["value"], d["latency_batch1_ms"], d["clocks"])
    print(d.get("roofline_roi_warp"))
    print(d.get("cpu_baseline"))
except Exception as e:
    print("ERR", e)
    print(open("gpurun_out/r02_bench_final.err").read()[-3000:])
EOF2
# launch list of two bench steps (per-kernel durations, cold-cache and serialised: shares only)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "timed/" -c 400 --csv --log-file gpurun_out/r02_launches_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-micro --no-graph > /dev/null 2>&1
python scripts/launch_summary.py gpurun_out/r02_launches_final.csv > gpurun_out/r02_launches_final_summary.txt 2>&1; head -30 gpurun_out/r02_launches_final_summary.txt
