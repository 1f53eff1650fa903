#!/bin/bash
# Pair-aware split-K model: every -m gpu test, default bench, batch-1 bench, launch lists of both.
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -q --durations=3 > gpurun_out/r02h_pytest.log 2>&1
tail -7 gpurun_out/r02h_pytest.log
timeout 400 python bench.py > gpurun_out/r02h_bench.json 2> gpurun_out/r02h_bench.err
timeout 200 python bench.py --batch 1 --streams 1 --no-cpu-baseline --no-micro > gpurun_out/r02h_bench_batch1.json 2> gpurun_out/r02h_b1.err
python - <<EOF2
import json
for f in ("gpurun_out/r02h_bench.json", "gpurun_out/r02h_bench_batch1.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d["roofline"]
        print(f, d["value"], d["ms_per_step"], r["frac"], d["e2e"]["value"], d["latency_batch1_ms"], d["forward_plus_voting"]["value"], d["gpu_launches_per_step"], d["clocks"])
        print([(s, m) for s, m in r["ms_by_launch_site"] if "100352" in s or s.startswith("300x") or "126x" in s])
    except Exception as e:
        print("ERR", f, e)
EOF2
tail -3 gpurun_out/r02h_bench.err gpurun_out/r02h_b1.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "timed/" -c 400 --csv --log-file gpurun_out/r02h_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-micro --no-graph > /dev/null 2>&1
python scripts/launch_summary.py gpurun_out/r02h_launches.csv > gpurun_out/r02h_launches_summary.txt 2>&1; head -12 gpurun_out/r02h_launches_summary.txt; tail -1 gpurun_out/r02h_launches_summary.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "timed/" -c 400 --csv --log-file gpurun_out/r02h_launches_batch1.csv python bench.py --batch 1 --steps 2 --warmup 3 --no-cpu-baseline --no-micro --no-graph > /dev/null 2>&1
python scripts/launch_summary.py gpurun_out/r02h_launches_batch1.csv > gpurun_out/r02h_launches_batch1_summary.txt 2>&1; head -12 gpurun_out/r02h_launches_batch1_summary.txt; tail -1 gpurun_out/r02h_launches_batch1_summary.txt
